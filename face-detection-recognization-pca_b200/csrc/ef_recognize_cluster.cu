// K2, single-kernel form: crops in, identities out.
//
// One thread-block CLUSTER of 4 CTAs owns a tile of 128 crops.  Each CTA streams one quarter of the pixel (K)
// dimension through TMA -> shared memory -> tcgen05.mma kind::i8 -> TMEM exactly like project_tc_kernel, so all SMs
// pull HBM although a 4096-crop batch has only 32 crop tiles.  The four partial int32 tiles are then exchanged through
// DISTRIBUTED SHARED MEMORY: CTA r of the cluster sums, for its 32 crops, the partials of all four CTAs (exact integer
// adds, 16-byte ld.shared::cluster), combines the digit planes into float64 features, and finds the nearest gallery
// row + threshold + label for those 32 crops.  No global accumulators, no atomics, no memset, no second launch: HBM
// traffic is the crops (once) + 24 B per crop out.
//
// Nearest gallery row (cosine metrics): a TENSOR-CORE FILTER followed by an exact float64 re-score.
//   * the normalised features p^ and the normalised gallery rows g^ are split into float16 hi + lo parts;
//     s~ = p_hi.g_hi + p_hi.g_lo + p_lo.g_hi is one tcgen05.mma kind::f16 (K = 3k) per 256 gallery rows, accumulated in
//     float32 in TMEM: |s~ - cos| <= kFilterEps (see the bound at kFilterEps);
//   * pass 0 finds the approximate maximum M per crop, pass 1 re-scores IN FLOAT64, with exactly the arithmetic of the
//     CUDA-core path (same fma order), every row with s~ >= M - 2 eps.  The exact arg-max (and every exact tie) is
//     among those rows, so the returned (score, index, label) are bit-identical to the full float64 scan -- typically
//     one or two rows per crop are re-scored instead of all of them;
//   * the float16 gallery image (prepared once per model in the canonical no-swizzle UMMA layout) streams through a
//     shared-memory ring with cp.async.bulk; the two TMEM score buffers alternate between the MMA and 8 scanning warps.
// The L2 metric (north-star extra) and shapes whose buffers do not fit keep the float64 scan over a shared-memory
// gallery tile.
//
// Covers k <= 32 and S*(k+1) <= 256 digit-plane columns (every shipped Gen-1 model and the k<=30 sklearn models at
// S = 8); other shapes use project_tc_kernel + the separate epilogue kernels.
//
// Replaces project_face_to_eigenspace + recognize_face (useless/scan.py:80-132) and scaler.transform + pca.transform +
// recognize_face_with_model (scan-template-v4.py:265-287) for a whole batch.
#include <climits>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kCluster = 4;                 // CTAs per crop tile = K splits
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;       // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 TMEM drain + sum of squares,
                                            // warp 6 gallery ring; warps 8..15 scan the filter scores
constexpr int QB = BLOCK_M / kCluster;      // crops finished by each CTA (one per lane)
static_assert(QB == 32, "one crop per lane");
constexpr int kScanWarps = 8;
constexpr int kGalTile = 256;               // gallery rows per filter MMA (UMMA N)
constexpr int kMaxRing = 8;
constexpr int kListCap = 256;               // re-score list entries per CTA (2 per scanning thread in the common case)
// Filter error bound.  Per component |a b - (a_hi b_hi + a_hi b_lo + a_lo b_hi)| <= 3 * 2^-22 |a b| + 2^-24 (float16
// hi/lo split, subnormal floor), summed with |a|,|b| <= 1 and Cauchy-Schwarz: < 2e-6 for k <= 32; the 3k <= 96 exact
// float16 products are accumulated in float32 with at most 2^-22 relative error per addition: < 2.3e-5.  5e-5 covers
// both with a margin; a too-large value only costs extra float64 re-scores, never correctness.
constexpr float kFilterEps = 5e-5f;

struct ClusterArgs {
  int B, D, NC, nc_pad, k, kq, S, kb_total, stages;
  int tmem_cols;
  const int32_t* col_exp;
  const double* bias;
  const double* sumsq_ext;   // precomputed weighted sum of squares (standardised models) or null
  int want_resid;
  double c0;
  const double* gp;          // prepared gallery [n][KR]
  const double* gnorm;
  const double* ginv;
  int n, tile_rows;
  const int32_t* labels;
  double threshold;
  double* out_proj;
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;
  double* out_resid;
  int* status;
  unsigned long long* probe;   // debugging aid (EF_TC_PROBE): [grid][32] globaltimer stamps
  // tensor-core filter
  int filter, kf, ring, g_tiles;
  const __half* gimg;          // [g_tiles][256 rows x kf] float16 image, swizzled K-major (swz_chunk_offset)
  // shared-memory offsets (bytes from the dynamic base), computed on the host
  int off_recv, off_ps, off_pe, off_gal, off_aimg, off_sh;
};

struct ClusterShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar;
  unsigned long long gal_full[kMaxRing];
  unsigned long long gal_empty[kMaxRing];
  unsigned long long score_full[2];
  unsigned long long score_empty[2];
  uint32_t tmem_base;
  int failed;
  int list_cnt, overflow;
  double pn[QB];
  double xu[QB];
  unsigned long long ssq_recv[kCluster][QB];       // written by the four CTAs of the cluster (DSMEM stores)
  float fmax_s[kScanWarps][QB];
  int list_L[kListCap], list_j[kListCap], list_label[kListCap];
  double list_key[kListCap], list_score[kListCap];
  int red_l[kScanWarps][QB];
  double red_s[kWarps][QB];
  double red_d[kWarps][QB];
  int red_i[kWarps][QB];
};

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  if (METRIC == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem));
}

// float16 hi / lo split of a value in [-1, 1]
__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// Exact float64 score of gallery row j for the crop in column L of pe; same fma order as the full float64 scan.
// key ranks the rows (cosine up to the positive factor 1/|p| for the Gen-1 rule), score is the value returned.
template <int METRIC, int KR>
__device__ __forceinline__ void exact_entry(const double* __restrict__ gp, const double* __restrict__ ginv,
                                            const double* __restrict__ gnorm, const int32_t* __restrict__ labels,
                                            const double* pe, int L, int j, double pn, double& key, double& score,
                                            int& label) {
  const double2* g = reinterpret_cast<const double2*>(gp + (size_t)j * KR);
  double d = 0.0;
#pragma unroll
  for (int c = 0; c < KR; c += 2) {
    const double2 gv = __ldg(g + (c >> 1));
    d = fma(pe[c * QB + L], gv.x, d);
    d = fma(pe[(c + 1) * QB + L], gv.y, d);
  }
  if (METRIC == EF_METRIC_COSINE_G1) {
    const double gi = __ldg(ginv + j), gn = __ldg(gnorm + j);
    key = d * gi;
    score = (pn == 0.0 || gn == 0.0) ? 0.0 : d / (pn * gn);       // useless/scan.py:70-77
  } else {
    key = d;
    score = d;
  }
  label = labels ? __ldg(labels + j) : j;
}

template <int METRIC, int KR>
__global__ void __launch_bounds__(kThreads, 1)
recognize_cluster_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                         const ClusterArgs a) {
  // 1024-byte alignment is required by the 128-byte swizzle atoms.  The array is used directly (no pointer rounding
  // through integers) so that the compiler keeps the shared address space and emits LDS/STS instead of generic loads.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  const int b_stage_bytes = a.nc_pad * BLOCK_K;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* sA = smem;                                            // [stages][128][128]
  uint8_t* sB = smem + (size_t)a.stages * A_STAGE_BYTES;         // [stages][nc_pad][128]
  int32_t* recv = reinterpret_cast<int32_t*>(smem + a.off_recv); // [4 source CTAs][nc_pad][32] partial sums of MY crops
  double* ps = reinterpret_cast<double*>(smem + a.off_ps);       // [KR][QB] features
  double* pe = reinterpret_cast<double*>(smem + a.off_pe);       // [KR][QB] features as the exact scorer uses them
  uint8_t* gal = smem + a.off_gal;                               // filter: ring of float16 gallery tiles
  double* gs = reinterpret_cast<double*>(gal);                   // float64 scan: [tile_rows][KR] + [tile_rows]
  double* gw = gs + (size_t)a.tile_rows * KR;
  uint8_t* aimg = smem + a.off_aimg;                             // filter: float16 A operand [128][kf]
  ClusterShared* sh = reinterpret_cast<ClusterShared*>(smem + a.off_sh);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                       // K quarter of this CTA
  const int m_tile = blockIdx.x / kCluster;
  const int kb0 = (int)((long long)a.kb_total * rank / kCluster);
  const int kb1 = (int)((long long)a.kb_total * (rank + 1) / kCluster);
  const bool fused_ssq = a.want_resid && a.sumsq_ext == nullptr;
  const bool filter = METRIC != EF_METRIC_L2 && a.filter != 0;
  const int row_bytes = a.kf * 2;
  const uint32_t gal_tile_bytes = (uint32_t)kGalTile * (uint32_t)row_bytes;
  const int n_seq = 2 * a.g_tiles;                               // filter: pass 0 + pass 1 over all gallery tiles
  // per-column constants of the feature combination, fetched before the main loop (off the critical path)
  int my_exp[3] = {0, 0, 0};
  double my_bias[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int c = warp + it * kWarps;
    if (c < a.kq) {
      my_exp[it] = a.col_exp[c];
      if (c < a.k) my_bias[it] = a.bias[c];
    }
  }

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], fused_ssq ? 5 : 1);
    }
    mbar_init(&sh->tmem_full_bar, 1);
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&sh->gal_full[s], 1);
      mbar_init(&sh->gal_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->score_full[s], 1);
      mbar_init(&sh->score_empty[s], kScanWarps);
    }
    sh->failed = 0;
    sh->list_cnt = 0;
    sh->overflow = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();                                 // "this CTA runs": awaited before the first DSMEM store below
  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as every CTA got here (its
  // CTAs start as ours retire); everything above touched only this CTA's own resources, everything below may read
  // buffers written by the previous kernel (the crops) or write buffers it reads or writes (the results).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 32 : nullptr;
  if (probe && tid == 0) probe[0] = globaltimer();

  auto load_gallery_tile = [&](int g0, int t0, int nthreads) {
    const int rows = min(a.tile_rows, a.n - g0);
    const char* src = reinterpret_cast<const char*>(a.gp + (size_t)g0 * KR);
    const int chunks = rows * KR / 2;
    for (int e = t0; e < chunks; e += nthreads) cp_async16(reinterpret_cast<char*>(gs) + e * 16, src + e * 16);
    if (METRIC == EF_METRIC_COSINE_G1)
      for (int e = t0; e < (rows + 1) / 2; e += nthreads)
        cp_async16(reinterpret_cast<char*>(gw) + e * 16, reinterpret_cast<const char*>(a.ginv + g0) + e * 16);
    asm volatile("cp.async.commit_group;\n" ::);
  };

  // ======================================================================= main loop (K quarter of this CTA)
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!mbar_wait(&sh->empty_bar[stage], phase ^ 1, failed)) break;
        mbar_arrive_expect_tx(&sh->full_bar[stage], (uint32_t)stage_bytes);
        tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, &tmap_x, &sh->full_bar[stage], kb * BLOCK_K, m_tile * BLOCK_M);
        tma_load_2d(sB + (size_t)stage * b_stage_bytes, &tmap_w, &sh->full_bar[stage], kb * BLOCK_K, 0);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
    cluster_wait();
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = umma_idesc_i8(a.nc_pad);
      bool ok = true;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
        tc_fence_after();
        if (probe && kb == kb0) probe[1] = globaltimer();
        const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
        const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
          umma_i8(tmem_base, umma_desc_sw128(a_addr + k * UMMA_K), umma_desc_sw128(b_addr + k * UMMA_K), idesc,
                  (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(&sh->empty_bar[stage]);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (ok) umma_commit(&sh->tmem_full_bar);
      if (probe) probe[2] = globaltimer();
    }
    __syncwarp();
    cluster_wait();
  } else if (warp < 6) {
    // TMEM lane group = warp % 4; the same warps compute the exact sum of squares from the staged crop tiles
    const int lane_group = warp & 3;
    const int row_in_tile = lane_group * 32 + lane;
    unsigned long long ssq = 0;
    bool ok = true;
    if (fused_ssq) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
        const uint4* line = reinterpret_cast<const uint4*>(sA + (size_t)stage * A_STAGE_BYTES + row_in_tile * BLOCK_K);
        unsigned int partial = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = line[(j + row_in_tile) & 7];
          partial = __dp4a(v.x, v.x, partial);
          partial = __dp4a(v.y, v.y, partial);
          partial = __dp4a(v.z, v.z, partial);
          partial = __dp4a(v.w, v.w, partial);
        }
        ssq += partial;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->empty_bar[stage]);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
    // drain the accumulator: rows of lane group q belong to CTA q of the cluster, which finishes those 32 crops --
    // PUSH the partial sums straight into its receive buffer (st.shared::cluster: no round-trip latency)
    if (ok && kb1 > kb0) ok = mbar_wait(&sh->tmem_full_bar, 0, failed);
    ok = __all_sync(0xffffffffu, ok);
    tc_fence_after();
    __syncwarp();
    cluster_wait();                                 // every CTA of the cluster has started: its shared memory exists
    const uint32_t dst = map_to_cta(smem_u32(recv) + (uint32_t)(((int)rank * a.nc_pad * 32 + lane) * 4),
                                    (uint32_t)lane_group);
    for (int c0 = 0; c0 < a.nc_pad; c0 += 16) {
      uint32_t v[16];
      if (ok && kb1 > kb0) {
        tmem_ld16(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) st_cluster_u32(dst + (uint32_t)(c0 + j) * 128u, v[j]);
    }
    st_cluster_u64(map_to_cta(smem_u32(&sh->ssq_recv[rank][lane]), (uint32_t)lane_group), ssq);
    tc_fence_before();
  } else {
    if (filter) {
      // warp 6: first fill of the float16 gallery ring while the crops stream
      if (warp == 6 && lane == 0) {
        const int first = min(a.ring, n_seq);
        for (int s = 0; s < first; ++s) {
          mbar_arrive_expect_tx(&sh->gal_full[s], gal_tile_bytes);
          bulk_load(gal + (size_t)s * gal_tile_bytes,
                    reinterpret_cast<const uint8_t*>(a.gimg) + (size_t)(s % a.g_tiles) * gal_tile_bytes, gal_tile_bytes,
                    &sh->gal_full[s]);
        }
      }
    } else {
      // warps 6..15: pull the first float64 gallery tile into shared memory while the crops stream
      load_gallery_tile(0, tid - 6 * 32, kThreads - 6 * 32);
    }
    __syncwarp();
    cluster_wait();
  }

  // ======================================================================= partial tiles have been exchanged (DSMEM)
  __syncthreads();
  if (probe && tid == 0) probe[3] = globaltimer();
  cluster_sync_all();                               // all four partial slabs of my 32 crops are in my receive buffer
  if (probe && tid == 0) probe[4] = globaltimer();
  const int b = m_tile * BLOCK_M + (int)rank * QB + lane;       // the crop this lane finishes
  const bool live = b < a.B;
  unsigned long long ssq_total = 0;
  if (warp == 0 && fused_ssq) {
#pragma unroll
    for (int q = 0; q < kCluster; ++q) ssq_total += sh->ssq_recv[q][lane];
  }
  // digit planes -> float64 features: exact integer sum over the four K quarters, small planes first
  for (int c = warp, it = 0; c < KR; c += kWarps, ++it) {
    double v = 0.0;
    if (c < a.kq) {
      int32_t plane[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        int sum = 0;
        if (s < a.S) {
          const int32_t* src = recv + (s * a.kq + c) * 32 + lane;
#pragma unroll
          for (int q = 0; q < kCluster; ++q) sum += src[q * a.nc_pad * 32];     // exact: |full-K sum| < 2^31
        }
        plane[s] = sum;
      }
      v = ldexp(ef::planes_to_double(plane), my_exp[it]);
    }
    if (c < a.k) {
      v -= my_bias[it];
      if (a.out_proj && live) a.out_proj[(size_t)b * a.k + c] = v;
    }
    ps[c * QB + lane] = c < a.k ? v : 0.0;        // padding columns (k .. KR) must be exact zeros
    if (c >= a.k && c < a.kq) sh->xu[lane] = v;    // residual column x . u
  }
  if (KR < a.kq && warp == 0) {                     // residual column beyond the padded feature count
    for (int c = KR; c < a.kq; ++c) {
      double v = 0.0;
      int32_t plane[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        int sum = 0;
        if (s < a.S) {
          const int32_t* src = recv + (s * a.kq + c) * 32 + lane;
#pragma unroll
          for (int q = 0; q < kCluster; ++q) sum += src[q * a.nc_pad * 32];     // exact: |full-K sum| < 2^31
        }
        plane[s] = sum;
      }
      sh->xu[lane] = ldexp(ef::planes_to_double(plane), a.col_exp[c]);
    }
  }
  __syncthreads();
  if (probe && tid == 0) probe[5] = globaltimer();

  // every thread derives the squared norm of ITS lane's crop (same fma order in all warps: bit-identical values)
  double n2 = 0.0;
  for (int c = 0; c < a.k; ++c) n2 = fma(ps[c * QB + lane], ps[c * QB + lane], n2);

  if (filter) {
    // ===================================================================== tensor-core filter + exact re-score
    const int KC = a.kf >> 3;                       // 16-byte chunks (8 halfs) per row
    {
      // A operand (float32 arithmetic is plenty: the filter is approximate by construction): rows r = lane + 32 q hold
      // crop `lane` (the four TMEM lane groups see the same 32 crops), K = [hi | hi | lo]; this thread writes the
      // 16-byte chunks (q, kc) = warp, warp + 16, ...
      const float rinv = n2 > 0.0 ? rsqrtf((float)n2) : 0.f;
      const int kc_log2 = 31 - __clz(KC);
      for (int e = warp; e < 4 * KC; e += kWarps) {
        const int qq = e >> kc_log2, kc = e & (KC - 1), r = lane + 32 * qq;
        __align__(16) __half h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int kk = kc * 8 + i;
          const int seg = kk >= 3 * a.k ? 3 : (kk >= 2 * a.k ? 2 : (kk >= a.k ? 1 : 0));
          __half hi = __float2half_rn(0.f), lo = hi;
          if (seg < 3) split_half((float)ps[(kk - seg * a.k) * QB + lane] * rinv, hi, lo);
          h[i] = seg < 2 ? hi : lo;
        }
        *reinterpret_cast<uint4*>(aimg + swz_chunk_offset(r, kc, row_bytes, BLOCK_M)) = *reinterpret_cast<const uint4*>(h);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (probe && tid == 0) probe[7] = globaltimer();
    if (warp == 0) {
      // gallery ring refills (the first `ring` tiles were requested while the crops streamed)
      if (lane == 0) {
        int slot = 0, t = a.ring % a.g_tiles;
        uint32_t ephase = 0;                       // parity of the (use - 1)-th completion of gal_empty[slot]
        for (int s = a.ring; s < n_seq; ++s) {
          if (!mbar_wait(&sh->gal_empty[slot], ephase, failed)) break;
          mbar_arrive_expect_tx(&sh->gal_full[slot], gal_tile_bytes);
          bulk_load(gal + (size_t)slot * gal_tile_bytes,
                    reinterpret_cast<const uint8_t*>(a.gimg) + (size_t)t * gal_tile_bytes, gal_tile_bytes,
                    &sh->gal_full[slot]);
          if (++slot == a.ring) { slot = 0; ephase ^= 1; }
          if (++t == a.g_tiles) t = 0;
        }
      }
      __syncwarp();
      // squared distance from face space (needs no square root)
      if (a.want_resid && live) {
        const double sq = fused_ssq ? (double)ssq_total : a.sumsq_ext[b];
        const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
        a.out_resid[b] = r > 0.0 ? r : 0.0;
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // single-thread issue loop, kept small: descriptors advance linearly with the shared-memory address, so the
        // descriptor of k-step ks in ring slot `slot` is a base value plus multiples of three constants
        const uint32_t idesc = umma_idesc_f16(kGalTile);
        const int n_ks = a.kf >> 4;
        const int swb = row_bytes < 128 ? row_bytes : 128;
        const int pa_log2 = swb == 128 ? 2 : (swb == 64 ? 1 : 0);           // k-steps per swizzle atom: 4, 2, 1
        const uint64_t adesc0 = umma_desc_swz(smem_u32(aimg), 0, row_bytes, BLOCK_M);
        const uint64_t bdesc0 = umma_desc_swz(smem_u32(gal), 0, row_bytes, kGalTile);
        const uint64_t a_atom = (uint64_t)((BLOCK_M * swb) >> 4), b_atom = (uint64_t)((kGalTile * swb) >> 4);
        const uint64_t slot_step = (uint64_t)(gal_tile_bytes >> 4);
        int slot = 0;
        uint32_t gphase = 0;
        uint64_t bslot = bdesc0;
        for (int s = 0; s < n_seq; ++s) {
          const int buf = s & 1;
          if (!mbar_wait(&sh->gal_full[slot], gphase, failed)) break;
          if (!mbar_wait(&sh->score_empty[buf], (uint32_t)(((s >> 1) & 1) ^ 1), failed)) break;
          tc_fence_after();
          if (probe && s < 4) probe[20 + s] = globaltimer();
          const uint32_t d_addr = tmem_base + (uint32_t)buf * kGalTile;
#pragma unroll 1
          for (int ks = 0; ks < n_ks; ++ks) {
            const uint64_t koff = (uint64_t)((ks & ((1 << pa_log2) - 1)) << 1);   // 32 bytes per k-step inside an atom
            const uint64_t katom = (uint64_t)(ks >> pa_log2);
            umma_f16(d_addr, adesc0 + koff + katom * a_atom, bslot + koff + katom * b_atom, idesc, ks > 0 ? 1u : 0u);
          }
          umma_commit(&sh->gal_empty[slot]);
          umma_commit(&sh->score_full[buf]);
          bslot += slot_step;
          if (++slot == a.ring) { slot = 0; gphase ^= 1; bslot = bdesc0; }
        }
      }
    } else if (warp < kWarps - kScanWarps) {
      // helper warps 2..7, off the critical path: the exact norm and the feature vectors as the exact scorer uses
      // them (divided by the norm for the sklearn rule) -- needed only when the re-score list is processed
      double pn = sqrt(n2);
      if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
      if (warp == 2) sh->pn[lane] = pn;
      for (int c = warp - 2; c < KR; c += kWarps - kScanWarps - 2) {
        double v = ps[c * QB + lane];
        if (METRIC == EF_METRIC_COSINE_SK) v = v / pn;
        pe[c * QB + lane] = v;
      }
      __threadfence_block();
      asm volatile("bar.arrive 3, 448;" ::: "memory");
    } else {
      // scanning warps: lane group q = warp % 4 (all groups hold the same 32 crops), 32 of the 256 columns each.
      // Pass 0: approximate maximum per crop.  Pass 1: rows within the filter band go to the re-score list.
      const int sw = warp - (kWarps - kScanWarps), stid = tid - (kWarps - kScanWarps) * 32;
      const int q = warp & 3, h = sw >> 2;
      const int col0 = (q * 2 + h) * 32;
      double best = -CUDART_INF, best_score = 0.0;
      int best_i = INT_MAX, best_label = -1;
      auto consider = [&](double key, double score, int label, int j) {
        if (better<METRIC>(key, j, best, best_i)) { best = key; best_score = score; best_label = label; best_i = j; }
      };
      float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F, thr = 0.f;
      bool ok = true;
      int t = -1, pass = 0;
      for (int s = 0; s < n_seq; ++s) {
        const int buf = s & 1, suse = s >> 1;
        if (++t == a.g_tiles) { t = 0; pass = 1; }
        if (pass == 1 && t == 0) {
          sh->fmax_s[sw][lane] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          asm volatile("bar.sync 2, 256;" ::: "memory");
          float M = sh->fmax_s[0][lane];
#pragma unroll
          for (int w = 1; w < kScanWarps; ++w) M = fmaxf(M, sh->fmax_s[w][lane]);
          thr = M - 2.f * kFilterEps;
          if (probe && stid == 0) probe[8] = globaltimer();
        }
        // warp-uniform health (tcgen05.ld and the named barriers need every thread); after a failure the loop keeps
        // running without touching the pipeline so that all scanning warps still meet at the barriers
        ok = __all_sync(0xffffffffu, ok && mbar_wait(&sh->score_full[buf], (uint32_t)(suse & 1), failed));
        if (!ok) continue;
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kGalTile + col0), v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->score_empty[buf]);
        if (probe && stid == 0 && s < 4) probe[28 + s] = globaltimer();
        const int j0 = t * kGalTile + col0;
        const int valid = min(32, a.n - j0);       // columns of this slice that are gallery rows (warp uniform)
        if (valid <= 0) continue;
        if (pass == 0) {
          if (valid == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < valid) m0 = fmaxf(m0, __uint_as_float(v[i]));
          }
        } else {
          unsigned mask = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
          if (valid < 32) mask &= (1u << valid) - 1u;
          while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int slot = atomicAdd(&sh->list_cnt, 1);
            if (slot < kListCap) {
              sh->list_L[slot] = lane;
              sh->list_j[slot] = j0 + i;
            } else {
              sh->overflow = 1;                     // more survivors than list entries: full float64 scan below
            }
          }
        }
        if (probe && stid == 0 && s < 4) probe[16 + s] = globaltimer();
      }
      __threadfence_block();
      asm volatile("bar.sync 3, 448;" ::: "memory");      // list complete; pn / pe ready (helper warps)
      if (probe && stid == 0) probe[10] = globaltimer();
      const bool overflow = *reinterpret_cast<volatile int*>(&sh->overflow) != 0;
      const int total = overflow ? 0 : *reinterpret_cast<volatile int*>(&sh->list_cnt);
      // ---- exact float64 scores of the surviving rows, one per thread (all L2 reads in flight together)
      for (int e = stid; e < total; e += kScanWarps * 32) {
        const int L = sh->list_L[e];
        double key, score; int label;
        exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe, L, sh->list_j[e], sh->pn[L], key, score, label);
        sh->list_key[e] = key;
        sh->list_score[e] = score;
        sh->list_label[e] = label;
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (probe && stid == 0) probe[11] = globaltimer();
      for (int e = sw; e < total; e += kScanWarps)
        if (sh->list_L[e] == lane) consider(sh->list_key[e], sh->list_score[e], sh->list_label[e], sh->list_j[e]);
      if (overflow) {
        // degenerate gallery (hundreds of rows within the filter band of one crop tile): exact scan of every row
        const double pn = sh->pn[lane];
        for (int j = sw; j < a.n; j += kScanWarps) {
          double key, score; int label;
          exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe, lane, j, pn, key, score, label);
          consider(key, score, label, j);
        }
      }
      sh->red_s[sw][lane] = best;
      sh->red_d[sw][lane] = best_score;
      sh->red_i[sw][lane] = best_i;
      sh->red_l[sw][lane] = best_label;
      if (probe && stid == 0) probe[12] = globaltimer();
    }
    __syncthreads();
    if (warp == 0 && live) {
      double bs = sh->red_s[0][lane], score = sh->red_d[0][lane];
      int bi = sh->red_i[0][lane], bl = sh->red_l[0][lane];
      for (int w = 1; w < kScanWarps; ++w)
        if (better<METRIC>(sh->red_s[w][lane], sh->red_i[w][lane], bs, bi)) {
          bs = sh->red_s[w][lane];
          score = sh->red_d[w][lane];
          bi = sh->red_i[w][lane];
          bl = sh->red_l[w][lane];
        }
      if (bi == INT_MAX) { bi = 0; bl = -1; }       // only after a pipeline failure (the status flag is raised below)
      a.out_score[b] = score;
      a.out_index[b] = bi;
      if (a.out_label) a.out_label[b] = score >= a.threshold ? bl : -1;
    }
  } else {
    // ===================================================================== float64 scan of a shared-memory gallery tile
    double pn = sqrt(n2);
    if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
    if (warp == 0 && a.want_resid && live) {
      const double sq = fused_ssq ? (double)ssq_total : a.sumsq_ext[b];
      const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
      a.out_resid[b] = r > 0.0 ? r : 0.0;
    }
    double best = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF, best_dot = 0.0;
    int best_i = INT_MAX;
    double p[KR];
#pragma unroll
    for (int c = 0; c < KR; ++c) {
      double v = ps[c * QB + lane];
      if (METRIC == EF_METRIC_COSINE_SK) v = v / pn;
      p[c] = v;
    }
    for (int g0 = 0; g0 < a.n; g0 += a.tile_rows) {
      const int rows = min(a.tile_rows, a.n - g0);
      if (g0 > 0) {
        __syncthreads();
        load_gallery_tile(g0, tid, kThreads);
      }
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();
      const int per = ((rows + kWarps - 1) / kWarps + 3) & ~3;
      const int r_begin = warp * per, r_end = min(rows, r_begin + per);
      for (int r = r_begin; r < r_end; r += 4) {
        double d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < KR; c += 2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const double2 g = *reinterpret_cast<const double2*>(gs + (size_t)(r + j) * KR + c);
            if (METRIC == EF_METRIC_L2) {
              const double t0 = p[c] - g.x, t1 = p[c + 1] - g.y;
              d[j] = fma(t0, t0, d[j]);
              d[j] = fma(t1, t1, d[j]);
            } else {
              d[j] = fma(p[c], g.x, d[j]);
              d[j] = fma(p[c + 1], g.y, d[j]);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (r + j >= r_end) break;
          const double s = (METRIC == EF_METRIC_COSINE_G1) ? d[j] * gw[r + j] : d[j];
          if (better<METRIC>(s, g0 + r + j, best, best_i)) {
            best = s;
            best_dot = d[j];
            best_i = g0 + r + j;
          }
        }
      }
    }
    sh->red_s[warp][lane] = best;
    sh->red_d[warp][lane] = best_dot;
    sh->red_i[warp][lane] = best_i;
    __syncthreads();
    if (warp == 0 && live) {
      double bs = sh->red_s[0][lane], bd = sh->red_d[0][lane];
      int bi = sh->red_i[0][lane];
      for (int w = 1; w < kWarps; ++w)
        if (better<METRIC>(sh->red_s[w][lane], sh->red_i[w][lane], bs, bi)) {
          bs = sh->red_s[w][lane];
          bd = sh->red_d[w][lane];
          bi = sh->red_i[w][lane];
        }
      double score = bs;
      if (METRIC == EF_METRIC_COSINE_G1) {
        const double gn = a.gnorm[bi];
        score = (pn == 0.0 || gn == 0.0) ? 0.0 : bd / (pn * gn);       // useless/scan.py:70-77
      }
      a.out_score[b] = score;
      a.out_index[b] = bi;
      if (a.out_label) {
        const bool pass = METRIC == EF_METRIC_L2 ? score <= a.threshold : score >= a.threshold;
        a.out_label[b] = pass ? (a.labels ? a.labels[bi] : bi) : -1;
      }
    }
  }

  // ======================================================================= teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
  if (probe && tid == 0) probe[6] = globaltimer();
}

// gallery rows -> float16 [g_hi | g_lo | g_hi] image, 256-row tiles in the swizzled K-major layout of swz_chunk_offset
__global__ void gallery_image_kernel(const double* __restrict__ gp, int kr, const double* __restrict__ ginv, int n, int k,
                                     int kf, int metric, __half* __restrict__ img) {
  const int KC = kf >> 3;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * KC) return;
  const int j = (int)(e / KC), kc = (int)(e - (long long)j * KC);
  const double scale = metric == EF_METRIC_COSINE_G1 ? ginv[j] : 1.0;
  __align__(16) __half h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = kc * 8 + i;
    const int seg = kk >= 3 * k ? 3 : (kk >= 2 * k ? 2 : (kk >= k ? 1 : 0));
    __half hi = __float2half_rn(0.f), lo = hi;
    if (seg < 3) {
      const double v = gp[(size_t)j * kr + (kk - seg * k)] * scale;
      hi = __double2half(v);
      lo = __double2half(v - (double)__half2float(hi));
    }
    h[i] = seg == 1 ? lo : hi;
  }
  const int tile = j / kGalTile, rr = j - tile * kGalTile;
  uint8_t* dst = reinterpret_cast<uint8_t*>(img) + (size_t)tile * kGalTile * kf * 2 +
                 swz_chunk_offset(rr, kc, kf * 2, kGalTile);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

template <int METRIC, int KR>
int launch_cluster(const CUtensorMap& mx, const CUtensorMap& mw, ClusterArgs& a, int m_tiles, cudaStream_t stream) {
  const int stage_bytes = A_STAGE_BYTES + a.nc_pad * BLOCK_K;
  // pipeline depth: 3 stages are enough (the main loop is throughput bound, see tools/tc_probe.py)
  int stages = 3;
  a.stages = stages;
  size_t off = (size_t)stages * stage_bytes;
  a.off_recv = (int)off; off += (size_t)a.nc_pad * 512;           // [4][nc_pad][32] int32
  a.off_ps = (int)off;   off += sizeof(double) * KR * QB;
  a.off_pe = (int)off;   off += sizeof(double) * KR * QB;
  off = (size_t)ef::round_up((int64_t)off, 1024);
  const size_t tail = (size_t)ef::round_up((int64_t)sizeof(ClusterShared), 128) + 128;
  bool filter = METRIC != EF_METRIC_L2 && a.gimg != nullptr && getenv("EF_NO_FILTER") == nullptr;
  if (filter) {
    const size_t aimg_bytes = (size_t)ef::round_up((int64_t)BLOCK_M * a.kf * 2, 1024);
    const size_t tile_bytes = (size_t)kGalTile * a.kf * 2;
    const size_t left = (size_t)kSmemLimit > off + aimg_bytes + tail ? (size_t)kSmemLimit - off - aimg_bytes - tail : 0;
    int ring = (int)std::min<size_t>(kMaxRing, left / tile_bytes);
    ring = std::min(ring, 2 * a.g_tiles);
    if (ring >= 2) {
      a.ring = ring;
      a.off_aimg = (int)off; off += aimg_bytes;
      a.off_gal = (int)off;  off += (size_t)ring * tile_bytes;
      a.tile_rows = 0;
      a.tmem_cols = 512;
    } else {
      filter = false;
    }
  }
  a.filter = filter ? 1 : 0;
  if (!filter) {
    const size_t left = (size_t)kSmemLimit > off + tail ? (size_t)kSmemLimit - off - tail : 0;
    int rows = (int)(left / (sizeof(double) * (KR + 1)));
    rows &= ~3;
    const int n4 = (a.n + 3) & ~3;
    if (rows > n4) rows = n4;
    if (rows < 64 && rows < n4) return EF_ERR_UNSUPPORTED;
    a.tile_rows = rows;
    a.off_gal = (int)off;  off += sizeof(double) * (size_t)rows * (KR + 1);
    a.off_aimg = a.off_gal;
    a.ring = 1;
  }
  off = (size_t)ef::round_up((int64_t)off, 128);
  a.off_sh = (int)off;
  const size_t smem = off + sizeof(ClusterShared);
  if (smem > (size_t)kSmemLimit) return EF_ERR_UNSUPPORTED;
  EF_ENSURE_SMEM((recognize_cluster_kernel<METRIC, KR>), smem);
  static unsigned long long* probe_buf = nullptr;
  const bool probing = getenv("EF_TC_PROBE") != nullptr;
  const int grid_n = m_tiles * kCluster;
  a.probe = nullptr;
  if (probing && grid_n <= 4096) {
    if (!probe_buf) EF_CUDA(cudaMalloc(&probe_buf, sizeof(unsigned long long) * 32 * 4096));
    EF_CUDA(cudaMemsetAsync(probe_buf, 0, sizeof(unsigned long long) * 32 * 4096, stream));
    a.probe = probe_buf;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(m_tiles * kCluster));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCluster;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = getenv("EF_NO_PDL") ? 1 : 2;
  EF_CUDA(cudaLaunchKernelEx(&cfg, recognize_cluster_kernel<METRIC, KR>, mx, mw, a));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (a.probe) {
    std::vector<unsigned long long> h((size_t)grid_n * 32);
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid_n; ++c) if (h[(size_t)c * 32] && h[(size_t)c * 32] < t0) t0 = h[(size_t)c * 32];
    const char* names[32] = {"start", "first_full", "mma_issued", "block_done", "sync1", "features", "end",
                             "aimg", "max_known", "-", "listed", "exact", "gathered", "-", "-", "-",
                             "scan_t0", "scan_t1", "scan_t2", "scan_t3", "mma_t0", "mma_t1", "mma_t2", "mma_t3",
                             "-", "-", "-", "-", "ld_t0", "ld_t1", "ld_t2", "ld_t3"};
    fprintf(stderr, "[ef_cluster_probe] grid %d stages %d filter %d ring %d tile_rows %d; us since first CTA start (mean/max):",
            grid_n, a.stages, a.filter, a.ring, a.tile_rows);
    for (int i = 0; i < 32; ++i) {
      if (names[i][0] == '-') continue;
      double sum = 0, mx = 0;
      for (int c = 0; c < grid_n; ++c) {
        const double v = h[(size_t)c * 32 + i] ? (double)(h[(size_t)c * 32 + i] - t0) * 1e-3 : 0.0;
        sum += v;
        if (v > mx) mx = v;
      }
      fprintf(stderr, " %s %.2f/%.2f", names[i], sum / grid_n, mx);
    }
    fprintf(stderr, "\n");
  }
  return EF_OK;
}

template <int METRIC>
int dispatch_kr(const CUtensorMap& mx, const CUtensorMap& mw, ClusterArgs& a, int kr, int m_tiles, cudaStream_t st) {
  switch (kr) {
    case 4: return launch_cluster<METRIC, 4>(mx, mw, a, m_tiles, st);
    case 8: return launch_cluster<METRIC, 8>(mx, mw, a, m_tiles, st);
    case 12: return launch_cluster<METRIC, 12>(mx, mw, a, m_tiles, st);
    case 16: return launch_cluster<METRIC, 16>(mx, mw, a, m_tiles, st);
    case 24: return launch_cluster<METRIC, 24>(mx, mw, a, m_tiles, st);
    default: return launch_cluster<METRIC, 32>(mx, mw, a, m_tiles, st);
  }
}

}  // namespace

namespace ef {

// float16 K extent of the filter operands: [hi | lo | hi] = 3k, padded to a power of two (one swizzle row)
int filter_kf(int k) {
  int kf = 16;
  while (kf < 3 * k) kf *= 2;
  return kf;
}

size_t gallery_image_bytes(int k, int64_t n) {
  return (size_t)ceil_div(n, kGalTile) * kGalTile * (size_t)filter_kf(k) * 2;
}

// float16 filter image of a prepared gallery (gp [n][kr], rows normalised for COSINE_SK; ginv = 1/|g| for COSINE_G1).
// img must hold gallery_image_bytes(k, n) bytes.
int gallery_image(const double* gp, int kr, const double* ginv, int64_t n, int k, int metric, void* img,
                  cudaStream_t stream) {
  if (n <= 0) return EF_OK;
  const int kf = filter_kf(k);
  EF_CUDA(cudaMemsetAsync(img, 0, gallery_image_bytes(k, n), stream));
  const int64_t work = n * (kf >> 3);
  EF_LAUNCH(gallery_image_kernel, (unsigned)ceil_div(work, 256), 256, 0, stream, gp, kr, ginv, (int)n, k, kf, metric,
            reinterpret_cast<__half*>(img));
  return EF_OK;
}

// EF_ERR_UNSUPPORTED when the shape / alignment is outside what the single-kernel form covers (caller falls back).
int recognize_cluster(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows,
                      int k, int kq, int S, const int32_t* col_exp, const double* bias, const double* sumsq_ext,
                      bool want_resid, double c0, const double* gp_padded, int kpad, const double* gnorm,
                      const double* ginv, const void* gimg, int64_t n, const int32_t* labels, int metric,
                      double threshold, double* out_proj, double* out_score, int32_t* out_index, int32_t* out_label,
                      double* out_resid, int* status, cudaStream_t stream) {
  using namespace ef_tc;
  if (B <= 0) return EF_OK;
  if (k > 32 || kpad != fused_epilogue_kpad(k) || n <= 0 || n >= (1ll << 31) - 512) return EF_ERR_UNSUPPORTED;
  const int nc_pad = (int)round_up(NC, 16);
  if (nc_pad > 256 || nc_pad > wq_rows) return EF_ERR_UNSUPPORTED;
  if ((ldx & 15) || (reinterpret_cast<uintptr_t>(X) & 15) || (ldw & 15) || (reinterpret_cast<uintptr_t>(Wq) & 15))
    return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  ClusterArgs a{};
  a.B = B; a.D = D; a.NC = NC; a.nc_pad = nc_pad; a.k = k; a.kq = kq; a.S = S;
  a.kb_total = (int)ceil_div(D, BLOCK_K);
  a.tmem_cols = 32;
  while (a.tmem_cols < nc_pad) a.tmem_cols *= 2;
  a.col_exp = col_exp; a.bias = bias; a.sumsq_ext = sumsq_ext; a.want_resid = want_resid ? 1 : 0; a.c0 = c0;
  a.gp = gp_padded; a.gnorm = gnorm; a.ginv = ginv; a.n = (int)n; a.labels = labels; a.threshold = threshold;
  a.out_proj = out_proj; a.out_score = out_score; a.out_index = out_index; a.out_label = out_label;
  a.out_resid = out_resid; a.status = status;
  a.gimg = reinterpret_cast<const __half*>(gimg);
  a.kf = filter_kf(k);
  a.g_tiles = (int)ceil_div(n, kGalTile);
  CUtensorMap mx, mw;
  if (!make_map(&mx, X, (uint64_t)D, (uint64_t)B, (uint64_t)ldx, BLOCK_M)) return EF_ERR_UNSUPPORTED;
  if (!make_map(&mw, Wq, (uint64_t)ldw, (uint64_t)wq_rows, (uint64_t)ldw, (uint32_t)nc_pad)) return EF_ERR_UNSUPPORTED;
  const int m_tiles = (int)ceil_div(B, BLOCK_M);
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(mx, mw, a, kpad, m_tiles, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(mx, mw, a, kpad, m_tiles, stream);
    case EF_METRIC_L2: return dispatch_kr<EF_METRIC_L2>(mx, mw, a, kpad, m_tiles, stream);
    default: return EF_ERR_INVALID;
  }
}

}  // namespace ef
