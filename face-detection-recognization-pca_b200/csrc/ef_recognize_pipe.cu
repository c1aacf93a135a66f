// K2, software-pipelined across batches: the streaming half of batch i and the matching half of batch i-1 in ONE launch.
//
// recognize_cluster_kernel (ef_recognize_cluster.cu) spends ~6 us streaming a 4096-crop batch at the HBM roofline and
// then ~10 us in dependent latency chains (cluster exchange, float64 features, tensor-core filter, exact re-score) during
// which HBM idles.  A serving loop submits batch after batch, so this kernel runs the two halves of CONSECUTIVE batches
// side by side on different warps of the same CTAs:
//   warps 0..5   stream half, batch i:   TMA -> tcgen05.mma kind::i8 -> TMEM -> push to the owning CTA (DSMEM) ->
//                float64 features.  Results written now: features, reconstruction error.  Carried to the next launch
//                (global memory): the exact-scorer vector pe, the norm, the float16 [hi|hi|lo] filter row.
//   warps 6..15  match half, batch i-1:  carried rows -> A operand, tcgen05.mma kind::f16 filter over the gallery ring,
//                two scanning passes, exact float64 re-score, arg-best.  Results written now: score, index, label of
//                batch i-1.
// The launch therefore costs max(stream half, match half) instead of their sum.  The arithmetic of both halves is the
// one of recognize_cluster_kernel (same integer sums, same fma order, same filter and band), so every output is bit
// identical to the unpipelined kernel; only WHEN the score / index / label of a batch appear changes (one launch later,
// or at ef_model_flush_device).  Covers the cosine metrics with k <= 21 (one 128-byte filter row) and <= 128 digit-plane
// columns; other shapes use recognize_cluster_kernel.
//
// Replaces, like recognize_cluster_kernel, project_face_to_eigenspace + recognize_face (useless/scan.py:80-132) and
// scaler.transform + pca.transform + recognize_face_with_model (scan-template-v4.py:265-287) for a stream of batches.
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

constexpr int kCluster = 4;
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int QB = BLOCK_M / kCluster;      // crops finished by each CTA (one per lane)
constexpr int kStreamWarps = 6;             // warp 0 TMA, warp 1 MMA, warps 2..5 drain; all six combine the features
constexpr int kScanWarps = 8;               // warps 8..15
constexpr int kGalTile = 128;               // gallery rows per filter MMA
constexpr int kScoreBufs = 3;               // 128-column TMEM score buffers (columns 128..511; the accumulator has 0..127)
constexpr int kMaxRing = 8;
constexpr int kListCap = 256;
constexpr float kFilterEps = 5e-5f;         // same bound as recognize_cluster_kernel

struct PipeArgs {
  // ---- stream half: the batch submitted with this launch
  int B, D, NC, nc_pad, k, kq, S, kb_total, stages;
  const int32_t* col_exp;
  const double* bias;
  const double* sumsq_ext;
  int want_resid;
  double c0;
  double* out_proj;
  double* out_resid;
  double* carry_pe;            // [B][KR]  exact-scorer vector of every crop
  double* carry_pn;            // [B]      norm
  __half* carry_img;           // [B][kf]  float16 filter row [hi | hi | lo]
  // ---- match half: the batch submitted with the previous launch
  int Bp;
  const double* prev_pe;
  const double* prev_pn;
  const __half* prev_img;
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;
  double threshold;
  // ---- model
  const double* gp;
  const double* gnorm;
  const double* ginv;
  const int32_t* labels;
  int n, kf, ring, g_tiles;
  const __half* gimg;
  int* status;
  int off_recv, off_ps, off_pe, off_aimg, off_gal, off_sh;
  unsigned long long* probe;   // debugging aid (EF_TC_PROBE): [grid][16] globaltimer stamps
};

struct PipeShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar;
  unsigned long long gal_full[kMaxRing];
  unsigned long long gal_empty[kMaxRing];
  unsigned long long score_full[kScoreBufs];
  unsigned long long score_empty[kScoreBufs];
  unsigned long long aimg_ready;
  uint32_t tmem_base;
  int failed;
  int list_cnt, overflow;
  double pn[QB];
  double xu[QB];
  unsigned long long ssq_recv[kCluster][QB];
  float fmax_s[kScanWarps][QB];
  int list_L[kListCap], list_j[kListCap], list_label[kListCap];
  double list_key[kListCap], list_score[kListCap];
  int red_l[kScanWarps][QB];
  double red_s[kScanWarps][QB];
  double red_d[kScanWarps][QB];
  int red_i[kScanWarps][QB];
};

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  return s > bs || (s == bs && i < bi);
}

__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

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

__device__ __forceinline__ void bar_stream() { asm volatile("bar.sync 4, 192;" ::: "memory"); }
__device__ __forceinline__ void bar_scan() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

template <int METRIC, int KR>
__global__ void __launch_bounds__(kThreads, 1)
recognize_pipe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const PipeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  const int b_stage_bytes = a.nc_pad * BLOCK_K;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)a.stages * A_STAGE_BYTES;
  int32_t* recv = reinterpret_cast<int32_t*>(smem + a.off_recv);
  double* ps = reinterpret_cast<double*>(smem + a.off_ps);       // [KR][QB] features of the current batch
  double* pe = reinterpret_cast<double*>(smem + a.off_pe);       // [KR][QB] exact-scorer vectors of the previous batch
  uint8_t* aimg = smem + a.off_aimg;
  uint8_t* gal = smem + a.off_gal;
  PipeShared* sh = reinterpret_cast<PipeShared*>(smem + a.off_sh);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int m_tile = blockIdx.x / kCluster;
  const int kb0 = (int)((long long)a.kb_total * rank / kCluster);
  const int kb1 = (int)((long long)a.kb_total * (rank + 1) / kCluster);
  const bool fused_ssq = a.want_resid && a.sumsq_ext == nullptr;
  const bool have_cur = m_tile * BLOCK_M < a.B;                 // uniform over the cluster
  const bool have_prev = m_tile * BLOCK_M < a.Bp;
  const int row_bytes = a.kf * 2;
  const uint32_t gal_tile_bytes = (uint32_t)kGalTile * (uint32_t)row_bytes;
  const int n_seq = 2 * a.g_tiles;
  const int b = m_tile * BLOCK_M + (int)rank * QB + lane;       // the crop this lane finishes (either batch)
  // per-column constants of the feature combination, fetched before the main loop (off the critical path)
  int my_exp[6] = {0, 0, 0, 0, 0, 0};
  double my_bias[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  if (warp < kStreamWarps) {
#pragma unroll
    for (int it = 0; it < 6; ++it) {
      const int c = warp + it * kStreamWarps;
      if (c < a.kq) {
        my_exp[it] = a.col_exp[c];
        if (c < a.k) my_bias[it] = a.bias[c];
      }
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
    for (int s = 0; s < kScoreBufs; ++s) {
      mbar_init(&sh->score_full[s], 1);
      mbar_init(&sh->score_empty[s], kScanWarps);
    }
    mbar_init(&sh->aimg_ready, kScanWarps);
    sh->failed = 0;
    sh->list_cnt = 0;
    sh->overflow = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();                                 // #1 "this CTA runs"
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory"); // the previous launch wrote the carried rows read below
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 16 : nullptr;
  if (probe && tid == 0) probe[0] = globaltimer();

  if (warp < kStreamWarps) {
    // =================================================================== stream half (current batch)
    if (warp == 0) {
      if (lane == 0 && have_cur) {
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
      cluster_wait();                               // #1
    } else if (warp == 1) {
      if (lane == 0 && have_cur) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t idesc = umma_idesc_i8(a.nc_pad);
        bool ok = true;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
          tc_fence_after();
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
        if (probe) probe[1] = globaltimer();
      }
      __syncwarp();
      cluster_wait();                               // #1
    } else {
      const int lane_group = warp & 3;
      const int row_in_tile = lane_group * 32 + lane;
      unsigned long long ssq = 0;
      bool ok = true;
      if (fused_ssq && have_cur) {
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
      if (have_cur && ok && kb1 > kb0) ok = mbar_wait(&sh->tmem_full_bar, 0, failed);
      ok = __all_sync(0xffffffffu, ok);
      tc_fence_after();
      __syncwarp();
      cluster_wait();                               // #1: the destination CTA's shared memory exists
      if (have_cur) {
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
      }
      tc_fence_before();
    }
    bar_stream();
    if (probe && tid == 0) probe[2] = globaltimer();
    cluster_arrive();                               // #2: my pushes are out
    cluster_wait();                                 //     all four partial slabs of my 32 crops have landed
    if (probe && tid == 0) probe[3] = globaltimer();
    if (have_cur) {
      const bool live = b < a.B;
      unsigned long long ssq_total = 0;
      if (warp == 0 && fused_ssq) {
#pragma unroll
        for (int q = 0; q < kCluster; ++q) ssq_total += sh->ssq_recv[q][lane];
      }
#pragma unroll
      for (int it = 0; it < 6; ++it) {
        const int c = warp + it * kStreamWarps;
        if (c >= KR) break;
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
        ps[c * QB + lane] = c < a.k ? v : 0.0;
        if (c >= a.k && c < a.kq) sh->xu[lane] = v;
      }
      if (KR < a.kq && warp == 0) {
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
      bar_stream();
      double n2 = 0.0;
      for (int c = 0; c < a.k; ++c) n2 = fma(ps[c * QB + lane], ps[c * QB + lane], n2);
      double pn = sqrt(n2);
      if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
      if (live) {
        // carried to the next launch: exact-scorer vector, norm, float16 filter row
        for (int c = warp; c < KR; c += kStreamWarps) {
          double v = ps[c * QB + lane];
          if (METRIC == EF_METRIC_COSINE_SK) v = v / pn;
          a.carry_pe[(size_t)b * KR + c] = v;
        }
        const float rinv = n2 > 0.0 ? rsqrtf((float)n2) : 0.f;
        const int KC = a.kf >> 3;
        for (int kc = warp; kc < KC; kc += kStreamWarps) {
          __align__(16) __half h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int kk = kc * 8 + i;
            const int seg = kk >= 3 * a.k ? 3 : (kk >= 2 * a.k ? 2 : (kk >= a.k ? 1 : 0));
            __half hi = __float2half_rn(0.f), lo = hi;
            if (seg < 3) split_half((float)ps[(kk - seg * a.k) * QB + lane] * rinv, hi, lo);
            h[i] = seg < 2 ? hi : lo;
          }
          *reinterpret_cast<uint4*>(a.carry_img + (size_t)b * a.kf + kc * 8) = *reinterpret_cast<const uint4*>(h);
        }
        if (warp == 0) {
          a.carry_pn[b] = pn;
          if (a.want_resid) {
            const double sq = fused_ssq ? (double)ssq_total : a.sumsq_ext[b];
            const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
            a.out_resid[b] = r > 0.0 ? r : 0.0;
          }
        }
      }
    }
    if (probe && tid == 0) probe[4] = globaltimer();
  } else {
    // =================================================================== match half (previous batch)
    cluster_wait();                                 // #1
    cluster_arrive();                               // #2 (nothing to publish); the wait is at the very end
    if (have_prev) {
      const int KC = a.kf >> 3;
      if (warp == 6) {
        if (lane == 0) {
          // gallery ring: first fill, then refills as the MMA warp releases slots
          const int first = min(a.ring, n_seq);
          int t = 0;
          for (int s = 0; s < first; ++s) {
            mbar_arrive_expect_tx(&sh->gal_full[s], gal_tile_bytes);
            bulk_load(gal + (size_t)s * gal_tile_bytes, reinterpret_cast<const uint8_t*>(a.gimg) + (size_t)t * gal_tile_bytes,
                      gal_tile_bytes, &sh->gal_full[s]);
            if (++t == a.g_tiles) t = 0;
          }
          int slot = 0;
          uint32_t ephase = 0;
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
      } else if (warp == 7) {
        if (lane == 0) {
          const uint32_t idesc = umma_idesc_f16(kGalTile);
          const int n_ks = a.kf >> 4;
          const int swb = row_bytes < 128 ? row_bytes : 128;
          const int pa_log2 = swb == 128 ? 2 : (swb == 64 ? 1 : 0);
          const uint64_t adesc0 = umma_desc_swz(smem_u32(aimg), 0, row_bytes, BLOCK_M);
          const uint64_t bdesc0 = umma_desc_swz(smem_u32(gal), 0, row_bytes, kGalTile);
          const uint64_t slot_step = (uint64_t)(gal_tile_bytes >> 4);
          bool ok = mbar_wait(&sh->aimg_ready, 0, failed);
          int slot = 0;
          uint32_t gphase = 0;
          uint64_t bslot = bdesc0;
          int buf = 0;
          uint32_t bphase = 1;                       // parity trick: a fresh barrier passes a wait on parity 1
          for (int s = 0; s < n_seq && ok; ++s) {
            if (!mbar_wait(&sh->gal_full[slot], gphase, failed)) break;
            if (!mbar_wait(&sh->score_empty[buf], bphase, failed)) break;
            tc_fence_after();
            const uint32_t d_addr = tmem_base + 128u + (uint32_t)buf * kGalTile;
#pragma unroll 1
            for (int ks = 0; ks < n_ks; ++ks) {
              const uint64_t koff = (uint64_t)((ks & ((1 << pa_log2) - 1)) << 1);
              umma_f16(d_addr, adesc0 + koff, bslot + koff, idesc, ks > 0 ? 1u : 0u);
            }
            umma_commit(&sh->gal_empty[slot]);
            umma_commit(&sh->score_full[buf]);
            bslot += slot_step;
            if (++slot == a.ring) { slot = 0; gphase ^= 1; bslot = bdesc0; }
            if (++buf == kScoreBufs) { buf = 0; bphase ^= 1; }
          }
        }
      } else {
        // ---- scanning warps 8..15
        const int sw = warp - (kWarps - kScanWarps), stid = tid - (kWarps - kScanWarps) * 32;
        const int q = warp & 3, hh = sw >> 2;
        const int col0 = (q * 2 + hh) * 16;           // 16 of the 128 score columns
        const bool livep = b < a.Bp;
        // carried rows -> shared memory: A operand (rows lane + 32 q' hold crop `lane`), exact-scorer vectors, norms
        for (int e = sw; e < 4 * KC; e += kScanWarps) {
          const int kc_log2 = 31 - __clz(KC);
          const int qq = e >> kc_log2, kc = e & (KC - 1), r = lane + 32 * qq;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (livep) v = __ldg(reinterpret_cast<const uint4*>(a.prev_img + (size_t)b * a.kf + kc * 8));
          *reinterpret_cast<uint4*>(aimg + swz_chunk_offset(r, kc, row_bytes, BLOCK_M)) = v;
        }
        for (int c = sw; c < KR; c += kScanWarps) pe[c * QB + lane] = livep ? __ldg(a.prev_pe + (size_t)b * KR + c) : 0.0;
        if (sw == 0) sh->pn[lane] = livep ? __ldg(a.prev_pn + b) : 1.0;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->aimg_ready);
        if (probe && stid == 0) probe[8] = globaltimer();
        double best = -CUDART_INF, best_score = 0.0;
        int best_i = INT_MAX, best_label = -1;
        auto consider = [&](double key, double score, int label, int j) {
          if (better<METRIC>(key, j, best, best_i)) { best = key; best_score = score; best_label = label; best_i = j; }
        };
        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F, thr = 0.f;
        bool ok = true;
        int t = -1, pass = 0, buf = -1;
        uint32_t fphase = 0;
        for (int s = 0; s < n_seq; ++s) {
          if (++buf == kScoreBufs) { buf = 0; fphase ^= 1; }
          if (++t == a.g_tiles) { t = 0; pass = 1; }
          if (pass == 1 && t == 0) {
            sh->fmax_s[sw][lane] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            bar_scan();
            float M = sh->fmax_s[0][lane];
#pragma unroll
            for (int w = 1; w < kScanWarps; ++w) M = fmaxf(M, sh->fmax_s[w][lane]);
            thr = M - 2.f * kFilterEps;
            if (probe && stid == 0) probe[9] = globaltimer();
          }
          ok = __all_sync(0xffffffffu, ok && mbar_wait(&sh->score_full[buf], fphase, failed));
          if (!ok) continue;
          tc_fence_after();
          uint32_t v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + 128u + (uint32_t)(buf * kGalTile + col0), v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh->score_empty[buf]);
          const int j0 = t * kGalTile + col0;
          const int valid = min(16, a.n - j0);
          if (valid <= 0) continue;
          if (pass == 0) {
            if (valid == 16) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                m0 = fmaxf(m0, __uint_as_float(v[i]));
                m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
                m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (i < valid) m0 = fmaxf(m0, __uint_as_float(v[i]));
            }
          } else {
            unsigned mask = 0u;
#pragma unroll
            for (int i = 0; i < 16; ++i) mask |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
            if (valid < 16) mask &= (1u << valid) - 1u;
            while (mask) {
              const int i = __ffs(mask) - 1;
              mask &= mask - 1u;
              const int slot = atomicAdd(&sh->list_cnt, 1);
              if (slot < kListCap) {
                sh->list_L[slot] = lane;
                sh->list_j[slot] = j0 + i;
              } else {
                sh->overflow = 1;
              }
            }
          }
        }
        __threadfence_block();
        bar_scan();
        if (probe && stid == 0) probe[10] = globaltimer();
        const bool overflow = *reinterpret_cast<volatile int*>(&sh->overflow) != 0;
        const int total = overflow ? 0 : *reinterpret_cast<volatile int*>(&sh->list_cnt);
        for (int e = stid; e < total; e += kScanWarps * 32) {
          const int L = sh->list_L[e];
          double key, score; int label;
          exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe, L, sh->list_j[e], sh->pn[L], key, score, label);
          sh->list_key[e] = key;
          sh->list_score[e] = score;
          sh->list_label[e] = label;
        }
        bar_scan();
        for (int e = sw; e < total; e += kScanWarps)
          if (sh->list_L[e] == lane) consider(sh->list_key[e], sh->list_score[e], sh->list_label[e], sh->list_j[e]);
        if (overflow) {
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
        bar_scan();
        if (sw == 0 && livep) {
          double bs = sh->red_s[0][lane], score = sh->red_d[0][lane];
          int bi = sh->red_i[0][lane], bl = sh->red_l[0][lane];
          for (int w = 1; w < kScanWarps; ++w)
            if (better<METRIC>(sh->red_s[w][lane], sh->red_i[w][lane], bs, bi)) {
              bs = sh->red_s[w][lane];
              score = sh->red_d[w][lane];
              bi = sh->red_i[w][lane];
              bl = sh->red_l[w][lane];
            }
          if (bi == INT_MAX) { bi = 0; bl = -1; }
          a.out_score[b] = score;
          a.out_index[b] = bi;
          if (a.out_label) a.out_label[b] = score >= a.threshold ? bl : -1;
        }
        if (probe && stid == 0) probe[11] = globaltimer();
      }
    }
    __syncwarp();
    cluster_wait();                                 // #2
  }

  // ======================================================================= teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
  if (probe && tid == 0) probe[6] = globaltimer();
}

template <int METRIC, int KR>
int launch_pipe(const CUtensorMap& mx, const CUtensorMap& mw, PipeArgs& a, int m_tiles, cudaStream_t stream) {
  const int stage_bytes = A_STAGE_BYTES + a.nc_pad * BLOCK_K;
  a.stages = 3;
  size_t off = (size_t)a.stages * stage_bytes;
  a.off_recv = (int)off; off += (size_t)a.nc_pad * 512;
  a.off_ps = (int)off;   off += sizeof(double) * KR * QB;
  a.off_pe = (int)off;   off += sizeof(double) * KR * QB;
  off = (size_t)ef::round_up((int64_t)off, 1024);
  const size_t aimg_bytes = (size_t)ef::round_up((int64_t)BLOCK_M * a.kf * 2, 1024);
  const size_t tile_bytes = (size_t)kGalTile * a.kf * 2;
  const size_t tail = (size_t)ef::round_up((int64_t)sizeof(PipeShared), 128) + 128;
  if ((size_t)kSmemLimit < off + aimg_bytes + tail + 2 * tile_bytes) return EF_ERR_UNSUPPORTED;
  int ring = (int)std::min<size_t>(kMaxRing, ((size_t)kSmemLimit - off - aimg_bytes - tail) / tile_bytes);
  ring = std::min(ring, 2 * a.g_tiles);
  if (ring < 2) return EF_ERR_UNSUPPORTED;
  a.ring = ring;
  a.off_aimg = (int)off; off += aimg_bytes;
  a.off_gal = (int)off;  off += (size_t)ring * tile_bytes;
  off = (size_t)ef::round_up((int64_t)off, 128);
  a.off_sh = (int)off;
  const size_t smem = off + sizeof(PipeShared);
  if (smem > (size_t)kSmemLimit) return EF_ERR_UNSUPPORTED;
  EF_ENSURE_SMEM((recognize_pipe_kernel<METRIC, KR>), smem);
  static unsigned long long* probe_buf = nullptr;
  const bool probing = getenv("EF_TC_PROBE") != nullptr;
  const int grid_n = m_tiles * kCluster;
  a.probe = nullptr;
  if (probing && grid_n <= 4096) {
    if (!probe_buf) EF_CUDA(cudaMalloc(&probe_buf, sizeof(unsigned long long) * 16 * 4096));
    EF_CUDA(cudaMemsetAsync(probe_buf, 0, sizeof(unsigned long long) * 16 * 4096, stream));
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
  EF_CUDA(cudaLaunchKernelEx(&cfg, recognize_pipe_kernel<METRIC, KR>, mx, mw, a));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (a.probe) {
    std::vector<unsigned long long> h((size_t)grid_n * 16);
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid_n; ++c) if (h[(size_t)c * 16] && h[(size_t)c * 16] < t0) t0 = h[(size_t)c * 16];
    const char* names[16] = {"start", "mma_issued", "pushed", "exchanged", "carried", "-", "end", "-",
                             "aimg", "max_known", "listed", "matched", "-", "-", "-", "-"};
    fprintf(stderr, "[ef_pipe_probe] grid %d B %d Bp %d ring %d; us since first CTA start (mean/max):", grid_n, a.B, a.Bp, a.ring);
    for (int i = 0; i < 16; ++i) {
      if (names[i][0] == '-') continue;
      double sum = 0, mx = 0;
      for (int c = 0; c < grid_n; ++c) {
        const double v = h[(size_t)c * 16 + i] ? (double)(h[(size_t)c * 16 + i] - t0) * 1e-3 : 0.0;
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
int dispatch_kr(const CUtensorMap& mx, const CUtensorMap& mw, PipeArgs& a, int kr, int m_tiles, cudaStream_t st) {
  switch (kr) {
    case 4: return launch_pipe<METRIC, 4>(mx, mw, a, m_tiles, st);
    case 8: return launch_pipe<METRIC, 8>(mx, mw, a, m_tiles, st);
    case 12: return launch_pipe<METRIC, 12>(mx, mw, a, m_tiles, st);
    case 16: return launch_pipe<METRIC, 16>(mx, mw, a, m_tiles, st);
    default: return launch_pipe<METRIC, 24>(mx, mw, a, m_tiles, st);
  }
}

}  // namespace

namespace ef {

bool pipe_supported(int k, int NC, int metric, int64_t n) {
  return metric != EF_METRIC_L2 && k <= 21 && round_up(NC, 16) <= 128 && n > 0 && n < (1ll << 31) - 512 &&
         filter_kf(k) <= 64;
}

// One pipelined launch: streams batch (X, B) and matches the previous batch (Bp rows carried in prev_*).  Either side
// may be empty (B = 0: flush; Bp = 0: first submit).  EF_ERR_UNSUPPORTED outside the kernel's coverage.
int recognize_pipe(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows, int k,
                   int kq, int S, const int32_t* col_exp, const double* bias, const double* sumsq_ext, bool want_resid,
                   double c0, double* out_proj, double* out_resid, double* carry_pe, double* carry_pn, void* carry_img,
                   int Bp, const double* prev_pe, const double* prev_pn, const void* prev_img, double* out_score,
                   int32_t* out_index, int32_t* out_label, double threshold, const double* gp_padded, int kpad,
                   const double* gnorm, const double* ginv, const void* gimg, int64_t n, const int32_t* labels,
                   int metric, int* status, cudaStream_t stream) {
  using namespace ef_tc;
  if (B <= 0 && Bp <= 0) return EF_OK;
  if (!pipe_supported(k, NC, metric, n) || kpad != fused_epilogue_kpad(k)) return EF_ERR_UNSUPPORTED;
  const int nc_pad = (int)round_up(NC, 16);
  if (nc_pad > wq_rows || (ldw & 15) || (reinterpret_cast<uintptr_t>(Wq) & 15)) return EF_ERR_UNSUPPORTED;
  if (B > 0 && ((ldx & 15) || (reinterpret_cast<uintptr_t>(X) & 15))) return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  PipeArgs a{};
  a.B = B > 0 ? B : 0; a.D = D; a.NC = NC; a.nc_pad = nc_pad; a.k = k; a.kq = kq; a.S = S;
  a.kb_total = (int)ceil_div(D, BLOCK_K);
  a.col_exp = col_exp; a.bias = bias; a.sumsq_ext = sumsq_ext; a.want_resid = want_resid ? 1 : 0; a.c0 = c0;
  a.out_proj = out_proj; a.out_resid = out_resid;
  a.carry_pe = carry_pe; a.carry_pn = carry_pn; a.carry_img = reinterpret_cast<__half*>(carry_img);
  a.Bp = Bp > 0 ? Bp : 0; a.prev_pe = prev_pe; a.prev_pn = prev_pn; a.prev_img = reinterpret_cast<const __half*>(prev_img);
  a.out_score = out_score; a.out_index = out_index; a.out_label = out_label; a.threshold = threshold;
  a.gp = gp_padded; a.gnorm = gnorm; a.ginv = ginv; a.labels = labels; a.n = (int)n;
  a.kf = filter_kf(k);
  a.g_tiles = (int)ceil_div(n, kGalTile);
  a.gimg = reinterpret_cast<const __half*>(gimg);
  a.status = status;
  CUtensorMap mx, mw;
  // an empty current batch still needs a valid crop map (never dereferenced): point it at the basis
  const void* xbase = B > 0 ? (const void*)X : (const void*)Wq;
  const uint64_t xrows = B > 0 ? (uint64_t)B : (uint64_t)wq_rows, xpitch = B > 0 ? (uint64_t)ldx : (uint64_t)ldw;
  if (!make_map(&mx, xbase, (uint64_t)D, xrows, xpitch, BLOCK_M)) return EF_ERR_UNSUPPORTED;
  if (!make_map(&mw, Wq, (uint64_t)ldw, (uint64_t)wq_rows, (uint64_t)ldw, (uint32_t)nc_pad)) return EF_ERR_UNSUPPORTED;
  const int m_tiles = (int)ceil_div(std::max(a.B, a.Bp), BLOCK_M);
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(mx, mw, a, kpad, m_tiles, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(mx, mw, a, kpad, m_tiles, stream);
    default: return EF_ERR_UNSUPPORTED;
  }
}

}  // namespace ef
