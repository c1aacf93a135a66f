// K2, single-kernel form: crops in, identities out.
//
// One thread-block CLUSTER of 4 CTAs owns a tile of 128 crops.  Each CTA streams one quarter of the pixel (K)
// dimension through TMA -> shared memory -> tcgen05.mma kind::i8 -> TMEM exactly like project_tc_kernel, so all SMs
// pull HBM although a 4096-crop batch has only 32 crop tiles.  The four partial int32 tiles are then exchanged through
// DISTRIBUTED SHARED MEMORY: CTA r of the cluster sums, for its 32 crops, the partials of all four CTAs (exact integer
// adds), combines the digit planes into float64 features, and runs the nearest-gallery search + threshold + label for
// those 32 crops against a gallery tile that was prefetched into shared memory (cp.async) while the crops streamed.
// No global accumulators, no atomics, no memset, no second launch: HBM traffic is the crops (once) + 24 B per crop out.
//
// Covers k <= 32 and S*(k+1) <= 256 digit-plane columns (every shipped Gen-1 model and the k<=30 sklearn models at
// S = 8); other shapes use project_tc_kernel + the separate epilogue kernels.
//
// Replaces project_face_to_eigenspace + recognize_face (useless/scan.py:80-132) and scaler.transform + pca.transform +
// recognize_face_with_model (scan-template-v4.py:265-287) for a whole batch.
#include <climits>
#include <cstdlib>
#include <vector>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kCluster = 4;                 // CTAs per crop tile = K splits
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;       // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 TMEM drain + sum of squares,
                                            // warps 6..15 gallery prefetch; all 16 warps combine + match
constexpr int QB = BLOCK_M / kCluster;      // crops finished by each CTA (one per lane)
static_assert(QB == 32, "one crop per lane");

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
  unsigned long long* probe;   // debugging aid (EF_TC_PROBE): [grid][8] globaltimer stamps
};

struct ClusterShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar;
  uint32_t tmem_base;
  int failed;
  double pn[QB];
  double xu[QB];
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
  // after the main loop the stage area is reused for the partial tile:
  int32_t* part = reinterpret_cast<int32_t*>(smem);             // [nc_pad][128] int32, plane-major
  unsigned long long* ssq_s = reinterpret_cast<unsigned long long*>(smem + (size_t)a.nc_pad * BLOCK_M * 4);   // [128]
  uint8_t* after = smem + (size_t)a.stages * stage_bytes;
  double* ps = reinterpret_cast<double*>(after);                 // [KR][QB]
  double* gs = ps + KR * QB;                                     // [tile_rows][KR]
  double* gw = gs + (size_t)a.tile_rows * KR;                    // [tile_rows]
  ClusterShared* sh = reinterpret_cast<ClusterShared*>(gw + a.tile_rows);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();                       // K quarter of this CTA
  const int m_tile = blockIdx.x / kCluster;
  const int kb0 = (int)((long long)a.kb_total * rank / kCluster);
  const int kb1 = (int)((long long)a.kb_total * (rank + 1) / kCluster);
  const bool fused_ssq = a.want_resid && a.sumsq_ext == nullptr;

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], fused_ssq ? 5 : 1);
    }
    mbar_init(&sh->tmem_full_bar, 1);
    sh->failed = 0;
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
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 8 : nullptr;
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
    // drain the accumulator into shared memory (the pipeline stages are free once the last MMA has completed and
    // all four warps have finished reading the last crop tile)
    if (ok && kb1 > kb0) ok = mbar_wait(&sh->tmem_full_bar, 0, failed);
    tc_fence_after();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int c0 = 0; c0 < a.nc_pad; c0 += 16) {
      uint32_t v[16];
      if (ok && kb1 > kb0) {
        tmem_ld16(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)c0, v);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) part[(size_t)(c0 + j) * BLOCK_M + row_in_tile] = (int32_t)v[j];
    }
    ssq_s[row_in_tile] = ssq;
    tc_fence_before();
  } else {
    // warps 6..15: pull the first gallery tile into shared memory while the crops stream
    load_gallery_tile(0, tid - 6 * 32, kThreads - 6 * 32);
  }

  // ======================================================================= exchange partial tiles through DSMEM
  __syncthreads();
  if (probe && tid == 0) probe[3] = globaltimer();
  cluster_sync_all();                               // every CTA's partial tile is in its shared memory
  if (probe && tid == 0) probe[4] = globaltimer();
  const int b = m_tile * BLOCK_M + (int)rank * QB + lane;       // the crop this lane finishes
  const bool live = b < a.B;
  for (int c = warp; c < KR; c += kWarps) ps[c * QB + lane] = 0.0;
  if (warp == 0) sh->xu[lane] = 0.0;
  __syncthreads();
  {
    const uint32_t part_local = smem_u32(part);
    uint32_t part_remote[kCluster];
#pragma unroll
    for (int q = 0; q < kCluster; ++q) part_remote[q] = map_to_cta(part_local, (uint32_t)q);
    const int my_row = (int)rank * QB + lane;
    for (int c = warp; c < a.kq; c += kWarps) {
      double v = 0.0;
      for (int s = a.S - 1; s >= 0; --s) {
        const uint32_t off = (uint32_t)(((s * a.kq + c) * BLOCK_M + my_row) * 4);
        int sum = 0;
#pragma unroll
        for (int q = 0; q < kCluster; ++q) sum += ld_cluster_s32(part_remote[q] + off);   // exact: |full-K sum| < 2^31
        v += (double)sum * __longlong_as_double((long long)(1023 - (7 * s + 6)) << 52);
      }
      v = ldexp(v, a.col_exp[c]);
      if (c < a.k) {
        v -= a.bias[c];
        ps[c * QB + lane] = v;
        if (a.out_proj && live) a.out_proj[(size_t)b * a.k + c] = v;
      } else {
        sh->xu[lane] = v;
      }
    }
  }
  unsigned long long ssq_total = 0;
  if (warp == 0 && fused_ssq) {
    const uint32_t ssq_local = smem_u32(ssq_s + (int)rank * QB + lane);
#pragma unroll
    for (int q = 0; q < kCluster; ++q) ssq_total += ld_cluster_u64(map_to_cta(ssq_local, (uint32_t)q));
  }
  __syncthreads();
  cluster_sync_all();                               // nobody reads remote shared memory after this point
  if (probe && tid == 0) probe[5] = globaltimer();

  if (warp == 0) {
    double n2 = 0.0;
    for (int c = 0; c < a.k; ++c) n2 = fma(ps[c * QB + lane], ps[c * QB + lane], n2);
    double pn = sqrt(n2);
    if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
    sh->pn[lane] = pn;
    if (a.want_resid && live) {
      const double sq = fused_ssq ? (double)ssq_total : a.sumsq_ext[b];
      const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
      a.out_resid[b] = r > 0.0 ? r : 0.0;
    }
  }
  __syncthreads();
  double p[KR];
#pragma unroll
  for (int c = 0; c < KR; ++c) {
    double v = ps[c * QB + lane];
    if (METRIC == EF_METRIC_COSINE_SK) v = v / sh->pn[lane];
    p[c] = v;
  }

  // ======================================================================= nearest gallery row (float64)
  double best = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF, best_dot = 0.0;
  int best_i = INT_MAX;
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
      const double pn = sh->pn[lane], gn = a.gnorm[bi];
      score = (pn == 0.0 || gn == 0.0) ? 0.0 : bd / (pn * gn);       // useless/scan.py:70-77
    }
    a.out_score[b] = score;
    a.out_index[b] = bi;
    if (a.out_label) {
      const bool pass = METRIC == EF_METRIC_L2 ? score <= a.threshold : score >= a.threshold;
      a.out_label[b] = pass ? (a.labels ? a.labels[bi] : bi) : -1;
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

template <int METRIC, int KR>
int launch_cluster(const CUtensorMap& mx, const CUtensorMap& mw, ClusterArgs& a, int m_tiles, cudaStream_t stream) {
  const int stage_bytes = A_STAGE_BYTES + a.nc_pad * BLOCK_K;
  const size_t fixed = 1024 + sizeof(ClusterShared) + sizeof(double) * KR * QB + 64;
  // pipeline depth: 3 stages are enough (the main loop is throughput bound, see tools/tc_probe.py); the partial tile
  // (nc_pad x 128 int32 + 128 x u64) must fit in the stage area
  int stages = 3;
  while ((size_t)stages * stage_bytes < (size_t)a.nc_pad * BLOCK_M * 4 + BLOCK_M * 8) ++stages;
  if (stages > kMaxStages) return EF_ERR_UNSUPPORTED;
  const size_t left = (size_t)kSmemLimit - fixed - (size_t)stages * stage_bytes;
  int rows = (int)(left / (sizeof(double) * (KR + 1)));
  rows &= ~3;
  const int n4 = (a.n + 3) & ~3;
  if (rows > n4) rows = n4;
  if (rows < 64 && rows < n4) return EF_ERR_UNSUPPORTED;
  a.tile_rows = rows;
  a.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes + sizeof(double) * (size_t)rows * (KR + 1);
  static size_t attr = 0;
  if (smem > attr) {
    EF_CUDA(cudaFuncSetAttribute(recognize_cluster_kernel<METRIC, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
    attr = smem;
  }
  static unsigned long long* probe_buf = nullptr;
  const bool probing = getenv("EF_TC_PROBE") != nullptr;
  const int grid_n = m_tiles * kCluster;
  a.probe = nullptr;
  if (probing && grid_n <= 4096) {
    if (!probe_buf) EF_CUDA(cudaMalloc(&probe_buf, sizeof(unsigned long long) * 8 * 4096));
    EF_CUDA(cudaMemsetAsync(probe_buf, 0, sizeof(unsigned long long) * 8 * 4096, stream));
    a.probe = probe_buf;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(m_tiles * kCluster));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCluster;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  EF_CUDA(cudaLaunchKernelEx(&cfg, recognize_cluster_kernel<METRIC, KR>, mx, mw, a));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (a.probe) {
    std::vector<unsigned long long> h((size_t)grid_n * 8);
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid_n; ++c) if (h[(size_t)c * 8] && h[(size_t)c * 8] < t0) t0 = h[(size_t)c * 8];
    const char* names[7] = {"start", "first_full", "mma_issued", "block_done", "sync1", "sync2", "end"};
    fprintf(stderr, "[ef_cluster_probe] grid %d stages %d tile_rows %d; us since first CTA start (mean/max):", grid_n, a.stages, a.tile_rows);
    for (int i = 0; i < 7; ++i) {
      double sum = 0, mx = 0;
      for (int c = 0; c < grid_n; ++c) {
        const double v = h[(size_t)c * 8 + i] ? (double)(h[(size_t)c * 8 + i] - t0) * 1e-3 : 0.0;
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

// EF_ERR_UNSUPPORTED when the shape / alignment is outside what the single-kernel form covers (caller falls back).
int recognize_cluster(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows,
                      int k, int kq, int S, const int32_t* col_exp, const double* bias, const double* sumsq_ext,
                      bool want_resid, double c0, const double* gp_padded, int kpad, const double* gnorm,
                      const double* ginv, int64_t n, const int32_t* labels, int metric, double threshold,
                      double* out_proj, double* out_score, int32_t* out_index, int32_t* out_label, double* out_resid,
                      int* status, cudaStream_t stream) {
  using namespace ef_tc;
  if (B <= 0) return EF_OK;
  if (k > 32 || kpad != fused_epilogue_kpad(k) || n <= 0 || n >= (1ll << 31) - 8) return EF_ERR_UNSUPPORTED;
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
