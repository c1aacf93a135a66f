// Large-gallery nearest-neighbour search on tensor cores (BASELINE config 3: 1 M identities x k = 128, sharded).
//
// The float64 scan of ef_match.cu costs 2 k n flops per query on the FP64 pipe (1.05 PFLOP for 4096 queries against a
// million rows).  Here the scan becomes a FILTER on the tensor cores and only the survivors are scored in float64:
//   * queries and gallery rows are normalised and split into float16 hi + lo; s~ = q_hi.g_hi + q_hi.g_lo + q_lo.g_hi
//     is a tcgen05.mma kind::f16 GEMM (M = 128 queries, N = 256 gallery rows, K = 3k padded to slabs of 64) with
//     float32 accumulation in TMEM; |s~ - cos| <= kEps (3k <= 384 products: 384 * 2^-22 accumulation + 2e-6 split);
//   * ONE pass over the gallery image: every scanning thread (= one query) keeps the running maximum of the scores
//     it has seen, publishes it to a per-query word in global memory (atomicMax) and picks up what the CTAs working
//     on other gallery chunks published at every tile boundary; a row goes to the candidate list when
//     s~ >= (best maximum known so far) - 2 kEps.  The known maximum only grows towards the final maximum M, so the list
//     is a superset of {s~ >= M - 2 kEps} (which contains the exact arg-max and all exact ties); entries below the
//     FINAL threshold are dropped by their stored s~ before the float64 work.  (Round 1 read the 768 MB image twice:
//     pass 0 for M, pass 1 for the rows.)  The survivors are scored in float64 with EXACTLY the arithmetic of
//     match_kernel (same norm, same sequential fma order, same tie rule), so score and index are bit identical to the
//     float64 scan.  A candidate-list overflow (degenerate galleries) reports EF_ERR_UNSUPPORTED at the next
//     synchronisation point of the caller through the flag word; the Python wrapper then runs the float64 scan.
// Euclidean distance (the reference's first-generation matcher variant) rides the same GEMM with ONE extra component:
//   |p - g|^2 / (2 |p| G) = |p| / (2 G)  -  ( p^ . g/G  -  r c ),   r = G / (2 |p|),  c = |g|^2 / G^2,  G = max |g|
// so arg-min of the distance = arg-max of key = [p^, -r] . [g/G, c]; both augmented rows are scaled to components
// <= 1 (the query side by t = 1 / max(1, r) > 0, which does not move the arg-max) and split hi/lo like the cosine
// operands.  The survivors are re-scored as sum (p - g)^2 in the float64 order of match_kernel.
// Blackwell mapping: the float16 gallery image streams as 32 KB (256 rows x one 128-byte K slab, SWIZZLE_128B) blocks
// through a cp.async.bulk ring; the query tile is resident in shared memory; two 256-column TMEM accumulators alternate
// between the MMA warp and four scanning warps (TMEM lane = query, tcgen05.ld x32 per 32 gallery rows).
//
// Replaces cosine_similarity + np.argmax of scan-template-v4.py:274-276 (and the loop of useless/scan.py:121-127) for a
// gallery far larger than the reference ever held.
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kThreads = 192;               // warp 0 bulk loads, warp 1 MMA + TMEM, warps 2..5 scan
constexpr int BN = 256;                     // gallery rows per tile (UMMA N)
constexpr int kSlab = 64;                   // halfs of K per slab = one 128-byte swizzle row
constexpr int kSlabBytesA = BLOCK_M * 128;  // 16 KB
constexpr int kSlabBytesB = BN * 128;       // 32 KB
constexpr int kMaxSlabs = 7;                // 3k <= 384 (cosine), 3(k + 1) <= 387 (L2: one extra component)
constexpr float kEps = 2e-4f;               // cosine: unit rows
constexpr float kEpsL2 = 4e-4f;             // L2: augmented rows of norm <= sqrt(2) on both sides
constexpr size_t kTrailer = 256;            // after the image: [0] double G = largest gallery norm (L2 only)

struct MatchTcArgs {
  const double* P;
  long long ldp;
  int B, k, n_slabs, stages, metric, pass;
  const __half* img;
  long long n;
  int g_tiles, tiles_per_chunk, b_pad;
  float* cmax;                 // [chunks][b_pad]           (pass 0 out)
  const float* thr;            // [b_pad]                   (pass 1 in)
  unsigned int* gmax;          // [b_pad] ordered-integer image of the best approximate score known per query (pass 2)
  int* cand_q;
  long long* cand_j;
  float* cand_f;               // approximate score of the candidate (pass 2: pruned against the final threshold)
  unsigned int* counter;       // [0] candidates, [1] overflow flag
  unsigned int cap;
  int* status;
  const double* gscale;        // L2: largest gallery norm G (image trailer)
  float eps;                   // |approximate key - exact key| bound of this metric
};

// order-preserving map float <-> unsigned (atomicMax on scores); 0 is below every finite score
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return u == 0u ? -CUDART_INF_F : __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct MatchTcShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar[2];
  unsigned long long tmem_empty_bar[2];
  uint32_t tmem_base;
  int failed;
};

__global__ void __launch_bounds__(kThreads, 1)
match_tc_kernel(const MatchTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  uint8_t* sA = smem;                                            // [n_slabs][128 rows][128 B]
  uint8_t* sB = smem + (size_t)a.n_slabs * kSlabBytesA;          // [stages][256 rows][128 B]
  MatchTcShared* sh = reinterpret_cast<MatchTcShared*>(sB + (size_t)a.stages * kSlabBytesB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, chunk = blockIdx.y;
  const int gt0 = chunk * a.tiles_per_chunk;
  const int gt1 = min(a.g_tiles, gt0 + a.tiles_per_chunk);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->tmem_full_bar[s], 1);
      mbar_init(&sh->tmem_empty_bar[s], 4);
    }
    sh->failed = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // ---- query tile -> float16 [hi | hi | lo] operand, rows normalised (float32 arithmetic: the filter is approximate)
  {
    const int KF = a.n_slabs * kSlab;
    const bool l2 = a.metric == EF_METRIC_L2;
    const int ka = l2 ? a.k + 1 : a.k;               // components per segment
    const double G = l2 ? *a.gscale : 0.0;
    for (int r = warp; r < BLOCK_M; r += kThreads / 32) {
      const int q = qt * BLOCK_M + r;
      double s2 = 0.0;
      if (q < a.B)
        for (int c = lane; c < a.k; c += 32) {
          const double v = a.P[(long long)q * a.ldp + c];
          s2 += v * v;
        }
      s2 = ef::warp_sum(s2);
      float rinv = s2 > 0.0 ? rsqrtf((float)s2) : 0.f;
      float last = 0.f;                              // L2: the extra query component -t r
      if (l2) {
        const double dinv = s2 > 0.0 ? 1.0 / sqrt(s2) : 0.0;
        const double rq = s2 > 0.0 ? 0.5 * G * dinv : 1.0;      // a zero query: key = -c (distance = |g|^2)
        const double t = rq > 1.0 ? 1.0 / rq : 1.0;
        rinv = (float)(dinv * t);
        last = -(float)(rq * t);
      }
      for (int kk = lane; kk < KF; kk += 32) {
        const int seg = kk >= 3 * ka ? 3 : (kk >= 2 * ka ? 2 : (kk >= ka ? 1 : 0));
        __half h = __float2half_rn(0.f);
        if (seg < 3 && q < a.B) {
          const int c = kk - seg * ka;
          const float v = c < a.k ? (float)a.P[(long long)q * a.ldp + c] * rinv : last;
          const __half hi = __float2half_rn(v);
          h = seg < 2 ? hi : __float2half_rn(v - __half2float(hi));
        }
        const int slab = kk >> 6, kin = kk & 63;
        *reinterpret_cast<__half*>(sA + (size_t)slab * kSlabBytesA + swz_chunk_offset(r, kin >> 3, 128, BLOCK_M) +
                                   (kin & 7) * 2) = h;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int gt = gt0; gt < gt1 && ok; ++gt)
        for (int slab = 0; slab < a.n_slabs; ++slab) {
          if (!mbar_wait(&sh->empty_bar[stage], phase ^ 1, failed)) { ok = false; break; }
          mbar_arrive_expect_tx(&sh->full_bar[stage], (uint32_t)kSlabBytesB);
          bulk_load(sB + (size_t)stage * kSlabBytesB,
                    reinterpret_cast<const uint8_t*>(a.img) + ((size_t)gt * a.n_slabs + slab) * kSlabBytesB,
                    (uint32_t)kSlabBytesB, &sh->full_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t aphase = 1;                          // a fresh barrier passes a wait on parity 1
      bool ok = true;
      for (int gt = gt0; gt < gt1 && ok; ++gt) {
        if (!mbar_wait(&sh->tmem_empty_bar[acc], aphase, failed)) break;
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)acc * BN;
        for (int slab = 0; slab < a.n_slabs; ++slab) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)slab * kSlabBytesA);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * kSlabBytesB);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(d_addr, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc,
                     (slab > 0 || ks > 0) ? 1u : 0u);
          umma_commit(&sh->empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        umma_commit(&sh->tmem_full_bar[acc]);
        if (++acc == 2) { acc = 0; aphase ^= 1; }
      }
    }
  } else {
    const int lane_group = warp & 3;
    const int q = qt * BLOCK_M + lane_group * 32 + lane;         // this thread's query
    const bool live = q < a.B;
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
    float thr = (a.pass == 1 && live) ? a.thr[q] : CUDART_INF_F;
    float runmax = -CUDART_INF_F, published = -CUDART_INF_F;     // pass 2: this thread's running maximum
    int acc = 0;
    uint32_t fphase = 0;
    bool ok = true;
    for (int gt = gt0; gt < gt1; ++gt) {
      ok = __all_sync(0xffffffffu, ok && mbar_wait(&sh->tmem_full_bar[acc], fphase, failed));
      if (!ok) break;
      tc_fence_after();
      const long long jbase = (long long)gt * BN;
      if (a.pass == 2 && live) {
        // what the CTAs of the other gallery chunks have found so far (a plain L2 read: any earlier value is valid)
        const float known = ord2f(*reinterpret_cast<volatile unsigned int*>(a.gmax + q));
        thr = fmaxf(runmax, known) - 2.f * a.eps;
      }
      for (int c0 = 0; c0 < BN; c0 += 32) {
        const long long j0 = jbase + c0;
        if (j0 >= a.n) break;                      // warp uniform
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)(acc * BN + c0), v);
        const int valid = (int)min((long long)32, a.n - j0);
        if (a.pass == 0) {
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
        } else if (a.pass == 2) {
          // maximum of the group first (four independent chains), threshold tightened with it, and only a group that
          // reaches the band is looked at entry by entry: in steady state that is ~ln(rows) groups per query
          float g0 = -CUDART_INF_F, g1 = -CUDART_INF_F, g2 = -CUDART_INF_F, g3 = -CUDART_INF_F;
          if (valid == 32) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              g0 = fmaxf(g0, __uint_as_float(v[i]));
              g1 = fmaxf(g1, __uint_as_float(v[i + 1]));
              g2 = fmaxf(g2, __uint_as_float(v[i + 2]));
              g3 = fmaxf(g3, __uint_as_float(v[i + 3]));
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < valid) g0 = fmaxf(g0, __uint_as_float(v[i]));
          }
          const float vm = fmaxf(fmaxf(g0, g1), fmaxf(g2, g3));
          runmax = fmaxf(runmax, vm);
          thr = fmaxf(thr, runmax - 2.f * a.eps);
          const bool hit = live && vm >= thr;
          if (__any_sync(0xffffffffu, hit)) {
            unsigned mask = 0u;
            if (hit) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
              if (valid < 32) mask &= (1u << valid) - 1u;
            }
            // one reservation per warp and column group: the lanes' counts are prefix-summed with shuffles
            const int cnt = __popc(mask);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned base = 0u;
            if (lane == 0) base = atomicAdd(a.counter, (unsigned)total);
            base = __shfl_sync(0xffffffffu, base, 0);
            unsigned slot = base + (unsigned)(incl - cnt);
#pragma unroll
            for (int i = 0; i < 32; ++i) {             // unrolled: v[] stays in registers (no dynamic index)
              if (mask & (1u << i)) {
                if (slot < a.cap) {
                  a.cand_q[slot] = q;
                  a.cand_j[slot] = j0 + i;
                  a.cand_f[slot] = __uint_as_float(v[i]);
                } else {
                  a.counter[1] = 1u;
                }
                ++slot;
              }
            }
          }
        } else {
          unsigned mask = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) mask |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
          if (valid < 32) mask &= (1u << valid) - 1u;
          while (mask) {
            const int i = __ffs(mask) - 1;
            mask &= mask - 1u;
            const unsigned slot = atomicAdd(a.counter, 1u);
            if (slot < a.cap) {
              a.cand_q[slot] = q;
              a.cand_j[slot] = j0 + i;
            } else {
              a.counter[1] = 1u;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty_bar[acc]);
      if (++acc == 2) { acc = 0; fphase ^= 1; }
      if (a.pass == 2 && live && runmax > published) {
        atomicMax(a.gmax + q, f2ord(runmax));
        published = runmax;
      }
    }
    if (a.pass == 0 && q < a.b_pad)
      a.cmax[(size_t)chunk * a.b_pad + q] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
}

// largest gallery norm (norms are >= 0: their bit patterns order like the values)
__global__ void match_tc_maxnorm_kernel(const double* __restrict__ norms, long long n, unsigned long long* out) {
  double m = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
    m = fmax(m, norms[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// gallery rows -> float16 [g_hi | g_lo | g_hi] image: [tile of 256 rows][slab][256 rows x 128 B, SWIZZLE_128B]
__global__ void match_tc_image_kernel(const double* __restrict__ gp, long long ldg, const double* __restrict__ norms,
                                      long long n, int k, int n_slabs, int metric, uint8_t* __restrict__ img,
                                      const double* __restrict__ gscale) {
  const int chunks_per_row = n_slabs * 8;          // 16-byte chunks of 8 halfs
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n * chunks_per_row) return;
  const long long j = e / chunks_per_row;
  const int ck = (int)(e - j * chunks_per_row);
  const bool l2 = metric == EF_METRIC_L2;
  const int ka = l2 ? k + 1 : k;
  double scale = 1.0, extra = 0.0;
  if (metric == EF_METRIC_COSINE_G1) {
    const double nr = norms[j];
    scale = nr == 0.0 ? 0.0 : 1.0 / nr;
  } else if (l2) {
    const double G = *gscale;
    scale = G == 0.0 ? 0.0 : 1.0 / G;
    const double u = norms[j] * scale;
    extra = u * u;                                 // c = |g|^2 / G^2 in [0, 1]
  }
  __align__(16) __half h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kk = ck * 8 + i;
    const int seg = kk >= 3 * ka ? 3 : (kk >= 2 * ka ? 2 : (kk >= ka ? 1 : 0));
    __half hi = __float2half_rn(0.f), lo = hi;
    if (seg < 3) {
      const int c = kk - seg * ka;
      const double v = c < k ? gp[j * ldg + c] * scale : extra;
      hi = __double2half(v);
      lo = __double2half(v - (double)__half2float(hi));
    }
    h[i] = seg == 1 ? lo : hi;
  }
  const long long tile = j / BN;
  const int rr = (int)(j - tile * BN);
  const int slab = ck >> 3, cin = ck & 7;
  uint8_t* dst = img + ((size_t)tile * n_slabs + slab) * kSlabBytesB + swz_chunk_offset(rr, cin, 128, BN);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
}

// thr[q] = max over chunks of cmax - 2 eps
__global__ void match_tc_thr_kernel(const float* __restrict__ cmax, int chunks, int b_pad, int B, float* __restrict__ thr,
                                    float eps) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b_pad) return;
  float m = -CUDART_INF_F;
  for (int c = 0; c < chunks; ++c) m = fmaxf(m, cmax[(size_t)c * b_pad + q]);
  thr[q] = q < B ? m - 2.f * eps : CUDART_INF_F;
}

// single pass: thr[q] = final maximum (published by every scanning thread) - 2 eps
__global__ void match_tc_thr_gmax_kernel(const unsigned int* __restrict__ gmax, int b_pad, int B, float* __restrict__ thr,
                                         float eps) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b_pad) return;
  thr[q] = q < B ? ord2f(gmax[q]) - 2.f * eps : CUDART_INF_F;
}

// query norms with the arithmetic of match_kernel (lane-strided partial sums, xor-shuffle tree)
__global__ void match_tc_norm_kernel(const double* __restrict__ P, long long ldp, int B, int k, int metric,
                                     double* __restrict__ pn, unsigned long long* __restrict__ best_key,
                                     long long* __restrict__ best_idx) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (q >= B) return;
  double s = 0.0;
  for (int c = lane; c < k; c += 32) {
    const double v = P[(long long)q * ldp + c];
    s += v * v;
  }
  s = ef::warp_sum(s);
  if (lane == 0) {
    double nrm = sqrt(s);
    if (metric == EF_METRIC_COSINE_SK && nrm == 0.0) nrm = 1.0;
    pn[q] = nrm;
    best_key[q] = 0ull;                              // below every encoded score
    best_idx[q] = LLONG_MAX;
  }
}

__device__ __forceinline__ unsigned long long order_key(double s) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(s);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);           // monotone in s; 0 is below every real score
}

// exact float64 score of every candidate, arithmetic of match_kernel: sequential fma over the components
__global__ void match_tc_rescore_kernel(const double* __restrict__ P, long long ldp, int k, const double* __restrict__ G,
                                        long long ldg, const double* __restrict__ gnorm, const double* __restrict__ pn,
                                        int* cand_q, const long long* __restrict__ cand_j,
                                        const unsigned int* counter, unsigned int cap, int metric,
                                        double* __restrict__ cand_s, unsigned long long* __restrict__ best_key,
                                        const float* __restrict__ cand_f, const float* __restrict__ thr,
                                        unsigned int* kept) {
  const unsigned int total = min(counter[0], cap);
  for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int q = cand_q[e];
    if (cand_f && cand_f[e] < thr[q]) {              // entered against an earlier, lower maximum: outside the final band
      cand_q[e] = -1;
      continue;
    }
    if (kept) atomicAdd(kept, 1u);
    const long long j = cand_j[e];
    const double* p = P + (long long)q * ldp;
    const double* g = G + j * ldg;
    const double nq = pn[q];
    double acc = 0.0;
    if (metric == EF_METRIC_L2) {
      for (int c = 0; c < k; ++c) {
        const double d = p[c] - g[c];
        acc = fma(d, d, acc);
      }
      acc = -acc;                                    // the lists and keys hold "higher is better"
    } else if (metric == EF_METRIC_COSINE_SK) {
      for (int c = 0; c < k; ++c) acc = fma(p[c] / nq, g[c], acc);
    } else {
      for (int c = 0; c < k; ++c) acc = fma(p[c], g[c], acc);
      const double gn = gnorm[j];
      acc = (nq == 0.0 || gn == 0.0) ? 0.0 : acc / (nq * gn);
    }
    cand_s[e] = acc;
    atomicMax(best_key + q, order_key(acc));
  }
}

__global__ void match_tc_pick_kernel(const int* __restrict__ cand_q, const long long* __restrict__ cand_j,
                                     const double* __restrict__ cand_s, const unsigned int* __restrict__ counter,
                                     unsigned int cap, const unsigned long long* __restrict__ best_key,
                                     long long* __restrict__ best_idx) {
  const unsigned int total = min(counter[0], cap);
  for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int q = cand_q[e];
    if (q < 0) continue;                             // pruned
    if (order_key(cand_s[e]) == best_key[q]) atomicMin(reinterpret_cast<unsigned long long*>(best_idx + q),
                                                       (unsigned long long)cand_j[e]);
  }
}

__global__ void match_tc_final_kernel(const unsigned long long* __restrict__ best_key, const long long* __restrict__ best_idx,
                                      int B, long long index_base, int metric, double* __restrict__ out_score,
                                      long long* __restrict__ out_index) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  const unsigned long long kbits = best_key[q];
  const long long bits = (kbits >> 63) ? (long long)(kbits & 0x7fffffffffffffffull) : (long long)~kbits;
  const double s = kbits == 0ull ? -CUDART_INF : __longlong_as_double(bits);
  out_score[q] = metric == EF_METRIC_L2 ? -s : s;
  out_index[q] = best_idx[q] == LLONG_MAX ? -1 : best_idx[q] + index_base;
}

int n_slabs_for(int k, int metric) { return (int)ef::ceil_div(3 * (int64_t)(metric == EF_METRIC_L2 ? k + 1 : k), kSlab); }
size_t image_body_bytes(int64_t n, int k, int metric) {
  return (size_t)ef::ceil_div(n, BN) * n_slabs_for(k, metric) * kSlabBytesB;
}

struct Layout {
  size_t cmax, thr, gmax, pn, best_key, best_idx, counter, cand_q, cand_j, cand_s, cand_f, status, total;
  int chunks, tiles_per_chunk, b_pad;
  unsigned int cap;
};

Layout work_layout(int B, int64_t n, int k) {
  Layout L{};
  const int q_tiles = (int)ef::ceil_div(B, BLOCK_M);
  const int g_tiles = (int)ef::ceil_div(n, BN);
  // gallery chunks per query tile: whole waves of CTAs (one CTA per SM).  32 query tiles x 10 chunks = 320 CTAs were
  // 2.16 waves on 148 SMs -- the third wave ran 16 % full; the wave count in 2..8 that fills its last wave best wins
  // (32 x 37 = 1184 = 8 x 148)
  const int64_t sms = ef::sm_count();
  int chunks = 1;
  double best_fill = 0.0;
  for (int waves = 2; waves <= 8; ++waves) {
    const int64_t c = std::max<int64_t>(1, (waves * sms) / q_tiles);
    const double fill = (double)(c * q_tiles) / (double)(ef::ceil_div(c * q_tiles, sms) * sms);
    if (fill > best_fill + 1e-9) { best_fill = fill; chunks = (int)c; }
  }
  if (chunks > g_tiles) chunks = g_tiles;
  if (chunks < 1) chunks = 1;
  L.tiles_per_chunk = (int)ef::ceil_div(g_tiles, chunks);
  L.chunks = (int)ef::ceil_div(g_tiles, L.tiles_per_chunk);
  L.b_pad = q_tiles * BLOCK_M;
  L.cap = (unsigned int)std::max<int64_t>(1 << 20, std::min<int64_t>((int64_t)B * 512, 1 << 26));
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  L.status = take(256);
  L.counter = take(256);
  L.cmax = take(sizeof(float) * (size_t)L.chunks * L.b_pad);
  L.thr = take(sizeof(float) * L.b_pad);
  L.gmax = take(sizeof(unsigned int) * L.b_pad);
  L.pn = take(sizeof(double) * B);
  L.best_key = take(sizeof(unsigned long long) * B);
  L.best_idx = take(sizeof(long long) * B);
  L.cand_q = take(sizeof(int) * (size_t)L.cap);
  L.cand_j = take(sizeof(long long) * (size_t)L.cap);
  L.cand_s = take(sizeof(double) * (size_t)L.cap);
  L.cand_f = take(sizeof(float) * (size_t)L.cap);
  L.total = off;
  return L;
}

}  // namespace

extern "C" {

size_t ef_match_tc_image_bytes_metric(int64_t n, int32_t k, int32_t metric) {
  if (n <= 0 || k <= 0) return 0;
  return image_body_bytes(n, k, metric) + kTrailer;
}

size_t ef_match_tc_image_bytes(int64_t n, int32_t k) { return ef_match_tc_image_bytes_metric(n, k, EF_METRIC_COSINE_SK); }

int ef_match_tc_prepare_device(const double* prepared, int64_t ldg, const double* norms, int64_t n, int32_t k,
                               int32_t metric, void* image, ef_stream_t stream) {
  if (!prepared || !image || n <= 0 || k <= 0 || ldg < k) return EF_ERR_INVALID;
  if (metric < EF_METRIC_COSINE_SK || metric > EF_METRIC_L2) return EF_ERR_INVALID;
  if (n_slabs_for(k, metric) > 49) return EF_ERR_UNSUPPORTED;    // k <= 1024 (ef_match_tc_device itself takes k <= 128)
  if (metric != EF_METRIC_COSINE_SK && !norms) return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  const size_t body = image_body_bytes(n, k, metric);
  EF_CUDA(cudaMemsetAsync(image, 0, body + kTrailer, st));
  double* gscale = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(image) + body);
  if (metric == EF_METRIC_L2)
    EF_LAUNCH(match_tc_maxnorm_kernel, (unsigned)std::min<int64_t>(1024, ef::ceil_div(n, 256)), 256, 0, st, norms,
              (long long)n, reinterpret_cast<unsigned long long*>(gscale));
  const int ns = n_slabs_for(k, metric);
  const int64_t work = n * ns * 8;
  EF_LAUNCH(match_tc_image_kernel, (unsigned)ef::ceil_div(work, 256), 256, 0, st, prepared, (long long)ldg, norms,
            (long long)n, k, ns, metric, reinterpret_cast<uint8_t*>(image), (const double*)gscale);
  return EF_OK;
}

size_t ef_match_tc_work_bytes(int32_t B, int64_t n, int32_t k) {
  if (B <= 0 || n <= 0 || k <= 0) return 256;
  return work_layout(B, n, k).total;
}

int ef_match_tc_device(const double* p, int64_t ldp, int32_t B, int32_t k, const double* prepared, int64_t ldg,
                       const double* norms, const void* image, int64_t n, int64_t index_base, int32_t metric,
                       double* out_score, int64_t* out_index, void* work, size_t work_bytes, ef_stream_t stream) {
  if (!p || !prepared || !image || !out_score || !out_index || !work || B < 0 || k <= 0 || ldp < k || ldg < k || n <= 0)
    return EF_ERR_INVALID;
  if (metric < EF_METRIC_COSINE_SK || metric > EF_METRIC_L2) return EF_ERR_INVALID;
  if (k > 128) return EF_ERR_UNSUPPORTED;
  if (metric == EF_METRIC_COSINE_G1 && !norms) return EF_ERR_INVALID;
  if (B == 0) return EF_OK;
  const Layout L = work_layout(B, n, k);
  if (work_bytes < L.total || (reinterpret_cast<uintptr_t>(work) & 255)) return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  char* w = reinterpret_cast<char*>(work);
  EF_CUDA(cudaMemsetAsync(w + L.status, 0, 512, st));           // status + counter words

  MatchTcArgs a{};
  a.P = p; a.ldp = ldp; a.B = B; a.k = k; a.n_slabs = n_slabs_for(k, metric); a.metric = metric;
  a.eps = metric == EF_METRIC_L2 ? kEpsL2 : kEps;
  a.gscale = reinterpret_cast<const double*>(reinterpret_cast<const uint8_t*>(image) + image_body_bytes(n, k, metric));
  a.img = reinterpret_cast<const __half*>(image);
  a.n = n; a.g_tiles = (int)ef::ceil_div(n, BN); a.tiles_per_chunk = L.tiles_per_chunk; a.b_pad = L.b_pad;
  a.cmax = reinterpret_cast<float*>(w + L.cmax);
  a.thr = reinterpret_cast<const float*>(w + L.thr);
  a.cand_q = reinterpret_cast<int*>(w + L.cand_q);
  a.cand_j = reinterpret_cast<long long*>(w + L.cand_j);
  a.cand_f = reinterpret_cast<float*>(w + L.cand_f);
  a.gmax = reinterpret_cast<unsigned int*>(w + L.gmax);
  a.counter = reinterpret_cast<unsigned int*>(w + L.counter);
  a.cap = L.cap;
  a.status = reinterpret_cast<int*>(w + L.status);
  const size_t fixed = (size_t)a.n_slabs * kSlabBytesA + sizeof(MatchTcShared) + 64;
  a.stages = (int)std::min<size_t>(kMaxStages, ((size_t)ef_tc::kSmemLimit - fixed) / kSlabBytesB);
  if (a.stages < 2) return EF_ERR_UNSUPPORTED;
  const size_t smem = fixed + (size_t)a.stages * kSlabBytesB;
  static size_t attr[64] = {0};                      // per device: a second GPU in the same process needs its own call
  int dev = 0;
  EF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return EF_ERR_UNSUPPORTED;
  if (smem > attr[dev]) {
    EF_CUDA(cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dev] = smem;
  }
  const dim3 grid((unsigned)ef::ceil_div(B, BLOCK_M), (unsigned)L.chunks);
  double* pn = reinterpret_cast<double*>(w + L.pn);
  unsigned long long* best_key = reinterpret_cast<unsigned long long*>(w + L.best_key);
  long long* best_idx = reinterpret_cast<long long*>(w + L.best_idx);
  double* cand_s = reinterpret_cast<double*>(w + L.cand_s);

  EF_LAUNCH(match_tc_norm_kernel, (unsigned)ef::ceil_div((int64_t)B * 32, 256), 256, 0, st, p, (long long)ldp, B, k,
            metric, pn, best_key, best_idx);
  const unsigned rgrid = (unsigned)std::min<int64_t>(4096, ef::ceil_div((int64_t)L.cap, 256));
  if (getenv("EF_MATCH_TC_TWO_PASS")) {
    // round-1 schedule (kept for A/B measurements): pass 0 = maxima, pass 1 = rows inside the band
    a.pass = 0;
    EF_LAUNCH(match_tc_kernel, grid, kThreads, smem, st, a);
    EF_LAUNCH(match_tc_thr_kernel, (unsigned)ef::ceil_div(L.b_pad, 256), 256, 0, st, a.cmax, L.chunks, L.b_pad, B,
              reinterpret_cast<float*>(w + L.thr), a.eps);
    a.pass = 1;
    EF_LAUNCH(match_tc_kernel, grid, kThreads, smem, st, a);
    EF_LAUNCH(match_tc_rescore_kernel, rgrid, 256, 0, st, p, (long long)ldp, k, prepared, (long long)ldg, norms, pn,
              a.cand_q, a.cand_j, a.counter, a.cap, metric, cand_s, best_key, (const float*)nullptr,
              (const float*)nullptr, (unsigned int*)nullptr);
  } else {
    EF_CUDA(cudaMemsetAsync(w + L.gmax, 0, sizeof(unsigned int) * L.b_pad, st));
    a.pass = 2;
    EF_LAUNCH(match_tc_kernel, grid, kThreads, smem, st, a);
    EF_LAUNCH(match_tc_thr_gmax_kernel, (unsigned)ef::ceil_div(L.b_pad, 256), 256, 0, st, a.gmax, L.b_pad, B,
              reinterpret_cast<float*>(w + L.thr), a.eps);
    EF_LAUNCH(match_tc_rescore_kernel, rgrid, 256, 0, st, p, (long long)ldp, k, prepared, (long long)ldg, norms, pn,
              a.cand_q, a.cand_j, a.counter, a.cap, metric, cand_s, best_key, a.cand_f, a.thr, a.counter + 2);
  }
  EF_LAUNCH(match_tc_pick_kernel, rgrid, 256, 0, st, a.cand_q, a.cand_j, cand_s, a.counter, a.cap, best_key, best_idx);
  EF_LAUNCH(match_tc_final_kernel, (unsigned)ef::ceil_div(B, 256), 256, 0, st, best_key, best_idx, B,
            (long long)index_base, (int)metric, out_score, reinterpret_cast<long long*>(out_index));
  return EF_OK;
}

/* [0] = pipeline-timeout flag, [1] = candidate count, [2] = candidate-list overflow (results incomplete: rerun with
 * ef_match_device).  Synchronous 12-byte read. */
int ef_match_tc_flags(const void* work, int32_t* flags3) {
  if (!work || !flags3) return EF_ERR_INVALID;
  const char* w = reinterpret_cast<const char*>(work);
  const Layout L = work_layout(1, 1, 1);             // status / counter offsets do not depend on the shape
  unsigned int c[3] = {0, 0, 0};                     // listed, overflow, kept after the final threshold (single pass)
  EF_CUDA(cudaMemcpy(&flags3[0], w + L.status, sizeof(int32_t), cudaMemcpyDeviceToHost));
  EF_CUDA(cudaMemcpy(c, w + L.counter, sizeof(c), cudaMemcpyDeviceToHost));
  flags3[1] = (int32_t)std::min<unsigned int>(c[2] ? c[2] : c[0], 0x7fffffffu);
  flags3[2] = (int32_t)c[1];
  return EF_OK;
}

}  // extern "C"
