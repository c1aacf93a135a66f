// K2 epilogue: digit planes -> float64 features (+ reconstruction error) -> nearest gallery row -> threshold/label.
//
// Two kernels share the first phase:
//   fused_epilogue_kernel   k <= 32: features stay in registers / shared memory, the (pre-padded) gallery is pulled
//                           into shared memory with cp.async while the planes are being combined, every warp sweeps a
//                           contiguous slice of gallery rows four at a time (independent DFMA chains), arg-best
//                           reduction and labelling in the same kernel: ONE launch after the projection;
//   finalize_kernel         any k: writes the features (and the residual) to global memory; the generic match
//                           kernel (ef_match.cu) follows.
// Both CONSUME AND CLEAR the int32 accumulators and the sum-of-squares buffer, so the next batch needs no memset.
//
// Replaces (per batch) the tail of project_face_to_eigenspace / pca.transform, the cosine loop + max of
// useless/scan.py:121-130 and cosine_similarity + argmax + threshold of scan-template-v4.py:274-287.
#include <climits>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

constexpr int QB = 32;          // crops per CTA (one per lane)
constexpr int kWarps = 16;
constexpr int kThreads = QB * kWarps;
constexpr int kSmemBudget = 200 * 1024;

// P[b][c] = 2^e_c * sum_s acc[s*kq + c][b] * 2^-(7s+6) - bias[c]; small planes first, then clear the planes.
__device__ __forceinline__ double combine_planes(int32_t* __restrict__ acc_t, int ld_acc, int b, int c, int kq, int S,
                                                 int exp_c) {
  // loads first, clearing stores afterwards (a store behind an in-flight load of the same address stalls the pipeline)
  int32_t plane[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) plane[s] = s < S ? __ldcg(acc_t + (size_t)(s * kq + c) * ld_acc + b) : 0;
#pragma unroll
  for (int s = 0; s < 8; ++s)
    if (s < S) __stcg(acc_t + (size_t)(s * kq + c) * ld_acc + b, 0);
  return ldexp(ef::planes_to_double(plane), exp_c);     // planes >= S are zero
}

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  if (METRIC == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

struct EpiArgs {
  int32_t* acc_t;
  int ld_acc, B, k, kq, S;
  const int32_t* col_exp;
  const double* bias;
  double* sumsq;          // consumed and cleared when resid is requested
  double c0;
  const double* gp;       // prepared gallery [n][KR] (zero padded to KR columns)
  const double* gnorm;    // [n] (COSINE_G1)
  const double* ginv;     // [n] 1/|g| or 0 (COSINE_G1)
  int n;
  int tile_rows;          // gallery rows per shared-memory tile (multiple of 4)
  const int32_t* labels;
  double threshold;
  double* out_proj;       // nullable
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;     // nullable
  double* out_resid;      // nullable
};

template <int METRIC, int KR>
__global__ void __launch_bounds__(kThreads)
fused_epilogue_kernel(const EpiArgs a) {
  // the projection of the next batch reads only until its own griddepcontrol.wait: it may be scheduled while this drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ __align__(16) double sm[];
  double* ps = sm;                                // [KR][QB] features
  double* gs = ps + KR * QB;                      // [tile_rows][KR] gallery tile
  double* gw = gs + (size_t)a.tile_rows * KR;     // [tile_rows] 1/|g| (COSINE_G1)
  __shared__ double pn_s[QB], xu_s[QB];
  __shared__ double red_s[kWarps][QB], red_d[kWarps][QB];
  __shared__ int red_i[kWarps][QB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x * QB + lane;
  const bool live = b < a.B;

  auto load_tile = [&](int g0) {
    const int rows = min(a.tile_rows, a.n - g0);
    const char* src = reinterpret_cast<const char*>(a.gp + (size_t)g0 * KR);
    const int chunks = rows * KR / 2;               // 16-byte chunks; KR is even
    for (int e = tid; e < chunks; e += kThreads) cp_async16(reinterpret_cast<char*>(gs) + e * 16, src + e * 16);
    if (METRIC == EF_METRIC_COSINE_G1)
      for (int e = tid; e < (rows + 1) / 2; e += kThreads) {
        // ginv is padded to an even count on the host side of the model
        cp_async16(reinterpret_cast<char*>(gw) + e * 16, reinterpret_cast<const char*>(a.ginv + g0) + e * 16);
      }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  load_tile(0);                                     // overlaps with the plane combination below

  // ---- phase 1: features
  for (int c = warp; c < KR; c += kWarps) ps[c * QB + lane] = 0.0;
  if (warp == 0) xu_s[lane] = 0.0;
  __syncthreads();
  for (int c = warp; c < a.kq; c += kWarps) {
    if (!live) continue;
    double v = combine_planes(a.acc_t, a.ld_acc, b, c, a.kq, a.S, a.col_exp[c]);
    if (c < a.k) {
      v -= a.bias[c];
      ps[c * QB + lane] = v;
      if (a.out_proj) a.out_proj[(size_t)b * a.k + c] = v;
    } else {
      xu_s[lane] = v;
    }
  }
  __syncthreads();
  if (warp == 0) {
    double n2 = 0.0;
    for (int c = 0; c < a.k; ++c) n2 = fma(ps[c * QB + lane], ps[c * QB + lane], n2);
    double pn = sqrt(n2);
    if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
    pn_s[lane] = pn;
    if (a.out_resid && live) {
      const double r = a.sumsq[b] - 2.0 * xu_s[lane] + a.c0 - n2;
      a.out_resid[b] = r > 0.0 ? r : 0.0;
      a.sumsq[b] = 0.0;
    }
  }
  __syncthreads();
  double p[KR];
#pragma unroll
  for (int c = 0; c < KR; ++c) {
    double v = ps[c * QB + lane];
    if (METRIC == EF_METRIC_COSINE_SK) v = v / pn_s[lane];        // sklearn normalize(): elementwise division
    p[c] = v;
  }

  // ---- phase 2: gallery sweep; warp w owns a contiguous slice of every tile, four rows per iteration
  double best = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF, best_dot = 0.0;
  int best_i = INT_MAX;
  for (int g0 = 0; g0 < a.n; g0 += a.tile_rows) {
    const int rows = min(a.tile_rows, a.n - g0);
    if (g0 > 0) {
      __syncthreads();                              // everyone is done with the previous tile
      load_tile(g0);
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    const int per = ((rows + kWarps - 1) / kWarps + 3) & ~3;      // rows per warp, multiple of 4
    const int r_begin = warp * per, r_end = min(rows, r_begin + per);
    for (int r = r_begin; r < r_end; r += 4) {
      double d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int c = 0; c < KR; c += 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // rows past the tile end read stale shared memory; their result is discarded below
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

  // ---- phase 3: reduce over the warps, score with the reference's own formula, threshold, label
  red_s[warp][lane] = best;
  red_d[warp][lane] = best_dot;
  red_i[warp][lane] = best_i;
  __syncthreads();
  if (warp == 0 && live) {
    double bs = red_s[0][lane], bd = red_d[0][lane];
    int bi = red_i[0][lane];
    for (int w = 1; w < kWarps; ++w)
      if (better<METRIC>(red_s[w][lane], red_i[w][lane], bs, bi)) {
        bs = red_s[w][lane];
        bd = red_d[w][lane];
        bi = red_i[w][lane];
      }
    double score = bs;
    if (METRIC == EF_METRIC_COSINE_G1) {
      // similarity = dot / (|p| |g|), zero norm -> 0.0   (useless/scan.py:70-77)
      const double pn = pn_s[lane], gn = a.gnorm[bi];
      score = (pn == 0.0 || gn == 0.0) ? 0.0 : bd / (pn * gn);
    }
    a.out_score[b] = score;
    a.out_index[b] = bi;
    if (a.out_label) {
      const bool pass = METRIC == EF_METRIC_L2 ? score <= a.threshold : score >= a.threshold;
      a.out_label[b] = pass ? (a.labels ? a.labels[bi] : bi) : -1;
    }
  }
}

// Any k: features (+ residual) to global memory, two kernels.
//   finalize_combine_kernel  one thread per (crop, column): the S plane loads of a column are independent and in flight
//                            together (crop index fastest: coalesced), the planes are cleared, the 32 x 32 tile of
//                            features is transposed through shared memory so that the feature rows are written
//                            coalesced; the residual column x.u goes to resid2 as a temporary
//   finalize_resid_kernel    one warp per crop: |p|^2 with a fixed summation order, then the reconstruction error
__global__ void __launch_bounds__(256)
finalize_combine_kernel(int32_t* __restrict__ acc_t, int ld_acc, int B, int k, int kq, int S,
                        const int32_t* __restrict__ col_exp, const double* __restrict__ bias, double* __restrict__ proj,
                        int64_t ldp, double* __restrict__ xu_out) {
  __shared__ double tile[32][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;          // 32 crops x 8 column phases
  const int b = blockIdx.x * 32 + lane;
  const int c_base = blockIdx.y * 32;
  // all plane loads of this thread first, the clearing stores afterwards: a store to an address whose load is still in
  // flight stalls the memory pipeline behind it (load latency x planes instead of one latency)
  int32_t plane[4][8];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = c_base + wy + 8 * it;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      plane[it][s] = (b < B && c < kq && s < S) ? __ldcg(acc_t + (size_t)(s * kq + c) * ld_acc + b) : 0;
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = c_base + wy + 8 * it;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (b < B && c < kq && s < S) __stcg(acc_t + (size_t)(s * kq + c) * ld_acc + b, 0);
  }
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int cy = wy + 8 * it;
    const int c = c_base + cy;
    double v = 0.0;
    if (b < B && c < kq) {
      v = ldexp(ef::planes_to_double(plane[it]), col_exp[c]);          // same combination as combine_planes
      if (c < k) v -= bias[c];
      else if (xu_out) xu_out[b] = v;                                  // the residual column x . u~
    }
    tile[cy][lane] = v;
  }
  __syncthreads();
  for (int by = wy; by < 32; by += 8) {
    const int bb = blockIdx.x * 32 + by, c = c_base + lane;
    if (bb < B && c < k) proj[(size_t)bb * ldp + c] = tile[lane][by];
  }
}

__global__ void __launch_bounds__(256)
finalize_resid_kernel(const double* __restrict__ proj, int64_t ldp, int B, int k, double* __restrict__ sumsq, double c0,
                      double* __restrict__ resid2) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch on both sides
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double n2 = 0.0;
  for (int c = lane; c < k; c += 32) {
    const double v = proj[(size_t)b * ldp + c];
    n2 = fma(v, v, n2);
  }
  n2 = ef::warp_sum(n2);
  if (lane == 0) {
    const double r = sumsq[b] - 2.0 * resid2[b] + c0 - n2;            // resid2[b] holds x . u~ from the combine kernel
    resid2[b] = r > 0.0 ? r : 0.0;
    sumsq[b] = 0.0;
  }
}

// Split-K slabs (ef_project_tc.cu, row-major part[split][crop][ld_part]) -> features: one thread per (crop, column),
// lanes along the columns of a plane (coalesced reads and feature writes), integer sum over the splits first (exact),
// then the same float64 combination as combine_planes.
__global__ void __launch_bounds__(256)
finalize_slabs_kernel(const int32_t* __restrict__ part, int splits, long long slab_stride, int ld_part, int B, int k,
                      int kq, int S, const int32_t* __restrict__ col_exp, const double* __restrict__ bias,
                      double* __restrict__ proj, int64_t ldp, double* __restrict__ xu_out) {
  // programmatic dependent launch on both sides: scheduled behind the projection (or the row sums), and the matcher's
  // query kernel may be scheduled behind this one
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // flattened (crop, column) index: every lane is busy whatever kq is; runs of kq consecutive lanes read one crop
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;          // B * kq < 2^31 (checked by the launcher)
  const int b = (int)(idx / (unsigned)kq);                             // 32-bit division: the 64-bit one costs more
  const int c = (int)(idx - (unsigned)b * (unsigned)kq);               // than the eight loads of the thread
  if (b >= B) return;
  int32_t plane[8];
#pragma unroll
  for (int s = 0; s < 8; ++s) plane[s] = 0;
  for (int sp = 0; sp < splits; ++sp) {
    const int32_t* src = part + (size_t)sp * slab_stride + (size_t)b * ld_part + c;
#pragma unroll
    for (int s = 0; s < 8; ++s)
      if (s < S) plane[s] += __ldcg(src + s * kq);
  }
  double v = ldexp(ef::planes_to_double(plane), col_exp[c]);            // same combination as combine_planes
  if (c < k) proj[(size_t)b * ldp + c] = v - bias[c];
  else if (xu_out) xu_out[b] = v;                                      // the residual column x . u~
}

// The same for COMBINED slabs (project_tc with combine: long long part[split][crop][ld_part / 4], component c at
// [2c] = hi, [2c + 1] = lo): the pairs of the splits add exactly, one rounding in hilo_to_double -- the same value.
__global__ void __launch_bounds__(256)
finalize_slabs_hilo_kernel(const long long* __restrict__ part, int splits, long long slab_stride, int ld64, int B, int k,
                           int kq, const int32_t* __restrict__ col_exp, const double* __restrict__ bias,
                           double* __restrict__ proj, int64_t ldp, double* __restrict__ xu_out, const ef::TcTail tail) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = (int)(idx / (unsigned)kq);
  const int c = (int)(idx - (unsigned)b * (unsigned)kq);
  if (b >= B) return;
  long long hi = 0, lo = 0;
  bool in_tail = false;
  if (tail.first >= 0) {
    // tail split of the projection: the tiles of its last, partial wave hold `tail.splits` partial tiles in the compact
    // region behind slab 0
    const int n_tile = (8 * c) / tail.block_n, m_tile = b >> 7;
    const int t = n_tile * tail.m_tiles + m_tile;
    if (t >= tail.first) {
      in_tail = true;
      const long long tile64 = (long long)128 * tail.block_n / 4;                      // long longs per partial tile
      const long long* base = part + tail.region / 4 + (long long)(t - tail.first) * tail.splits * tile64 +
                              (long long)(b & 127) * (tail.block_n / 4) + (8 * c - n_tile * tail.block_n) / 4;
      for (int sp = 0; sp < tail.splits; ++sp) {
        const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(base + (long long)sp * tile64));
        hi += v.x;
        lo += v.y;
      }
    }
  }
  if (!in_tail)
    for (int sp = 0; sp < splits; ++sp) {
      const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(part + (size_t)sp * slab_stride + (size_t)b * ld64) + c);
      hi += v.x;
      lo += v.y;
    }
  double v = ldexp(ef::hilo_to_double(hi, lo), col_exp[c]);
  if (c < k) proj[(size_t)b * ldp + c] = v - bias[c];
  else if (xu_out) xu_out[b] = v;
}

template <int METRIC, int KR>
int launch_fused(EpiArgs& a, cudaStream_t stream) {
  // gallery tile: as many rows as fit (the whole gallery for the shipped sizes), multiple of 4
  int rows = (int)((kSmemBudget - sizeof(double) * KR * QB) / (sizeof(double) * (KR + 1)));
  rows &= ~3;
  const int n4 = (a.n + 3) & ~3;
  if (rows > n4) rows = n4;
  a.tile_rows = rows;
  const size_t smem = sizeof(double) * ((size_t)KR * QB + (size_t)rows * (KR + 1)) + 64;
  // always: dynamic + ~11 KB of static shared memory crosses the 48 KB default well below 48 KB dynamic
  EF_ENSURE_SMEM((fused_epilogue_kernel<METRIC, KR>), smem);
  EF_LAUNCH((fused_epilogue_kernel<METRIC, KR>), (unsigned)ef::ceil_div(a.B, QB), kThreads, smem, stream, a);
  return EF_OK;
}

template <int METRIC>
int dispatch_kr(EpiArgs& a, int kr, cudaStream_t stream) {
  switch (kr) {
    case 4: return launch_fused<METRIC, 4>(a, stream);
    case 8: return launch_fused<METRIC, 8>(a, stream);
    case 12: return launch_fused<METRIC, 12>(a, stream);
    case 16: return launch_fused<METRIC, 16>(a, stream);
    case 24: return launch_fused<METRIC, 24>(a, stream);
    default: return launch_fused<METRIC, 32>(a, stream);
  }
}

}  // namespace

namespace ef {

bool fused_epilogue_supported(int k, int64_t n) { return k <= 32 && n > 0 && n < (1ll << 31) - 8; }

// Column padding of the prepared gallery used by the fused kernel (the kernel is specialised on it).
int fused_epilogue_kpad(int k) {
  const int steps[] = {4, 8, 12, 16, 24, 32};
  for (int s : steps)
    if (k <= s) return s;
  return k;
}

int fused_epilogue(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp, const double* bias,
                   double* sumsq, double c0, const double* gp_padded, const double* gnorm, const double* ginv, int64_t n,
                   const int32_t* labels, int metric, double threshold, double* out_proj, double* out_score,
                   int32_t* out_index, int32_t* out_label, double* out_resid, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  EpiArgs a{acc_t, ld_acc, B, k, kq, S, col_exp, bias, sumsq, c0, gp_padded, gnorm, ginv, (int)n, 0, labels, threshold,
            out_proj, out_score, out_index, out_label, out_resid};
  const int kr = fused_epilogue_kpad(k);
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(a, kr, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(a, kr, stream);
    case EF_METRIC_L2: return dispatch_kr<EF_METRIC_L2>(a, kr, stream);
    default: return EF_ERR_INVALID;
  }
}

int project_finalize_slabs(const int32_t* part, int splits, int ld_part, int B, int k, int kq, int S,
                           const int32_t* col_exp, const double* bias, double* proj, int64_t ldp, double* resid2,
                           cudaStream_t stream, bool combined, const TcTail* tail) {
  if (B <= 0) return EF_OK;
  if (S > 8 || (int64_t)B * kq >= (1ll << 31) - 256) return EF_ERR_INVALID;
  const unsigned grid = (unsigned)ceil_div((int64_t)B * kq, 256);
  if (combined) {
    if (S != 8 || (ld_part & 3)) return EF_ERR_INVALID;
    const TcTail tl = tail ? *tail : TcTail{-1, 1, 0, 0, 0};
    EF_LAUNCH_PDL(finalize_slabs_hilo_kernel, grid, 256, 0, stream, reinterpret_cast<const long long*>(part), splits,
                  (long long)B * (ld_part / 4), ld_part / 4, B, k, kq, col_exp, bias, proj, (int64_t)ldp,
                  (kq > k) ? resid2 : (double*)nullptr, tl);
    return EF_OK;
  }
  EF_LAUNCH_PDL(finalize_slabs_kernel, grid, 256, 0, stream, part, splits, (long long)B * ld_part, ld_part, B, k, kq, S,
                col_exp, bias, proj, (int64_t)ldp, (kq > k) ? resid2 : (double*)nullptr);
  return EF_OK;
}

// the residual pass alone: resid2 holds x . u~ on entry
int project_resid(const double* proj, int64_t ldp, int B, int k, double* sumsq, double c0, double* resid2,
                  cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  EF_LAUNCH_PDL(finalize_resid_kernel, (unsigned)ceil_div((int64_t)B * 32, 256), 256, 0, stream, proj, (int64_t)ldp, B, k,
                sumsq, c0, resid2);
  return EF_OK;
}

int project_finalize(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp,
                     const double* bias, double* proj, int64_t ldp, double* sumsq, double c0, double* resid2,
                     bool resid_pass, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  if (S > 8) return EF_ERR_INVALID;
  dim3 grid((unsigned)ceil_div(B, 32), (unsigned)ceil_div(kq, 32));
  EF_LAUNCH(finalize_combine_kernel, grid, 256, 0, stream, acc_t, ld_acc, B, k, kq, S, col_exp, bias, proj, ldp,
            (kq > k) ? resid2 : nullptr);
  if (resid2 && resid_pass)
    EF_LAUNCH(finalize_resid_kernel, (unsigned)ceil_div((int64_t)B * 32, 256), 256, 0, stream, proj, ldp, B, k, sumsq,
              c0, resid2);
  return EF_OK;
}

}  // namespace ef
