// K2 epilogue: digit planes -> float64 features (+ reconstruction error) -> nearest gallery row -> threshold/label.
//
// Two kernels share the first phase:
//   fused_epilogue_kernel   k <= 32: features stay in registers / shared memory, the gallery streams through shared
//                           memory once per 32 crops, arg-best reduction and labelling in the same kernel (one launch
//                           after the projection instead of five);
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
constexpr int kWarps = 8;
constexpr int kThreads = QB * kWarps;

// P[b][c] = 2^e_c * sum_s acc[s*kq + c][b] * 2^-(7s+6) - bias[c]; small planes first, then clear the planes.
__device__ __forceinline__ double combine_planes(int32_t* __restrict__ acc_t, int ld_acc, int b, int c, int kq, int S,
                                                 int exp_c) {
  double v = 0.0;
  for (int s = S - 1; s >= 0; --s) {
    int32_t* p = acc_t + (size_t)(s * kq + c) * ld_acc + b;
    v += ldexp((double)*p, -(7 * s + 6));
    *p = 0;
  }
  return ldexp(v, exp_c);
}

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  if (METRIC == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

struct EpiArgs {
  int32_t* acc_t;
  int ld_acc, B, k, kq, S;
  const int32_t* col_exp;
  const double* bias;
  double* sumsq;          // consumed and cleared when resid is requested
  double c0;
  const double* gp;       // prepared gallery [n][k]
  const double* gnorm;    // [n] (COSINE_G1)
  int n;
  int tile_rows;          // gallery rows per shared-memory tile
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
  extern __shared__ double sm[];
  double* ps = sm;                         // [KR][QB] features (normalised for COSINE_SK)
  double* gs = ps + KR * QB;               // [tile_rows][KR] gallery tile, zero padded to KR
  double* gw = gs + (size_t)a.tile_rows * KR;   // [tile_rows] 1/|g| (COSINE_G1)
  __shared__ double pn_s[QB], xu_s[QB];
  __shared__ double red_s[kWarps][QB], red_d[kWarps][QB];
  __shared__ int red_i[kWarps][QB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x * QB + lane;
  const bool live = b < a.B;

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

  // ---- phase 2: gallery sweep, warp w takes rows w, w+8, ... of every tile
  double best = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF, best_dot = 0.0;
  int best_i = INT_MAX;
  for (int g0 = 0; g0 < a.n; g0 += a.tile_rows) {
    const int rows = min(a.tile_rows, a.n - g0);
    __syncthreads();
    for (int e = tid; e < rows * KR; e += kThreads) {
      const int r = e / KR, c = e - r * KR;
      gs[e] = (c < a.k) ? a.gp[(size_t)(g0 + r) * a.k + c] : 0.0;
    }
    if (METRIC == EF_METRIC_COSINE_G1)
      for (int r = tid; r < rows; r += kThreads) {
        const double gn = a.gnorm[g0 + r];
        gw[r] = gn == 0.0 ? 0.0 : 1.0 / gn;
      }
    __syncthreads();
    for (int r = warp; r < rows; r += kWarps) {
      const double* g = gs + (size_t)r * KR;
      double d = 0.0;
#pragma unroll
      for (int c = 0; c < KR; ++c) {
        if (METRIC == EF_METRIC_L2) {
          const double t = p[c] - g[c];
          d = fma(t, t, d);
        } else {
          d = fma(p[c], g[c], d);
        }
      }
      const double s = (METRIC == EF_METRIC_COSINE_G1) ? d * gw[r] : d;
      if (better<METRIC>(s, g0 + r, best, best_i)) {
        best = s;
        best_dot = d;
        best_i = g0 + r;
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

// Any k: features (+ residual) to global memory.
__global__ void __launch_bounds__(kThreads)
finalize_kernel(int32_t* __restrict__ acc_t, int ld_acc, int B, int k, int kq, int S,
                const int32_t* __restrict__ col_exp, const double* __restrict__ bias, double* __restrict__ proj,
                int64_t ldp, double* __restrict__ sumsq, double c0, double* __restrict__ resid2) {
  __shared__ double n2_s[kWarps][QB], xu_s[QB];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * QB + lane;
  const bool live = b < B;
  if (warp == 0) xu_s[lane] = 0.0;
  __syncthreads();
  double n2 = 0.0;
  for (int c = warp; c < kq; c += kWarps) {
    if (!live) continue;
    double v = combine_planes(acc_t, ld_acc, b, c, kq, S, col_exp[c]);
    if (c < k) {
      v -= bias[c];
      proj[(size_t)b * ldp + c] = v;
      n2 = fma(v, v, n2);
    } else {
      xu_s[lane] = v;
    }
  }
  n2_s[warp][lane] = n2;
  __syncthreads();
  if (warp == 0 && live && resid2) {
    double t = 0.0;
    for (int w = 0; w < kWarps; ++w) t += n2_s[w][lane];
    const double r = sumsq[b] - 2.0 * xu_s[lane] + c0 - t;
    resid2[b] = r > 0.0 ? r : 0.0;
    sumsq[b] = 0.0;
  }
}

template <int METRIC, int KR>
int launch_fused(const EpiArgs& a, cudaStream_t stream) {
  const size_t smem = sizeof(double) * ((size_t)KR * QB + (size_t)a.tile_rows * (KR + 1));
  static size_t attr = 0;
  if (smem > 48 * 1024 && smem > attr) {
    EF_CUDA(cudaFuncSetAttribute(fused_epilogue_kernel<METRIC, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  EF_LAUNCH((fused_epilogue_kernel<METRIC, KR>), (unsigned)ef::ceil_div(a.B, QB), kThreads, smem, stream, a);
  return EF_OK;
}

template <int METRIC>
int dispatch_kr(EpiArgs& a, cudaStream_t stream) {
  const int kr = a.k <= 8 ? 8 : (a.k <= 16 ? 16 : 32);
  // gallery tile: as much as fits in ~64 KB, at least 8 rows per warp round
  int rows = (int)((64 * 1024) / (sizeof(double) * (kr + 1)));
  rows = rows / kWarps * kWarps;
  if (rows > a.n) rows = a.n;
  a.tile_rows = rows;
  switch (kr) {
    case 8: return launch_fused<METRIC, 8>(a, stream);
    case 16: return launch_fused<METRIC, 16>(a, stream);
    default: return launch_fused<METRIC, 32>(a, stream);
  }
}

}  // namespace

namespace ef {

bool fused_epilogue_supported(int k, int64_t n) { return k <= 32 && n > 0 && n < (1ll << 31); }

int fused_epilogue(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp, const double* bias,
                   double* sumsq, double c0, const double* gp, const double* gnorm, int64_t n, const int32_t* labels,
                   int metric, double threshold, double* out_proj, double* out_score, int32_t* out_index,
                   int32_t* out_label, double* out_resid, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  EpiArgs a{acc_t, ld_acc, B, k, kq, S, col_exp, bias, sumsq, c0, gp, gnorm, (int)n, 0, labels, threshold,
            out_proj, out_score, out_index, out_label, out_resid};
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(a, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(a, stream);
    case EF_METRIC_L2: return dispatch_kr<EF_METRIC_L2>(a, stream);
    default: return EF_ERR_INVALID;
  }
}

int project_finalize(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp,
                     const double* bias, double* proj, int64_t ldp, double* sumsq, double c0, double* resid2,
                     cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  EF_LAUNCH(finalize_kernel, (unsigned)ceil_div(B, QB), kThreads, 0, stream, acc_t, ld_acc, B, k, kq, S, col_exp, bias,
            proj, ldp, sumsq, c0, resid2);
  return EF_OK;
}

}  // namespace ef
