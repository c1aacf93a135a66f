// K2b: nearest-gallery search in float64 with the reference's tie rule (first maximum wins).
//
// Replaces the Python loop over projected_data calling cosine_similarity + max() (useless/scan.py:58-78,121-127)
// and sklearn cosine_similarity + np.argmax (scan-template-v4.py:274-276).  A small float64 GEMM
// (32 queries x 128 gallery rows per CTA, k in chunks of 16 through shared memory) with the arg-best reduction
// fused in the epilogue: scores never touch HBM.
#include <algorithm>
#include <climits>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

constexpr int QT = 32;    // queries per CTA
constexpr int GT = 128;   // gallery rows per tile
constexpr int KC = 16;    // k chunk
constexpr int kThreads = 256;

struct Best {
  double s;
  long long i;
};

// true when candidate (s, i) beats (bs, bi): higher score (cosine) / lower distance (L2); ties -> lower index.
template <int METRIC>
__device__ __forceinline__ bool better(double s, long long i, double bs, long long bi) {
  if (METRIC == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

template <int METRIC>
__global__ void __launch_bounds__(kThreads)
match_kernel(const double* __restrict__ P, int64_t ldp, int B, int k, const double* __restrict__ G, int64_t ldg,
             const double* __restrict__ gnorm, int64_t n, int64_t index_base, int64_t rows_per_split,
             double* __restrict__ out_score, int64_t* __restrict__ out_index) {
  __shared__ double ps[KC][QT];
  __shared__ double gs[KC][GT + 2];
  __shared__ double pnorm[QT];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * QT;
  const int64_t g_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t g_end = min(n, g_begin + rows_per_split);

  // query norms (sqrt of the sum of squares; zero -> 1 for the sklearn rule)
  for (int qi = warp; qi < QT; qi += kThreads / 32) {
    const int q = q0 + qi;
    double s = 0.0;
    if (q < B)
      for (int c = lane; c < k; c += 32) {
        const double v = P[(int64_t)q * ldp + c];
        s += v * v;
      }
    s = ef::warp_sum(s);
    if (lane == 0) {
      double nrm = sqrt(s);
      if (METRIC == EF_METRIC_COSINE_SK && nrm == 0.0) nrm = 1.0;
      pnorm[qi] = nrm;
    }
  }
  __syncthreads();

  Best best[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    best[i].s = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
    best[i].i = LLONG_MAX;
  }

  for (int64_t gt = g_begin; gt < g_end; gt += GT) {
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int k0 = 0; k0 < k; k0 += KC) {
      __syncthreads();
      // queries: QT x KC -> ps[kk][q]
      for (int e = tid; e < QT * KC; e += kThreads) {
        const int qi = e / KC, kk = e % KC;
        const int q = q0 + qi, c = k0 + kk;
        double v = 0.0;
        if (q < B && c < k) {
          v = P[(int64_t)q * ldp + c];
          if (METRIC == EF_METRIC_COSINE_SK) v = v / pnorm[qi];
        }
        ps[kk][qi] = v;
      }
      // gallery: GT x KC -> gs[kk][row]; each thread takes one row half (8 contiguous doubles)
      {
        const int r = tid & (GT - 1), h = tid >> 7;
        const int64_t row = gt + r;
        const double* src = G + row * ldg + k0 + h * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = k0 + h * 8 + i;
          gs[h * 8 + i][r] = (row < g_end && c < k) ? src[i] : 0.0;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        double pv[4], gv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) pv[i] = ps[kk][warp * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) gv[j] = gs[kk][lane * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (METRIC == EF_METRIC_L2) {
              const double d = pv[i] - gv[j];
              acc[i][j] = fma(d, d, acc[i][j]);
            } else {
              acc[i][j] = fma(pv[i], gv[j], acc[i][j]);
            }
          }
      }
    }
    // epilogue for this gallery tile: ascending j keeps the lowest index on ties inside a thread
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t row = gt + lane * 4 + j;
      if (row >= g_end) continue;
      double gn = 1.0;
      if (METRIC == EF_METRIC_COSINE_G1) gn = gnorm[row];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        double s = acc[i][j];
        if (METRIC == EF_METRIC_COSINE_G1) {
          const double pn = pnorm[warp * 4 + i];
          s = (pn == 0.0 || gn == 0.0) ? 0.0 : s / (pn * gn);
        }
        if (better<METRIC>(s, row, best[i].s, best[i].i)) {
          best[i].s = s;
          best[i].i = row;
        }
      }
    }
  }

  // the 32 lanes of a warp hold candidates for the same 4 queries
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = best[i].s;
    long long idx = best[i].i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const long long i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      if (better<METRIC>(s2, i2, s, idx)) {
        s = s2;
        idx = i2;
      }
    }
    const int q = q0 + warp * 4 + i;
    if (lane == 0 && q < B) {
      out_score[(int64_t)blockIdx.y * B + q] = s;
      out_index[(int64_t)blockIdx.y * B + q] = (idx == LLONG_MAX) ? -1 : idx + index_base;
    }
  }
}

// A few queries (the reference's own call pattern is ONE face per call) against a gallery of any size: the tiled kernel
// above would run one query tile per gallery split through 37 load -> sync -> multiply rounds (110 us for 1 x 590 x 590),
// and so does any thread that walks its gallery row straight from global memory (every load its own cache line, 19
// warps on the whole GPU: 3 us per 16 elements).  Here a warp first pulls its 16 / 32 gallery rows WHOLE into shared
// memory with coalesced cp.async (all of the tile in flight at once: one memory latency), then lane r runs the same
// ascending fma chain per (query, row) as match_kernel over its row -- every score is bit identical to it.
constexpr int FQ = 8;        // queries per call of the few-query kernel

template <int METRIC>
__global__ void __launch_bounds__(32)
match_few_kernel(const double* __restrict__ P, int64_t ldp, int B, int k, const double* __restrict__ G, int64_t ldg,
                 const double* __restrict__ gnorm, int64_t n, int64_t index_base, int rows_per_cta,
                 double* __restrict__ out_score, int64_t* __restrict__ out_index) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch on both sides
  asm volatile("griddepcontrol.wait;" ::: "memory");
  extern __shared__ double fsm[];
  double* qs = fsm;                                // [B][k] queries (divided by their norm for the sklearn metric)
  const int kp = k | 1;                            // odd row pitch: lanes reading one column of 32 rows hit 32 banks
  double* tile = fsm + (size_t)B * k;              // [rows_per_cta][kp]
  __shared__ double pnorm[FQ];
  const int lane = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * rows_per_cta;
  const int rows = (int)min((int64_t)rows_per_cta, n - row0);
  // gallery rows -> shared memory, element by element (8-byte copies: any ld, any alignment), coalesced along a row
  {
    const unsigned t0 = (unsigned)__cvta_generic_to_shared(tile);
    for (int r = 0; r < rows; ++r) {
      const double* g = G + (row0 + r) * ldg;
      for (int c = lane; c < k; c += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(t0 + 8u * (unsigned)(r * kp + c)), "l"(g + c));
    }
    asm volatile("cp.async.commit_group;\n" ::);
  }
  for (int qi = 0; qi < B; ++qi) {                 // norms with match_kernel's arithmetic (lane-strided sums, xor tree)
    double s = 0.0;
    for (int c = lane; c < k; c += 32) {
      const double v = P[(int64_t)qi * ldp + c];
      s += v * v;
    }
    s = ef::warp_sum(s);
    double nrm = sqrt(s);
    if (METRIC == EF_METRIC_COSINE_SK && nrm == 0.0) nrm = 1.0;
    if (lane == 0) pnorm[qi] = nrm;
    for (int c = lane; c < k; c += 32) {
      double v = P[(int64_t)qi * ldp + c];
      if (METRIC == EF_METRIC_COSINE_SK) v = v / nrm;
      qs[qi * k + c] = v;
    }
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
  __syncwarp();
  const bool live = lane < rows;
  const int64_t row = row0 + lane;
  double acc[FQ];
#pragma unroll
  for (int i = 0; i < FQ; ++i) acc[i] = 0.0;
  if (live) {
    const double* g = tile + (size_t)lane * kp;
    for (int c = 0; c < k; ++c) {
      const double gv = g[c];
#pragma unroll
      for (int i = 0; i < FQ; ++i) {
        if (i < B) {
          if (METRIC == EF_METRIC_L2) {
            const double d = qs[i * k + c] - gv;
            acc[i] = fma(d, d, acc[i]);
          } else {
            acc[i] = fma(qs[i * k + c], gv, acc[i]);
          }
        }
      }
    }
  }
  double gn = 1.0;
  if (METRIC == EF_METRIC_COSINE_G1 && live) gn = gnorm[row];
#pragma unroll
  for (int i = 0; i < FQ; ++i) {
    if (i >= B) break;
    double s = acc[i];
    long long idx = live ? row : LLONG_MAX;
    if (!live) s = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
    else if (METRIC == EF_METRIC_COSINE_G1) {
      const double pn = pnorm[i];
      s = (pn == 0.0 || gn == 0.0) ? 0.0 : s / (pn * gn);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const long long i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      if (better<METRIC>(s2, i2, s, idx)) {
        s = s2;
        idx = i2;
      }
    }
    if (lane == 0) {
      out_score[(int64_t)blockIdx.x * B + i] = s;
      out_index[(int64_t)blockIdx.x * B + i] = (idx == LLONG_MAX) ? -1 : idx + index_base;
    }
  }
}

template <int METRIC>
__global__ void match_reduce_kernel(const double* __restrict__ scores, const int64_t* __restrict__ indices, int R,
                                    int B, double* __restrict__ out_score, int64_t* __restrict__ out_index) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch on both sides
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  double bs = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
  long long bi = LLONG_MAX;
  for (int r = 0; r < R; ++r) {
    const double s = scores[(int64_t)r * B + q];
    const long long i = indices[(int64_t)r * B + q];
    if (i < 0) continue;  // empty shard
    if (better<METRIC>(s, i, bs, bi)) {
      bs = s;
      bi = i;
    }
  }
  out_score[q] = bs;
  out_index[q] = (bi == LLONG_MAX) ? -1 : bi;
}

__global__ void gallery_prepare_kernel(const double* __restrict__ g, int64_t ldg, int64_t n, int k, int metric,
                                       double* __restrict__ gp, int64_t ldgp, double* __restrict__ gnorm,
                                       double* __restrict__ ginv) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* src = g + row * ldg;
  double s = 0.0;
  for (int c = lane; c < k; c += 32) s += src[c] * src[c];
  s = ef::warp_sum(s);
  const double nrm = sqrt(s);
  const double div = (metric == EF_METRIC_COSINE_SK) ? (nrm == 0.0 ? 1.0 : nrm) : 1.0;
  for (int c = lane; c < k; c += 32) gp[row * ldgp + c] = (metric == EF_METRIC_COSINE_SK) ? src[c] / div : src[c];
  if (lane == 0 && gnorm) gnorm[row] = nrm;
  if (lane == 0 && ginv) ginv[row] = nrm == 0.0 ? 0.0 : 1.0 / nrm;
}

__global__ void label_kernel(const double* __restrict__ score, const int64_t* __restrict__ index, int B,
                             const int32_t* __restrict__ labels, int metric, double threshold,
                             int32_t* __restrict__ out_index, int32_t* __restrict__ out_label) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch on both sides
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  const long long idx = index[q];
  if (out_index) out_index[q] = (int32_t)idx;
  if (out_label) {
    const double s = score[q];
    const bool pass = idx >= 0 && (metric == EF_METRIC_L2 ? s <= threshold : s >= threshold);
    out_label[q] = pass ? (labels ? labels[idx] : (int32_t)idx) : -1;
  }
}

int gallery_splits(int B, int64_t n) {
  const int qtiles = (int)ef::ceil_div(B, QT);
  int64_t want = ef::ceil_div(2 * (int64_t)ef::sm_count(), qtiles);
  const int64_t max_splits = ef::ceil_div(n, GT);
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace

namespace ef {

int gallery_prepare(const double* g, int64_t ldg, int64_t n, int k, int metric, double* gp, int64_t ldgp, double* gnorm,
                    double* ginv, cudaStream_t stream) {
  if (n <= 0) return EF_OK;
  const int threads = 256;
  const int64_t grid = ceil_div(n * 32, threads);
  EF_LAUNCH(gallery_prepare_kernel, (unsigned)grid, threads, 0, stream, g, ldg, n, k, metric, gp, ldgp, gnorm, ginv);
  return EF_OK;
}

// few-query kernel: rows per CTA so that the tile + the queries fit shared memory (k <= 1024: 16 rows, k <= 448: 32)
static int few_rows(int B, int k) {
  const size_t q = sizeof(double) * (size_t)B * k;
  for (int rows : {32, 16, 8})
    if (q + sizeof(double) * (size_t)rows * (k | 1) <= 200 * 1024) return rows;
  return 0;
}
static bool few_queries(int B, int k) { return B <= FQ && few_rows(B, k) > 0; }

size_t match_work_bytes(int B, int64_t n) {
  const int splits = gallery_splits(B, n);
  size_t bytes = splits > 1 ? (size_t)splits * B * (sizeof(double) + sizeof(int64_t)) : 0;
  if (B <= FQ) bytes = std::max(bytes, (size_t)ceil_div(n, 8) * B * (sizeof(double) + sizeof(int64_t)));
  return bytes;
}

template <int METRIC>
static int match_impl(const double* p, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm,
                      int64_t n, int64_t index_base, double* out_score, int64_t* out_index, void* work,
                      cudaStream_t stream) {
  if (work && few_queries(B, k) && n > 32) {
    // 8 / 16 / 32 gallery rows per one-warp CTA, per-CTA bests reduced by match_reduce_kernel
    const int rows = few_rows(B, k);
    const int ctas = (int)ceil_div(n, rows);
    const size_t smem = sizeof(double) * ((size_t)B * k + (size_t)rows * (k | 1));
    double* ws = reinterpret_cast<double*>(work);
    int64_t* wi = reinterpret_cast<int64_t*>(ws + (size_t)ctas * B);
    EF_ENSURE_SMEM(match_few_kernel<METRIC>, smem);
    // the one-face chain (residual -> few-query match -> reduce -> label) is launch bound: every kernel of it may be
    // scheduled while its predecessor drains (programmatic dependent launch; each waits before it reads)
    EF_LAUNCH_PDL(match_few_kernel<METRIC>, (unsigned)ctas, 32, smem, stream, p, ldp, B, k, gp, ldgp, gnorm, n, index_base,
                  rows, ws, wi);
    EF_LAUNCH_PDL(match_reduce_kernel<METRIC>, (unsigned)ceil_div(B, 256), 256, 0, stream, (const double*)ws,
                  (const int64_t*)wi, ctas, B, out_score, out_index);
    return EF_OK;
  }
  const int splits = work ? gallery_splits(B, n) : 1;
  const int64_t rows_per_split = round_up(ceil_div(n, splits), GT);
  const int real_splits = (int)ceil_div(n, rows_per_split);
  dim3 grid((unsigned)ceil_div(B, QT), (unsigned)real_splits);
  if (real_splits == 1) {
    EF_LAUNCH(match_kernel<METRIC>, grid, kThreads, 0, stream, p, ldp, B, k, gp, ldgp, gnorm, n, index_base,
              rows_per_split, out_score, out_index);
    return EF_OK;
  }
  double* ws = reinterpret_cast<double*>(work);
  int64_t* wi = reinterpret_cast<int64_t*>(ws + (size_t)splits * B);
  EF_LAUNCH(match_kernel<METRIC>, grid, kThreads, 0, stream, p, ldp, B, k, gp, ldgp, gnorm, n, index_base,
            rows_per_split, ws, wi);
  EF_LAUNCH(match_reduce_kernel<METRIC>, (unsigned)ceil_div(B, 256), 256, 0, stream, ws, wi, real_splits, B, out_score,
            out_index);
  return EF_OK;
}

int match(const double* p, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm, int64_t n,
          int64_t index_base, int metric, double* out_score, int64_t* out_index, void* work, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  if (n <= 0) return EF_ERR_INVALID;
  switch (metric) {
    case EF_METRIC_COSINE_SK:
      return match_impl<EF_METRIC_COSINE_SK>(p, ldp, B, k, gp, ldgp, gnorm, n, index_base, out_score, out_index, work, stream);
    case EF_METRIC_COSINE_G1:
      return match_impl<EF_METRIC_COSINE_G1>(p, ldp, B, k, gp, ldgp, gnorm, n, index_base, out_score, out_index, work, stream);
    case EF_METRIC_L2:
      return match_impl<EF_METRIC_L2>(p, ldp, B, k, gp, ldgp, gnorm, n, index_base, out_score, out_index, work, stream);
    default:
      return EF_ERR_INVALID;
  }
}

int label_lookup(const double* score, const int64_t* index, int B, const int32_t* labels, int metric, double threshold,
                 int32_t* out_index32, int32_t* out_label, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  EF_LAUNCH_PDL(label_kernel, (unsigned)ceil_div(B, 256), 256, 0, stream, score, index, B, labels, metric, threshold,
                out_index32, out_label);
  return EF_OK;
}

}  // namespace ef

extern "C" {

size_t ef_match_work_bytes(int32_t B, int64_t n) { return ef::match_work_bytes(B, n); }

int ef_gallery_prepare_device(const double* gallery, int64_t ldg, int64_t n, int32_t k, int32_t metric, double* prepared,
                              int64_t ldp, double* norms, ef_stream_t stream) {
  if (!gallery || !prepared || n < 0 || k <= 0 || ldg < k || ldp < k) return EF_ERR_INVALID;
  if (metric == EF_METRIC_COSINE_G1 && !norms) return EF_ERR_INVALID;
  return ef::gallery_prepare(gallery, ldg, n, k, metric, prepared, ldp, norms, nullptr, ef::as_stream(stream));
}

int ef_match_device(const double* p, int64_t ldp, int32_t B, int32_t k, const double* prepared, int64_t ldg,
                    const double* norms, int64_t n, int64_t index_base, int32_t metric, double* out_score,
                    int64_t* out_index, void* work, ef_stream_t stream) {
  if (!p || !prepared || !out_score || !out_index || B < 0 || k <= 0 || ldp < k || ldg < k) return EF_ERR_INVALID;
  if (metric == EF_METRIC_COSINE_G1 && !norms) return EF_ERR_INVALID;
  return ef::match(p, ldp, B, k, prepared, ldg, norms, n, index_base, metric, out_score, out_index, work,
                   ef::as_stream(stream));
}

int ef_match_reduce_device(const double* scores, const int64_t* indices, int32_t R, int32_t B, int32_t metric,
                           double* out_score, int64_t* out_index, ef_stream_t stream) {
  if (!scores || !indices || !out_score || !out_index || R <= 0 || B < 0) return EF_ERR_INVALID;
  if (B == 0) return EF_OK;
  cudaStream_t st = ef::as_stream(stream);
  const unsigned grid = (unsigned)ef::ceil_div(B, 256);
  switch (metric) {
    case EF_METRIC_COSINE_SK:
      EF_LAUNCH(match_reduce_kernel<EF_METRIC_COSINE_SK>, grid, 256, 0, st, scores, indices, R, B, out_score, out_index);
      return EF_OK;
    case EF_METRIC_COSINE_G1:
      EF_LAUNCH(match_reduce_kernel<EF_METRIC_COSINE_G1>, grid, 256, 0, st, scores, indices, R, B, out_score, out_index);
      return EF_OK;
    case EF_METRIC_L2:
      EF_LAUNCH(match_reduce_kernel<EF_METRIC_L2>, grid, 256, 0, st, scores, indices, R, B, out_score, out_index);
      return EF_OK;
    default:
      return EF_ERR_INVALID;
  }
}

}  // extern "C"
