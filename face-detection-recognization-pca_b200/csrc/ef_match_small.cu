// K2b'': residual + nearest-gallery search + threshold/label in ONE launch for any k when the gallery is small enough
// for every CTA to sweep all of it (the shipped models: 229 x 50, 178 x 178, 590 x 50 rows x components).
//
// Same float64 arithmetic as the generic chain finalize_resid_kernel -> match_kernel -> match_reduce_kernel ->
// label_kernel (ef_epilogue.cu / ef_match.cu), bit for bit:
//   * |p|^2: lane-strided fma chains + xor-shuffle tree (finalize_resid_kernel / match_kernel's query norms);
//   * dot products: one fma chain per (query, gallery row) over ascending component index, starting from 0.0;
//   * scores: dot / (|p| |g|) (COSINE_G1, zero norm -> 0.0), dot of the elementwise-normalised rows (COSINE_SK),
//     squared distance (L2); first best wins (lowest index on ties).
// What changes is the schedule: 32 (or 16) queries per CTA stay in shared memory for all of k, the gallery streams
// through a double-buffered cp.async pipeline in 128-row x 32-component blocks, every thread owns 8 (4) queries x 2
// gallery rows, eight warps per CTA, and the COSINE_G1 division is only carried out for rows whose
// reciprocal-multiply estimate is within 1e-13 of the warp's best estimate for that query (the exact quotient decides,
// as before).
//
// Replaces (per batch) the cosine loop + max of useless/scan.py:121-130 and cosine_similarity + argmax + threshold of
// scan-template-v4.py:274-287 for the shapes the fused k <= 32 kernels do not cover.
#include <climits>
#include <cstdlib>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

constexpr int GT = 128;          // gallery rows per block
constexpr int KC = 32;           // components per block
constexpr int GS = KC + 2;       // row pitch of a staged block in doubles: 16-byte aligned, conflict-free LDS.128
constexpr int kThreads = 256;

struct MsArgs {
  const double* P;
  int64_t ldp;
  int B, k, kc_total;            // kc_total = k rounded up to KC
  const double* G;
  int64_t ldg;
  const double* gnorm;
  int n;
  const int32_t* labels;
  double threshold;
  double* sumsq;                 // nullable; consumed and cleared
  double c0;
  double* resid2;                // nullable; holds x . u~ on entry, the reconstruction error on exit
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;            // nullable
};

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  if (METRIC == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

// The 8 warps are (query group of QPT queries: warp & 3) x (half of a gallery block, 64 rows: warp >> 2).
template <int METRIC, int QPT>
__global__ void __launch_bounds__(kThreads)
match_small_kernel(const MsArgs a) {
  // the projection of the next batch (launched with programmatic stream serialization: it only reads until its own
  // griddepcontrol.wait) may be scheduled while this grid drains
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  constexpr int QT = 4 * QPT;
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) double sm[];
  const int kst = a.kc_total + 2;                  // query row pitch (even: double2 loads)
  double* ps = sm;                                 // [QT][kst]
  double* gs = ps + (size_t)QT * kst;              // [2][GT][GS]
  __shared__ double pn_s[QT];
  __shared__ double red_s[QT];
  __shared__ int red_i[QT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int qg = warp & 3, rh = warp >> 2;
  const int q0 = blockIdx.x * QT;
  const int n_chunks = a.kc_total / KC;
  const int n_tiles = (a.n + GT - 1) / GT;
  const int n_stages = n_tiles * n_chunks;

  auto issue = [&](int stage) {
    const int t = stage / n_chunks, c = stage - t * n_chunks;
    double* dst = gs + (size_t)(stage & 1) * GT * GS;
    const double* src = a.G + (int64_t)(t * GT) * a.ldg + c * KC;
    const int rows = min(GT, a.n - t * GT);
#pragma unroll 2
    for (int r = warp; r < GT; r += kWarps) {
#pragma unroll
      for (int kk = lane; kk < KC; kk += 32) {
        if (r < rows && c * KC + kk < a.k) cp_async8(dst + r * GS + kk, src + (int64_t)r * a.ldg + kk);
        else dst[r * GS + kk] = 0.0;
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  issue(0);

  // ---- queries: warp (qg, rh) brings in half of its group's rows (zero padded to a multiple of KC) and forms |p|^2
  // in the lane-strided order of finalize_resid_kernel / match_kernel; lanes 0.. then finish one query each
  {
    constexpr int H = QPT / 2;
    double n2[H];
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const int qi = qg * QPT + rh * H + i, q = q0 + qi;
      double acc = 0.0;
      for (int c = lane; c < a.kc_total; c += 32) {
        const double v = (q < a.B && c < a.k) ? a.P[(int64_t)q * a.ldp + c] : 0.0;
        ps[qi * kst + c] = v;
        if (c < a.k) acc = fma(v, v, acc);
      }
      n2[i] = ef::warp_sum(acc);
    }
    double mine = 0.0;
#pragma unroll
    for (int i = 0; i < H; ++i)
      if (lane == i) mine = n2[i];
    if (lane < H) {
      const int qi = qg * QPT + rh * H + lane, q = q0 + qi;
      double pn = sqrt(mine);
      if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
      pn_s[qi] = pn;
      if (a.resid2 && q < a.B) {
        const double r = a.sumsq[q] - 2.0 * a.resid2[q] + a.c0 - mine;
        a.resid2[q] = r > 0.0 ? r : 0.0;
        a.sumsq[q] = 0.0;
      }
    }
    if (METRIC == EF_METRIC_COSINE_SK) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const int qi = qg * QPT + rh * H + i;
        const double pn = pn_s[qi];
        for (int c = lane; c < a.k; c += 32) ps[qi * kst + c] = ps[qi * kst + c] / pn;   // sklearn normalize()
      }
    }
  }
  __syncthreads();
  double pn[QPT], pinv[QPT];
#pragma unroll
  for (int i = 0; i < QPT; ++i) {
    pn[i] = pn_s[qg * QPT + i];
    pinv[i] = (METRIC == EF_METRIC_COSINE_G1 && pn[i] != 0.0) ? 1.0 / pn[i] : 0.0;
  }

  double best[QPT];
  int best_i[QPT];
#pragma unroll
  for (int i = 0; i < QPT; ++i) {
    best[i] = (METRIC == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
    best_i[i] = INT_MAX;
  }
  double acc[QPT][2];

  for (int stage = 0; stage < n_stages; ++stage) {
    const int t = stage / n_chunks, c = stage - t * n_chunks;
    if (stage + 1 < n_stages) {
      issue(stage + 1);
      asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    if (c == 0) {
#pragma unroll
      for (int i = 0; i < QPT; ++i) acc[i][0] = acc[i][1] = 0.0;
    }
    const double* gb = gs + (size_t)(stage & 1) * GT * GS + (rh * 64 + lane) * GS;
    const double* pb = ps + (size_t)(qg * QPT) * kst + c * KC;
#pragma unroll 4
    for (int kk = 0; kk < KC; kk += 2) {
      double2 g[2], p[QPT];
#pragma unroll
      for (int j = 0; j < 2; ++j) g[j] = *reinterpret_cast<const double2*>(gb + j * 32 * GS + kk);
#pragma unroll
      for (int i = 0; i < QPT; ++i) p[i] = *reinterpret_cast<const double2*>(pb + i * kst + kk);
#pragma unroll
      for (int i = 0; i < QPT; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (METRIC == EF_METRIC_L2) {
            const double d0 = p[i].x - g[j].x;
            acc[i][j] = fma(d0, d0, acc[i][j]);
            const double d1 = p[i].y - g[j].y;
            acc[i][j] = fma(d1, d1, acc[i][j]);
          } else {
            acc[i][j] = fma(p[i].x, g[j].x, acc[i][j]);
            acc[i][j] = fma(p[i].y, g[j].y, acc[i][j]);
          }
        }
    }
    if (c == n_chunks - 1) {
      // ---- scores of this gallery block; ascending rows inside a thread, `better` orders across threads
      const int row0 = t * GT + rh * 64 + lane;
      double gn[2] = {1.0, 1.0}, ginv[2] = {0.0, 0.0};
      if (METRIC == EF_METRIC_COSINE_G1) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (row0 + 32 * j < a.n) {
            gn[j] = a.gnorm[row0 + 32 * j];
            ginv[j] = gn[j] != 0.0 ? 1.0 / gn[j] : 0.0;
          }
      }
#pragma unroll
      for (int i = 0; i < QPT; ++i) {
        if (METRIC == EF_METRIC_COSINE_G1) {
          // The quotient dot / (|p| |g|) is only formed for rows whose reciprocal-multiply estimate is within 1e-13
          // of the best estimate among the 64 rows this warp holds for the query: |estimate - quotient| is a few ulp
          // of a number <= 1, so the row(s) with the largest quotient always pass, and the quotient decides.
          double est[2], m = -CUDART_INF;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            est[j] = acc[i][j] * pinv[i] * ginv[j];
            if (row0 + 32 * j < a.n && est[j] > m) m = est[j];
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = row0 + 32 * j;
            if (row < a.n && est[j] >= m - 1e-13) {
              const double s = (pn[i] == 0.0 || gn[j] == 0.0) ? 0.0 : acc[i][j] / (pn[i] * gn[j]);
              if (better<METRIC>(s, row, best[i], best_i[i])) {
                best[i] = s;
                best_i[i] = row;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int row = row0 + 32 * j;
            if (row < a.n && better<METRIC>(acc[i][j], row, best[i], best_i[i])) {
              best[i] = acc[i][j];
              best_i[i] = row;
            }
          }
        }
      }
    }
    __syncthreads();                                 // the buffer is refilled two stages from now
  }

  // ---- the 32 lanes of the two warps (qg, 0) and (qg, 1) hold candidates for the same QPT queries
#pragma unroll
  for (int i = 0; i < QPT; ++i) {
    double s = best[i];
    int idx = best_i[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double s2 = __shfl_xor_sync(0xffffffffu, s, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, idx, o);
      if (better<METRIC>(s2, i2, s, idx)) {
        s = s2;
        idx = i2;
      }
    }
    best[i] = s;
    best_i[i] = idx;
  }
  if (rh == 1 && lane == 0) {
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      red_s[qg * QPT + i] = best[i];
      red_i[qg * QPT + i] = best_i[i];
    }
  }
  __syncthreads();
  if (rh == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < QPT; ++i) {
      const int qi = qg * QPT + i, q = q0 + qi;
      double s = best[i];
      int idx = best_i[i];
      if (better<METRIC>(red_s[qi], red_i[qi], s, idx)) {
        s = red_s[qi];
        idx = red_i[qi];
      }
      if (q < a.B) {
        const bool found = idx != INT_MAX;
        a.out_score[q] = s;
        a.out_index[q] = found ? idx : -1;
        if (a.out_label) {
          const bool pass = found && (METRIC == EF_METRIC_L2 ? s <= a.threshold : s >= a.threshold);
          a.out_label[q] = pass ? (a.labels ? a.labels[idx] : idx) : -1;
        }
      }
    }
  }
}

size_t smem_bytes(int qt, int kc_total) {
  return sizeof(double) * ((size_t)qt * (kc_total + 2) + 2 * (size_t)GT * GS);
}

template <int METRIC, int QPT>
int launch_q(const MsArgs& a, cudaStream_t stream) {
  const size_t smem = smem_bytes(4 * QPT, a.kc_total);
  EF_ENSURE_SMEM((match_small_kernel<METRIC, QPT>), smem);
  EF_LAUNCH((match_small_kernel<METRIC, QPT>), (unsigned)ef::ceil_div(a.B, 4 * QPT), kThreads, smem, stream, a);
  return EF_OK;
}

template <int METRIC>
int launch(const MsArgs& a, cudaStream_t stream) {
  // 32 queries per CTA while the query rows fit next to the gallery blocks and the sweep is long enough to amortise the
  // block staging; 16 for the largest k and for short sweeps (<= 6 blocks: twice the CTAs, two per SM, so that one
  // CTA's prologue and barriers overlap the other's arithmetic -- measured 5 us faster on the 229 x 50 gallery,
  // 4 us slower on the 178 x 178 one; 8 queries per CTA is slower again)
  const int n_stages = (int)ef::ceil_div(a.n, GT) * (a.kc_total / KC);
  const bool wide = smem_bytes(32, a.kc_total) <= 200 * 1024 && n_stages > 6;
  return wide ? launch_q<METRIC, 8>(a, stream) : launch_q<METRIC, 4>(a, stream);
}

}  // namespace

namespace ef {

// Every CTA sweeps the whole gallery: worth it while the gallery is small, or while the batch alone fills the GPU.
bool match_small_supported(int B, int k, int64_t n) {
  if (k <= 0 || n <= 0 || n >= (1ll << 31) - GT) return false;
  if (smem_bytes(16, (int)round_up(k, KC)) > 200 * 1024) return false;      // k <= 1024
  return n <= 1024 || (ceil_div(B, 32) >= sm_count() && n <= 16384);
}

int match_small(const double* proj, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm,
                int64_t n, const int32_t* labels, int metric, double threshold, double* sumsq, double c0, double* resid2,
                double* out_score, int32_t* out_index, int32_t* out_label, cudaStream_t stream) {
  if (B <= 0) return EF_OK;
  if (!match_small_supported(B, k, n)) return EF_ERR_UNSUPPORTED;
  MsArgs a{proj, ldp, B, k, (int)round_up(k, KC), gp, ldgp, gnorm, (int)n, labels, threshold,
           resid2 ? sumsq : nullptr, c0, resid2, out_score, out_index, out_label};
  switch (metric) {
    case EF_METRIC_COSINE_SK: return launch<EF_METRIC_COSINE_SK>(a, stream);
    case EF_METRIC_COSINE_G1: return launch<EF_METRIC_COSINE_G1>(a, stream);
    case EF_METRIC_L2: return launch<EF_METRIC_L2>(a, stream);
    default: return EF_ERR_INVALID;
  }
}

}  // namespace ef
