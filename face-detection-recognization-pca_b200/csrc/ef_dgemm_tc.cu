// Float64 GEMM on the FP64 tensor-core path (mma.sync m8n8k4 f64 -- the only FP64 matrix instruction Blackwell has;
// tcgen05 has no f64 kind) for the float64 products of the PCA fit: the covariance x block products of the subspace
// solver (config 4: 10 000 x 10 000 by 10 000 x 320, 64 Gflop each, ~150 per fit), its thin deflation / Rayleigh-Ritz /
// orthonormalisation products, and the projection of the training rows.
//
// Why not the CUDA-core kernels of ef_linalg.cu: an 8 x 8 register tile reads 2 bytes of shared memory per DFMA -- at 64
// DFMA per clock and SM that is the whole 128 B/clk shared-memory port, so dgemm_big_kernel stalls at 16 TFLOP/s (40 %
// of the FP64 peak).  A DMMA takes its 8 x 4 and 4 x 8 operands from ONE register per lane: a 64 x 32 warp tile needs
// 12 shared-memory loads per 32 DMMAs = 0.375 B per FMA.
//
// CTA tile 128 x 128 x 16, eight warps (2 x 4), warp tile 64 x 32 = 8 x 4 DMMA tiles, three cp.async stages.  Operands
// land in shared memory in whichever order their memory layout is contiguous (k-major rows padded to 20 doubles,
// row-major k-slices padded to 132): both give conflict-free fragment loads, so no transposing stores are needed.
// Split-K (grid.z) for products with a small output and a long K (Q^T Y: 320 x 320 x 10 000): every split writes its
// partial tile, a second kernel adds the partials in a fixed order -- the result is deterministic and independent of M,
// which is what keeps row-sharded products (multi-GPU solver) bit identical to the single-GPU ones.
//
// Replaces np.dot / np.cov products of useless/train.py:84-122 at the sizes where the reference calls LAPACK/BLAS.
#include <algorithm>

#include "ef_common.cuh"

namespace {

constexpr int TM = 128, BK = 16;
constexpr int kThreads = 256;
constexpr int kStages = 3;
constexpr int LDK = BK + 4;        // k-major: [row][k], 20 doubles per row
// row-major: [k][row], rows + 4 doubles per k (132 / 68: = 4 mod 16, conflict-free fragment loads)
__host__ __device__ constexpr int operand_doubles(int rows) { return rows * LDK; }          // >= BK * (rows + 4)
__host__ __device__ constexpr int stage_doubles(int tn) { return operand_doubles(TM) + operand_doubles(tn); }
constexpr size_t smem_bytes(int tn) { return sizeof(double) * (size_t)kStages * stage_doubles(tn); }   // 122 880 / 92 160

__device__ __forceinline__ void cp_async_16(unsigned dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_8(unsigned dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// One operand tile (ROWS rows x 16 k) of element (row, k) = base[row * s_row + k * s_k] into shared memory.
// KMAJOR (s_k == 1): smem[row][LDK]; else (s_row == 1): smem[k][ROWS + 4].  Out-of-range elements become zeros.
template <bool KMAJOR, int ROWS>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ base, int64_t s_row, int64_t s_k,
                                          int row0, int rows, int k0, int k1, bool vec16, int tid) {
  constexpr int LDR = ROWS + 4;
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
#pragma unroll
  for (int i = 0; i < ROWS * 8 / kThreads; ++i) {
    const int c = tid + kThreads * i;                       // ROWS x 8 chunks of two doubles
    int r, k;
    if (KMAJOR) { r = c >> 3; k = (c & 7) * 2; } else { k = c / (ROWS / 2); r = (c % (ROWS / 2)) * 2; }
    const int gr = row0 + r, gk = k0 + k;
    const unsigned dst = s0 + 8u * (unsigned)(KMAJOR ? r * LDK + k : k * LDR + r);
    // validity of the chunk's two elements (second element: next k or next row)
    int n_valid;
    if (KMAJOR) n_valid = (gr < rows) ? max(0, min(2, k1 - gk)) : 0;
    else n_valid = (gk < k1) ? max(0, min(2, rows - gr)) : 0;
    const double* src = base + (int64_t)(gr < rows ? gr : 0) * s_row + (int64_t)(gk < k1 ? gk : k0) * s_k;
    if (vec16) {
      cp_async_16(dst, n_valid ? src : base, 8 * n_valid);
    } else {
      cp_async_8(dst, n_valid >= 1 ? src : base, n_valid >= 1 ? 8 : 0);
      const double* src1 = src + (KMAJOR ? s_k : s_row);
      cp_async_8(dst + 8u, n_valid >= 2 ? src1 : base, n_valid >= 2 ? 8 : 0);
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR, int TN>
__global__ void __launch_bounds__(kThreads, TN == 128 ? 1 : 2)
dgemm_tc_kernel(int M, int N, int K, int k_per_split, double alpha, const double* __restrict__ A, int64_t sam,
                int64_t sak, const double* __restrict__ Bm, int64_t sbk, int64_t sbn, double beta, double* __restrict__ C,
                int64_t ldc, double* __restrict__ partial, int vec_a, int vec_b) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr int WN = TN / 4, NJ = WN / 8;                   // warp tile 64 x (TN / 4): 8 x NJ DMMA tiles
  constexpr int LDRA = TM + 4, LDRB = TN + 4;
  constexpr int kStageDoubles = stage_doubles(TN), kOperandDoubles = operand_doubles(TM);
  const int wm = warp >> 2, wn = warp & 3;                  // 2 x 4 warps
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  const int n_chunks = (ke - kb + BK - 1) / BK;

  double acc[8][NJ][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int chunk) {
    if (chunk < n_chunks) {
      double* sa = smem + (size_t)(chunk % kStages) * kStageDoubles;
      double* sb = sa + kOperandDoubles;
      const int k0 = kb + chunk * BK;
      load_tile<A_KMAJOR, TM>(sa, A, sam, sak, m0, M, k0, ke, vec_a != 0, tid);
      load_tile<B_KMAJOR, TN>(sb, Bm, sbn, sbk, n0, N, k0, ke, vec_b != 0, tid);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  issue(0);
  issue(1);
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    issue(chunk + 2);
    asm volatile("cp.async.wait_group 2;\n" ::);
    __syncthreads();
    const double* sa = smem + (size_t)(chunk % kStages) * kStageDoubles;
    const double* sb = sa + kOperandDoubles;
#pragma unroll
    for (int s = 0; s < BK / 4; ++s) {
      double a[8], b[NJ];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        a[i] = A_KMAJOR ? sa[(wm * 64 + 8 * i + g) * LDK + 4 * s + t] : sa[(4 * s + t) * LDRA + wm * 64 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        b[j] = B_KMAJOR ? sb[(wn * WN + 8 * j + g) * LDK + 4 * s + t] : sb[(4 * s + t) * LDRB + wn * WN + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncthreads();                                        // the stage is refilled by the next iteration's issue
  }
  asm volatile("cp.async.wait_group 0;\n" ::);

  // epilogue: lane holds C[8 i + g][8 j + 2 t + {0, 1}] of its warp tile
  const bool split = partial != nullptr;
  double* out = split ? partial + (size_t)blockIdx.z * (size_t)M * (size_t)N : C;
  const int64_t ldo = split ? N : ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + wm * 64 + 8 * i + g;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int n = n0 + wn * WN + 8 * j + 2 * t;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (n + e >= N) continue;
        double* p = out + (int64_t)m * ldo + n + e;
        if (split) {
          *p = acc[i][j][e];
        } else {
          double v = alpha * acc[i][j][e];
          if (beta != 0.0) v += beta * *p;
          *p = v;
        }
      }
    }
  }
}

// C = alpha * (partial[0] + partial[1] + ...) + beta * C, partials added in ascending split order
__global__ void dgemm_tc_reduce_kernel(const double* __restrict__ partial, int splits, int64_t MN, int N, double alpha,
                                       double beta, double* __restrict__ C, int64_t ldc) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < MN; e += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * MN + e];
    const int64_t m = e / N, n = e - m * N;
    double v = alpha * s;
    if (beta != 0.0) v += beta * C[m * ldc + n];
    C[m * ldc + n] = v;
  }
}

template <bool AK, bool BKM, int TN>
int launch(cudaStream_t st, int M, int N, int K, int kps, int splits, double alpha, const double* A, int64_t sam, int64_t sak,
           const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc, double* partial, int va,
           int vb) {
  const dim3 grid((unsigned)ef::ceil_div(N, TN), (unsigned)ef::ceil_div(M, TM), (unsigned)splits);
  EF_ENSURE_SMEM((dgemm_tc_kernel<AK, BKM, TN>), smem_bytes(TN));
  EF_LAUNCH((dgemm_tc_kernel<AK, BKM, TN>), grid, kThreads, smem_bytes(TN), st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn,
            beta, C, ldc, partial, va, vb);
  return EF_OK;
}

template <int TN>
int dispatch(bool ak, bool bk, cudaStream_t st, int M, int N, int K, int kps, int splits, double alpha, const double* A,
             int64_t sam, int64_t sak, const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc,
             double* partial, int va, int vb) {
  if (ak && bk) return launch<true, true, TN>(st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  if (ak) return launch<true, false, TN>(st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  if (bk) return launch<false, true, TN>(st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  return launch<false, false, TN>(st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
}

// Relative time of a grid of `tiles` CTAs when `per_sm` of them are resident per SM and a lone CTA on an SM runs at
// `solo` of the SM's rate: full rounds plus the cost of the ragged last round.
double grid_cost(int64_t tiles, int per_sm, double tile_time, double solo) {
  const int64_t sms = ef::sm_count(), slots = sms * per_sm;
  const int64_t full = tiles / slots, rest = tiles % slots;
  double cost = (double)full * per_sm * tile_time;
  if (rest > 0) cost += rest <= sms && per_sm > 1 ? tile_time / solo : per_sm * tile_time;
  return cost;
}

}  // namespace

extern "C" {

size_t ef_dgemm_tc_work_bytes(int32_t M, int32_t N, int32_t splits) {
  if (M <= 0 || N <= 0 || splits <= 1) return 0;
  return sizeof(double) * (size_t)M * (size_t)N * (size_t)splits;
}

int ef_dgemm_tc_device(int32_t M, int32_t N, int32_t K, double alpha, const double* A, int64_t sam, int64_t sak,
                       const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc, int32_t splits,
                       void* work, ef_stream_t stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0 || ldc < N || splits < 0) return EF_ERR_INVALID;
  if (M == 0 || N == 0) return EF_OK;
  // one of the two strides of each operand must be 1 (every caller in this library: row- or column-major views)
  if ((sak != 1 && sam != 1) || (sbk != 1 && sbn != 1)) return EF_ERR_UNSUPPORTED;
  cudaStream_t st = ef::as_stream(stream);
  if (splits <= 1) splits = 1;
  int kps = (int)ef::round_up(ef::ceil_div(std::max(K, 1), splits), BK);
  splits = (int)ef::ceil_div(std::max(K, 1), kps);
  if (splits > 1 && !work) return EF_ERR_INVALID;
  const bool ak = sak == 1, bk = sbk == 1;
  // 16-byte copies need an even leading stride and a 16-byte aligned base; otherwise 8-byte copies
  const int va = ((reinterpret_cast<uintptr_t>(A) & 15) == 0 && ((ak ? sam : sak) % 2 == 0)) ? 1 : 0;
  const int vb = ((reinterpret_cast<uintptr_t>(B) & 15) == 0 && ((bk ? sbn : sbk) % 2 == 0)) ? 1 : 0;
  double* partial = splits > 1 ? reinterpret_cast<double*>(work) : nullptr;
  // tile width: 128 x 128 (one CTA per SM) or 128 x 64 (two per SM), whichever fills the SMs' last round better.  Both
  // accumulate an output element over the same K chunks in the same order: the choice never changes a bit.
  const int64_t rows_t = ef::ceil_div(M, TM);
  const double cost128 = grid_cost(rows_t * ef::ceil_div(N, 128) * splits, 1, 1.0, 1.0);
  const double cost64 = grid_cost(rows_t * ef::ceil_div(N, 64) * splits, 2, 0.5 * 1.08, 0.7);
  int rc;
  if (cost64 < cost128)
    rc = dispatch<64>(ak, bk, st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  else
    rc = dispatch<128>(ak, bk, st, M, N, K, kps, splits, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  EF_TRY(rc);
  if (splits > 1) {
    const int64_t MN = (int64_t)M * N;
    EF_LAUNCH(dgemm_tc_reduce_kernel, (unsigned)std::min<int64_t>(1024, ef::ceil_div(MN, 256)), 256, 0, st, partial, splits,
              MN, N, alpha, beta, C, ldc);
  }
  return EF_OK;
}

}  // extern "C"
