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

constexpr int TM = 128, TN = 128, BK = 16;
constexpr int kThreads = 256;
constexpr int kStages = 3;
constexpr int LDK = BK + 4;        // k-major: [row][k], 20 doubles per row
constexpr int LDR = TM + 4;        // row-major: [k][row], 132 doubles per k
constexpr int kOperandDoubles = TM * LDK;                  // 2560 >= BK * LDR = 2112
constexpr int kStageDoubles = 2 * kOperandDoubles;
constexpr size_t kSmemBytes = sizeof(double) * (size_t)kStages * kStageDoubles;   // 122 880

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

// One operand tile (128 rows x 16 k) of element (row, k) = base[row * s_row + k * s_k] into shared memory.
// KMAJOR (s_k == 1): smem[row][LDK]; else (s_row == 1): smem[k][LDR].  Out-of-range elements become zeros.
template <bool KMAJOR>
__device__ __forceinline__ void load_tile(double* smem, const double* __restrict__ base, int64_t s_row, int64_t s_k,
                                          int row0, int rows, int k0, int k1, bool vec16, int tid) {
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = tid + kThreads * i;                       // 1024 chunks of two doubles
    int r, k;
    if (KMAJOR) { r = c >> 3; k = (c & 7) * 2; } else { k = c >> 6; r = (c & 63) * 2; }
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

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(kThreads, 1)
dgemm_tc_kernel(int M, int N, int K, int k_per_split, double alpha, const double* __restrict__ A, int64_t sam,
                int64_t sak, const double* __restrict__ Bm, int64_t sbk, int64_t sbn, double beta, double* __restrict__ C,
                int64_t ldc, double* __restrict__ partial, int vec_a, int vec_b) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;                  // 2 x 4 warps
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kb = blockIdx.z * k_per_split, ke = min(K, kb + k_per_split);
  const int n_chunks = (ke - kb + BK - 1) / BK;

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int chunk) {
    if (chunk < n_chunks) {
      double* sa = smem + (size_t)(chunk % kStages) * kStageDoubles;
      double* sb = sa + kOperandDoubles;
      const int k0 = kb + chunk * BK;
      load_tile<A_KMAJOR>(sa, A, sam, sak, m0, M, k0, ke, vec_a != 0, tid);
      load_tile<B_KMAJOR>(sb, Bm, sbn, sbk, n0, N, k0, ke, vec_b != 0, tid);
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
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        a[i] = A_KMAJOR ? sa[(wm * 64 + 8 * i + g) * LDK + 4 * s + t] : sa[(4 * s + t) * LDR + wm * 64 + 8 * i + g];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        b[j] = B_KMAJOR ? sb[(wn * 32 + 8 * j + g) * LDK + 4 * s + t] : sb[(4 * s + t) * LDR + wn * 32 + 8 * j + g];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
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
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + wn * 32 + 8 * j + 2 * t;
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

template <bool AK, bool BKM>
int launch(dim3 grid, cudaStream_t st, int M, int N, int K, int kps, double alpha, const double* A, int64_t sam, int64_t sak,
           const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc, double* partial, int va,
           int vb) {
  EF_ENSURE_SMEM((dgemm_tc_kernel<AK, BKM>), kSmemBytes);
  EF_LAUNCH((dgemm_tc_kernel<AK, BKM>), grid, kThreads, kSmemBytes, st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn, beta,
            C, ldc, partial, va, vb);
  return EF_OK;
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
  const dim3 grid((unsigned)ef::ceil_div(N, TN), (unsigned)ef::ceil_div(M, TM), (unsigned)splits);
  double* partial = splits > 1 ? reinterpret_cast<double*>(work) : nullptr;
  int rc;
  if (ak && bk) rc = launch<true, true>(grid, st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  else if (ak) rc = launch<true, false>(grid, st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  else if (bk) rc = launch<false, true>(grid, st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  else rc = launch<false, false>(grid, st, M, N, K, kps, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc, partial, va, vb);
  EF_TRY(rc);
  if (splits > 1) {
    const int64_t MN = (int64_t)M * N;
    EF_LAUNCH(dgemm_tc_reduce_kernel, (unsigned)std::min<int64_t>(1024, ef::ceil_div(MN, 256)), 256, 0, st, partial, splits,
              MN, N, alpha, beta, C, ldc);
  }
  return EF_OK;
}

}  // extern "C"
