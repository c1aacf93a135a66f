// Cholesky factor of a small symmetric positive definite matrix and the inverse of that factor, in ONE launch of ONE
// CTA: the orthonormalisation step (CholeskyQR) of the subspace solver behind the config-4 PCA fit.
//
// The reference gets its eigenvectors from np.linalg.eigh of the whole D x D covariance (useless/train.py:103); for
// D = 10 000 this engine iterates on an n x m block instead (dist.eigh_topk_device) and has to re-orthonormalise the
// block after every filter application: G = Y^T Y (ef_dgemm_device), G = L L^T and L^-1 (here), Q = Y L^-T
// (ef_dgemm_device).  The m x m work is tiny (m <= 640: m^3 / 3 = 11 Mflop at m = 320) and strictly sequential across
// panels, so it is latency, not throughput: one CTA of 1024 threads walks 32-column panels with the panel in shared
// memory and the trailing matrix L2 resident -- ~0.1 ms instead of the 10 ms of a 320 x 320 Jacobi eigensolve.
#include "ef_common.cuh"

namespace {

constexpr int kNB = 32;
constexpr int kPad = kNB + 1;
constexpr int kCholThreads = 1024;
constexpr int kCholMax = 640;

__global__ void __launch_bounds__(kCholThreads, 1)
chol_inverse_kernel(double* A, int m, double* Linv, int* info) {
  extern __shared__ double sm[];
  double* Dg = sm;                       // [32][33] diagonal block of L (identity padded)
  double* Dv = Dg + kNB * kPad;          // [32][33] inverse of a diagonal block
  double* Ts = Dv + kNB * kPad;          // [32][33] block product
  double* P = Ts + kNB * kPad;           // [m][33] panel below the diagonal block
  __shared__ int bad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) bad = 0;
  __syncthreads();

  // ------------------------------------------------------------------ right-looking blocked Cholesky (lower triangle)
  for (int j0 = 0; j0 < m; j0 += kNB) {
    const int nb = min(kNB, m - j0);
    for (int e = tid; e < kNB * kNB; e += kCholThreads) {
      const int r = e / kNB, c = e - r * kNB;
      Dg[r * kPad + c] = (r < nb && c < nb) ? A[(size_t)(j0 + r) * m + j0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      for (int k = 0; k < nb; ++k) {
        const double dkk = Dg[k * kPad + k];
        if (!(dkk > 0.0)) {                       // not positive definite to working precision (also catches NaN)
          if (lane == 0) bad = j0 + k + 1;
          break;
        }
        const double s = sqrt(dkk);
        __syncwarp();
        if (lane == k) Dg[k * kPad + k] = s;
        double lik = 0.0;
        if (lane > k) {
          lik = Dg[lane * kPad + k] / s;
          Dg[lane * kPad + k] = lik;
        }
        __syncwarp();
        if (lane > k)
          for (int c = k + 1; c <= lane; ++c) Dg[lane * kPad + c] -= lik * Dg[c * kPad + k];
        __syncwarp();
      }
    }
    __syncthreads();
    if (bad) break;
    for (int e = tid; e < nb * nb; e += kCholThreads) {
      const int r = e / nb, c = e - r * nb;
      A[(size_t)(j0 + r) * m + j0 + c] = c <= r ? Dg[r * kPad + c] : 0.0;
    }
    // panel: rows below the diagonal block, one thread per row: x D^T = a  (forward substitution, D in shared memory)
    const int rows = m - j0 - nb;
    for (int i = tid; i < rows; i += kCholThreads) {
      double* a = A + (size_t)(j0 + nb + i) * m + j0;
      double x[kNB];
#pragma unroll
      for (int c = 0; c < kNB; ++c) x[c] = c < nb ? a[c] : 0.0;
#pragma unroll
      for (int c = 0; c < kNB; ++c) {
        double s = x[c];
#pragma unroll
        for (int t = 0; t < c; ++t) s -= x[t] * Dg[c * kPad + t];
        x[c] = s / Dg[c * kPad + c];
      }
#pragma unroll
      for (int c = 0; c < kNB; ++c) {
        if (c < nb) a[c] = x[c];
        P[i * kPad + c] = x[c];
      }
    }
    __syncthreads();
    // trailing update (lower triangle): A22 -= P P^T
    for (int e = tid; e < rows * rows; e += kCholThreads) {
      const int i = e / rows, c = e - i * rows;
      if (c > i) continue;
      double s = 0.0;
#pragma unroll
      for (int t = 0; t < kNB; ++t) s += P[i * kPad + t] * P[c * kPad + t];
      A[(size_t)(j0 + nb + i) * m + j0 + nb + c] -= s;
    }
    __syncthreads();
  }
  if (bad) {
    if (tid == 0) *info = bad;
    return;
  }

  // ------------------------------------------------------------------ inverse of the lower-triangular factor, by blocks:
  // Linv[I][I] = inv(L[I][I]);  Linv[I][J] = -inv(L[I][I]) * sum_{J <= K < I} L[I][K] Linv[K][J]   (J < I)
  for (int e = tid; e < m * m; e += kCholThreads) Linv[e] = 0.0;
  __syncthreads();
  const int nblk = (m + kNB - 1) / kNB;
  const int r = tid >> 5, c = lane;                // this thread's element of a 32 x 32 block
  for (int I = 0; I < nblk; ++I) {
    const int i0 = I * kNB, nbI = min(kNB, m - i0);
    for (int e = tid; e < kNB * kNB; e += kCholThreads) {
      const int rr = e / kNB, cc = e - rr * kNB;
      Dg[rr * kPad + cc] = (rr < nbI && cc < nbI && cc <= rr) ? A[(size_t)(i0 + rr) * m + i0 + cc] : (rr == cc ? 1.0 : 0.0);
    }
    __syncthreads();
    if (warp == 0) {
      // column `lane` of the inverse of the diagonal block by forward substitution
      double x[kNB];
#pragma unroll
      for (int rr = 0; rr < kNB; ++rr) {
        double s = rr == lane ? 1.0 : 0.0;
#pragma unroll
        for (int t = 0; t < rr; ++t) s -= Dg[rr * kPad + t] * x[t];
        x[rr] = s / Dg[rr * kPad + rr];
      }
#pragma unroll
      for (int rr = 0; rr < kNB; ++rr) Dv[rr * kPad + lane] = x[rr];
    }
    __syncthreads();
    if (r < nbI && c < nbI) Linv[(size_t)(i0 + r) * m + i0 + c] = c <= r ? Dv[r * kPad + c] : 0.0;
    for (int J = 0; J < I; ++J) {
      const int j0 = J * kNB;
      double t = 0.0;
      if (r < nbI) {
        const double* lrow = A + (size_t)(i0 + r) * m;
        const double* col = Linv + j0 + c;
#pragma unroll 4
        for (int kk = j0; kk < i0; ++kk) t += lrow[kk] * col[(size_t)kk * m];
      }
      Ts[r * kPad + c] = t;
      __syncthreads();
      double o = 0.0;
#pragma unroll
      for (int rr = 0; rr < kNB; ++rr) o -= Dv[r * kPad + rr] * Ts[rr * kPad + c];
      if (r < nbI) Linv[(size_t)(i0 + r) * m + j0 + c] = o;
      __syncthreads();
    }
    __syncthreads();                               // block row I of Linv is complete before block row I + 1 reads it
  }
  if (tid == 0) *info = 0;
}

}  // namespace

extern "C" int ef_chol_inverse_device(double* G, int32_t m, double* Linv, int32_t* info, ef_stream_t stream) {
  if (!G || !Linv || !info || m <= 0) return EF_ERR_INVALID;
  if (m > kCholMax) return EF_ERR_UNSUPPORTED;
  const size_t smem = sizeof(double) * ((size_t)3 * kNB * kPad + (size_t)m * kPad);
  static size_t attr[64] = {0};
  int dev = 0;
  EF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return EF_ERR_UNSUPPORTED;
  if (smem > attr[dev]) {
    EF_CUDA(cudaFuncSetAttribute(chol_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr[dev] = smem;
  }
  EF_LAUNCH(chol_inverse_kernel, 1, kCholThreads, smem, ef::as_stream(stream), G, (int)m, Linv, info);
  return EF_OK;
}
