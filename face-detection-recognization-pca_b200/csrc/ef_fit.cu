// PCA fit orchestration (host entry points): Gen-1 manual_pca and Gen-2 StandardScaler + PCA(full).
//
//   ef_fit_gen1_host   useless/train.py:56-128   mean, centre, Gram (N<D) or covariance, eigh, back-projection,
//                                                column normalisation, descending sort, top-k, projection
//   ef_fit_gen2_host   train-v5.py:349-385       pixel mean, StandardScaler.fit_transform, PCA(k).fit_transform
//                                                (solver "full": SVD of the centred matrix, svd_flip, U*S)
// Everything numeric runs on the device in float64; the host only sequences launches and does O(n) bookkeeping on
// the eigenvalue vector.
#include <climits>
#include <cmath>
#include <vector>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

__global__ void mean_from_colsum_kernel(const long long* __restrict__ colsum, int D, long long N,
                                        double* __restrict__ mean) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) mean[d] = (double)colsum[d] / (double)N;   // np.mean: exact integer sum, one division
}

// sklearn _incremental_mean_and_var (first batch): per column, sequential over rows like np.sum(axis=0).
__global__ void scaler_var_kernel(const uint8_t* __restrict__ X, int64_t ldx, int64_t N, int D,
                                  const double* __restrict__ mean, double* __restrict__ var,
                                  double* __restrict__ scale) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const double T = mean[d];
  double corr = 0.0, s2 = 0.0;
  for (int64_t n = 0; n < N; ++n) {
    const double t = (double)X[n * ldx + d] - T;
    corr += t;
    s2 += t * t;
  }
  const double nn = (double)N;
  double v = (s2 - corr * corr / nn) / nn;
  var[d] = v;
  const double eps = 2.220446049250313e-16;
  const double nme = nn * T * eps;
  const bool constant = v <= nn * eps * v + nme * nme;
  double sc = sqrt(v);
  if (constant || sc == 0.0) sc = 1.0;
  scale[d] = sc;
}

// ManualStandardScaler.fit (scripts/manual/train-v2.py:57-63): np.std (population) per column, exact zeros -> 1.
__global__ void scaler_var_manual_kernel(const uint8_t* __restrict__ X, int64_t ldx, int64_t N, int D,
                                         const double* __restrict__ mean, double* __restrict__ var,
                                         double* __restrict__ scale) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const double T = mean[d];
  double s2 = 0.0;
  for (int64_t n = 0; n < N; ++n) {
    const double t = (double)X[n * ldx + d] - T;
    s2 += t * t;
  }
  const double v = s2 / (double)N;
  var[d] = v;
  double sc = sqrt(v);
  if (sc == 0.0) sc = 1.0;
  scale[d] = sc;
}

// column means of a float64 matrix, sequential over rows (np.mean(axis=0) order), then subtract in place
__global__ void colmean_center_kernel(double* __restrict__ Z, int64_t ldz, int64_t N, int D,
                                      double* __restrict__ mean_out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double s = 0.0;
  for (int64_t n = 0; n < N; ++n) s += Z[n * ldz + d];
  const double m = s / (double)N;
  mean_out[d] = m;
  for (int64_t n = 0; n < N; ++n) Z[n * ldz + d] -= m;
}

// E [D][k] row-major: divide every column by its 2-norm (useless/train.py:94-95). One CTA per column.
__global__ void colnorm_kernel(double* __restrict__ E, int D, int k) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  double s = 0.0;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double v = E[(int64_t)d * k + c];
    s = fma(v, v, s);
  }
  s = ef::warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
  const double nrm = sqrt(tot);
  for (int d = threadIdx.x; d < D; d += blockDim.x) E[(int64_t)d * k + c] = E[(int64_t)d * k + c] / nrm;
}

// Vt [k][D] row-major: normalise every row, then sklearn svd_flip(u_based_decision=False): make the entry of
// largest magnitude (first occurrence) positive.  sign[c] receives +-1.  One CTA per row.
__global__ void rownorm_flip_kernel(double* __restrict__ Vt, int k, int D, double* __restrict__ sign) {
  __shared__ double red[8];
  __shared__ double bestv[8];
  __shared__ int besti[8];
  const int c = blockIdx.x;
  double* row = Vt + (int64_t)c * D;
  double s = 0.0, bv = -1.0;
  int bi = INT_MAX;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const double v = row[d];
    s = fma(v, v, s);
    const double a = fabs(v);
    if (a > bv) { bv = a; bi = d; }
  }
  s = ef::warp_sum(s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    if (v2 > bv || (v2 == bv && i2 < bi)) { bv = v2; bi = i2; }
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[w] = s; bestv[w] = bv; besti[w] = bi; }
  __syncthreads();
  double tot = 0.0;
  bv = -1.0; bi = INT_MAX;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
    tot += red[i];
    if (bestv[i] > bv || (bestv[i] == bv && besti[i] < bi)) { bv = bestv[i]; bi = besti[i]; }
  }
  const double nrm = sqrt(tot);
  const double sg = (bi != INT_MAX && row[bi] < 0.0) ? -1.0 : 1.0;
  __syncthreads();
  const double f = (nrm > 0.0) ? sg / nrm : sg;
  for (int d = threadIdx.x; d < D; d += blockDim.x) row[d] = row[d] * f;
  if (threadIdx.x == 0) sign[c] = sg;
}

// features[n][c] = sign[c] * U[c][n] * S[c]   (U rows are eigenvectors of the N x N Gram)
__global__ void features_kernel(const double* __restrict__ U, int N, int k, const double* __restrict__ S,
                                const double* __restrict__ sign, double* __restrict__ F) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)N * k) return;
  const int n = (int)(e / k), c = (int)(e % k);
  F[e] = sign[c] * U[(int64_t)c * N + n] * S[c];
}

// E[d][c] = V[c][d] for c < k  (V rows of length D)
__global__ void rows_to_cols_kernel(const double* __restrict__ V, int D, int k, double* __restrict__ E) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)D * k) return;
  const int d = (int)(e / k), c = (int)(e % k);
  E[e] = V[(int64_t)c * D + d];
}

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  ~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
};

}  // namespace

extern "C" {

int ef_fit_gen1_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, double* eigenfaces,
                     double* mean, double* projected, double* eigenvalues, ef_fit_info_t* info) {
  if (!X || !eigenfaces || !mean || !projected || !eigenvalues || N < 2 || D <= 0 || ldx < D) return EF_ERR_INVALID;
  const bool snapshot = N < D;                                 // useless/train.py:82
  const int n = snapshot ? N : D;
  if (n > 4096) return EF_ERR_UNSUPPORTED;
  if (k <= 0) k = std::min(N - 1, D);                          // :111-112
  k = std::min(k, n);                                          // :114
  int dev_count = 0;
  EF_CUDA(cudaGetDeviceCount(&dev_count));
  cudaStream_t st = nullptr;
  EF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
  Timer tm;
  EF_CUDA(cudaEventCreate(&tm.a));
  EF_CUDA(cudaEventCreate(&tm.b));

  // workspaces are kept per host thread and only ever grow: repeated fits (one model per person) pay for cudaMalloc /
  // cudaFree once, not ~15 times per call
  static thread_local ef::DevBuf dX, dsum, dmean, dZ, dA, dwork, devals, devecs, dE, dP;
  static thread_local ef::DevBuf dG, dgw, dcw;
  static thread_local int ws_device = -1;
  int cur_device = 0;
  EF_CUDA(cudaGetDevice(&cur_device));
  if (cur_device != ws_device) {     // the cached buffers belong to another device: drop them
    for (ef::DevBuf* bptr : {&dX, &dsum, &dmean, &dZ, &dA, &dwork, &devals, &devecs, &dE, &dP, &dG, &dgw, &dcw}) bptr->release();
    ws_device = cur_device;
  }
  const int64_t ldxd = ef::round_up(D, 16);
  EF_TRY(dX.ensure((size_t)N * ldxd));
  EF_TRY(dsum.ensure(sizeof(int64_t) * D));
  EF_TRY(dmean.ensure(sizeof(double) * D));
  EF_TRY(dZ.ensure(sizeof(double) * (size_t)N * D));
  EF_TRY(dA.ensure(sizeof(double) * (size_t)n * n));
  EF_TRY(dwork.ensure(ef_eigh_work_bytes(n)));
  EF_TRY(devals.ensure(sizeof(double) * n));
  EF_TRY(devecs.ensure(sizeof(double) * (size_t)n * n));
  EF_TRY(dE.ensure(sizeof(double) * (size_t)D * k));
  EF_TRY(dP.ensure(sizeof(double) * (size_t)N * k));
  EF_TRY(dG.ensure(sizeof(int64_t) * (size_t)n * n));
  EF_TRY(dgw.ensure(ef_gram_u8_tc_work_bytes(N, D, snapshot ? 0 : 1)));
  EF_TRY(dcw.ensure(ef_gram_center_work_bytes(n) + 16));
  EF_CUDA(cudaMemcpy2DAsync(dX.p, ldxd, X, ldx, D, N, cudaMemcpyHostToDevice, st));
  EF_CUDA(cudaEventRecord(tm.a, st));

  // mean face, centred data
  EF_TRY(ef_colsum_u8_device(dX.as<uint8_t>(), ldxd, N, D, dsum.as<int64_t>(), st));
  EF_LAUNCH(mean_from_colsum_kernel, (unsigned)ef::ceil_div(D, 256), 256, 0, st, dsum.as<long long>(), D, (long long)N,
            dmean.as<double>());
  EF_TRY(ef_standardize_u8_device(dX.as<uint8_t>(), ldxd, N, D, dmean.as<double>(), nullptr, nullptr, dZ.as<double>(), D, st));
  const double alpha = 1.0 / (double)(N - 1);
  double* Z = dZ.as<double>();
  // cov = Xc Xc^T / (N-1)  (:84)  or  Xc^T Xc / (N-1)  (:99): exact integer Gram of the raw pixels on tensor cores,
  // centred on the small matrix with an exact integer numerator (one rounding per entry)
  const int side = snapshot ? 0 : 1;
  int gst = ef_gram_u8_tc_store_device(dX.as<uint8_t>(), ldxd, N, D, 0, D, side, dG.as<int64_t>(), dgw.p, dgw.bytes, st);
  if (gst == EF_ERR_UNSUPPORTED) {
    EF_CUDA(cudaMemsetAsync(dG.p, 0, sizeof(int64_t) * (size_t)n * n, st));
    gst = ef_gram_u8_device(dX.as<uint8_t>(), ldxd, N, D, 0, D, side, dG.as<int64_t>(), st);
  }
  EF_TRY(gst);
  EF_TRY(ef_gram_center_device(dG.as<int64_t>(), n, side, dsum.as<int64_t>(), N, alpha, dA.as<double>(), dcw.p, st));
  int sweeps = 0;
  double off = 0.0;
  const int est = ef_eigh_jacobi_device(dA.as<double>(), n, devals.as<double>(), devecs.as<double>(), dwork.p, 0, 0.0,
                                        &sweeps, &off, st);
  if (est != EF_OK) return est;
  double* E = dE.as<double>();
  if (snapshot) {
    // eigenfaces = Xc^T V (top-k columns), then column normalisation                          :91-95
    EF_TRY(ef_dgemm_device(D, k, N, 1.0, Z, 1, D, devecs.as<double>(), 1, N, 0.0, E, k, st));
    EF_LAUNCH(colnorm_kernel, k, 256, 0, st, E, D, k);
  } else {
    // eigenvectors of the covariance are the eigenfaces: E[d][c] = evecs[c][d]
    EF_LAUNCH(rows_to_cols_kernel, (unsigned)ef::ceil_div((int64_t)D * k, 256), 256, 0, st, devecs.as<double>(), D, k, E);
  }
  // projected_data = Xc E                                                                      :122
  EF_TRY(ef_dgemm_device(N, k, D, 1.0, Z, D, 1, E, k, 1, 0.0, dP.as<double>(), k, st));
  EF_CUDA(cudaEventRecord(tm.b, st));

  std::vector<double> Eh((size_t)D * k);
  EF_CUDA(cudaMemcpyAsync(Eh.data(), E, sizeof(double) * Eh.size(), cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(mean, dmean.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(projected, dP.p, sizeof(double) * (size_t)N * k, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(eigenvalues, devals.p, sizeof(double) * k, cudaMemcpyDeviceToHost, st));
  int32_t gram_flag = 0;
  EF_CUDA(cudaMemcpyAsync(&gram_flag, dgw.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaStreamSynchronize(st));
  if (gram_flag) {
    ef::set_error_detail("tcgen05 Gram pipeline timed out (mbarrier wait > 2 s)", cudaErrorLaunchTimeout);
    return EF_ERR_CUDA;
  }
  for (int c = 0; c < k; ++c)
    for (int d = 0; d < D; ++d) eigenfaces[(size_t)c * D + d] = Eh[(size_t)d * k + c];   // Fortran order [D][k]
  if (info) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tm.a, tm.b);
    info->sweeps = sweeps;
    info->branch = snapshot ? 0 : 1;
    info->off_norm = off;
    info->gpu_ms = ms;
  }
  return EF_OK;
}

// flavour 0: sklearn StandardScaler + PCA(full) (train-v5.py:349-385); flavour 1: ManualStandardScaler + ManualPCA
// (scripts/manual/train-v2.py:9-72: np.std with exact zeros -> 1; np.cov + eigh, whose top-k eigenvectors are the same
// directions as the SVD's -- the reference leaves their signs to LAPACK, here they follow the svd_flip rule).
// Zh != null: PCA only, of a float64 host matrix (what PCA.fit / ManualPCA.fit receive); the scaler outputs are unused.
static int fit_gen2_impl(const uint8_t* X, const double* Zh, int64_t ldx, int32_t N, int32_t D, int32_t k, int flavour,
                         const ef_gen2_fit_t* out, ef_fit_info_t* info) {
  if ((!X && !Zh) || !out || N < 2 || D <= 0 || ldx < D || k <= 0) return EF_ERR_INVALID;
  if (!out->pca_mean || !out->components || !out->explained_variance || !out->explained_variance_ratio ||
      !out->singular_values || !out->noise_variance || !out->features)
    return EF_ERR_INVALID;
  if (X && (!out->mean_face || !out->scaler_mean || !out->scaler_var || !out->scaler_scale)) return EF_ERR_INVALID;
  const bool snapshot = N <= D;
  const int n = snapshot ? N : D;
  if (n > 4096) return EF_ERR_UNSUPPORTED;
  if (k > n) return EF_ERR_INVALID;      // sklearn: n_components must be <= min(n_samples, n_features)
  cudaStream_t st = nullptr;
  EF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
  Timer tm;
  EF_CUDA(cudaEventCreate(&tm.a));
  EF_CUDA(cudaEventCreate(&tm.b));

  static thread_local ef::DevBuf dX, dsum, dmean, dvar, dscale, dpm, dZ, dA, dwork, devals, devecs, dVt, dS, dsign, dF,
      dcoef;
  static thread_local int ws_device = -1;
  int cur_device = 0;
  EF_CUDA(cudaGetDevice(&cur_device));
  if (cur_device != ws_device) {
    for (ef::DevBuf* bptr : {&dX, &dsum, &dmean, &dvar, &dscale, &dpm, &dZ, &dA, &dwork, &devals, &devecs, &dVt, &dS,
                             &dsign, &dF, &dcoef})
      bptr->release();
    ws_device = cur_device;
  }
  const int64_t ldxd = ef::round_up(D, 16);
  EF_TRY(dX.ensure((size_t)N * ldxd));
  EF_TRY(dsum.ensure(sizeof(int64_t) * D));
  EF_TRY(dmean.ensure(sizeof(double) * D));
  EF_TRY(dvar.ensure(sizeof(double) * D));
  EF_TRY(dscale.ensure(sizeof(double) * D));
  EF_TRY(dpm.ensure(sizeof(double) * D));
  EF_TRY(dZ.ensure(sizeof(double) * (size_t)N * D));
  EF_TRY(dA.ensure(sizeof(double) * (size_t)n * n));
  EF_TRY(dwork.ensure(ef_eigh_work_bytes(n)));
  EF_TRY(devals.ensure(sizeof(double) * n));
  EF_TRY(devecs.ensure(sizeof(double) * (size_t)n * n));
  EF_TRY(dVt.ensure(sizeof(double) * (size_t)k * D));
  EF_TRY(dS.ensure(sizeof(double) * n));
  EF_TRY(dsign.ensure(sizeof(double) * k));
  EF_TRY(dF.ensure(sizeof(double) * (size_t)N * k));
  double* Z = dZ.as<double>();
  if (X) {
    EF_CUDA(cudaMemcpy2DAsync(dX.p, ldxd, X, ldx, D, N, cudaMemcpyHostToDevice, st));
    EF_CUDA(cudaEventRecord(tm.a, st));
    // pixel mean (:366) == StandardScaler.mean_ ; var_, scale_ (:370)
    EF_TRY(ef_colsum_u8_device(dX.as<uint8_t>(), ldxd, N, D, dsum.as<int64_t>(), st));
    EF_LAUNCH(mean_from_colsum_kernel, (unsigned)ef::ceil_div(D, 256), 256, 0, st, dsum.as<long long>(), D, (long long)N,
              dmean.as<double>());
    if (flavour == 1) {
      EF_LAUNCH(scaler_var_manual_kernel, (unsigned)ef::ceil_div(D, 128), 128, 0, st, dX.as<uint8_t>(), ldxd, (int64_t)N, D,
                dmean.as<double>(), dvar.as<double>(), dscale.as<double>());
    } else {
      EF_LAUNCH(scaler_var_kernel, (unsigned)ef::ceil_div(D, 128), 128, 0, st, dX.as<uint8_t>(), ldxd, (int64_t)N, D,
                dmean.as<double>(), dvar.as<double>(), dscale.as<double>());
    }
    EF_TRY(ef_standardize_u8_device(dX.as<uint8_t>(), ldxd, N, D, dmean.as<double>(), dscale.as<double>(), nullptr, Z, D, st));
  } else {
    EF_CUDA(cudaMemcpy2DAsync(Z, sizeof(double) * D, Zh, sizeof(double) * ldx, sizeof(double) * D, N,
                              cudaMemcpyHostToDevice, st));
    EF_CUDA(cudaEventRecord(tm.a, st));
  }
  // PCA: centre (mean_ of the standardised data is ~1e-16 but sklearn subtracts it), :373
  EF_LAUNCH(colmean_center_kernel, (unsigned)ef::ceil_div(D, 128), 128, 0, st, Z, (int64_t)D, (int64_t)N, D,
            dpm.as<double>());
  if (snapshot) {
    EF_TRY(ef_dgemm_device(N, N, D, 1.0, Z, D, 1, Z, 1, D, 0.0, dA.as<double>(), N, st));      // Zc Zc^T = U S^2 U^T
  } else {
    EF_TRY(ef_dgemm_device(D, D, N, 1.0, Z, 1, D, Z, D, 1, 0.0, dA.as<double>(), D, st));      // Zc^T Zc = V S^2 V^T
  }
  int sweeps = 0;
  double off = 0.0;
  const int est = ef_eigh_jacobi_device(dA.as<double>(), n, devals.as<double>(), devecs.as<double>(), dwork.p, 0, 0.0,
                                        &sweeps, &off, st);
  if (est != EF_OK) return est;
  // O(n) host bookkeeping on the eigenvalues: S = sqrt(lambda), explained variance, ratio, noise variance
  std::vector<double> lam(n);
  EF_CUDA(cudaMemcpyAsync(lam.data(), devals.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaStreamSynchronize(st));
  std::vector<double> S(n), ev(n);
  double total = 0.0;
  for (int i = 0; i < n; ++i) {
    const double l = lam[i] > 0.0 ? lam[i] : 0.0;
    S[i] = std::sqrt(l);
    ev[i] = l / (double)(N - 1);
    total += ev[i];
  }
  EF_CUDA(cudaMemcpyAsync(dS.p, S.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
  double* Vt = dVt.as<double>();
  if (snapshot) {
    // Vt[c][:] = u_c^T Zc / S_c, evaluated as the normalised row (unit norm also when S_c ~ 0)
    EF_TRY(ef_dgemm_device(k, D, N, 1.0, devecs.as<double>(), N, 1, Z, D, 1, 0.0, Vt, D, st));
  } else {
    EF_CUDA(cudaMemcpyAsync(Vt, devecs.p, sizeof(double) * (size_t)k * D, cudaMemcpyDeviceToDevice, st));
  }
  EF_LAUNCH(rownorm_flip_kernel, k, 256, 0, st, Vt, k, D, dsign.as<double>());
  if (snapshot) {
    // Numerically null singular values (the centred matrix has rank <= N - 1, so with k = N the last one always is):
    // u^T Zc is rounding noise there, and its normalisation would be a direction INSIDE the row space.  LAPACK (and so
    // the reference's pickles) returns a unit vector orthogonal to all other components, for which Zc v = sigma u ~ 0:
    // the training crops then project to ~0 on it, matching the stored face_features.  Build exactly that: take the
    // noise row, project out every other component twice (classical Gram-Schmidt with re-orthogonalisation), renormalise.
    // (the spectrum comes from the Gram matrix, whose eigenvalues carry an absolute error ~ n eps lambda_1: singular
    // values below sqrt(n eps) sigma_1 are unresolved and count as null)
    const double null_tol = S[0] * std::sqrt((double)std::max(N, D) * 2.220446049250313e-16);
    for (int c = 0; c < k; ++c) {
      if (S[c] > null_tol) continue;
      EF_TRY(dcoef.ensure(sizeof(double) * k));
      double* w = Vt + (size_t)c * D;
      for (int pass = 0; pass < 2; ++pass) {
        // coef = Vt w (k values); the component's own coefficient is zeroed; w -= Vt^T coef
        EF_TRY(ef_dgemm_device(k, 1, D, 1.0, Vt, D, 1, w, 1, 1, 0.0, dcoef.as<double>(), 1, st));
        EF_CUDA(cudaMemsetAsync(dcoef.as<double>() + c, 0, sizeof(double), st));
        EF_TRY(ef_dgemm_device(D, 1, k, -1.0, Vt, 1, D, dcoef.as<double>(), 1, 1, 1.0, w, 1, st));
      }
      EF_LAUNCH(rownorm_flip_kernel, 1, 256, 0, st, w, 1, D, dsign.as<double>() + c);
    }
  }
  if (snapshot) {
    EF_LAUNCH(features_kernel, (unsigned)ef::ceil_div((int64_t)N * k, 256), 256, 0, st, devecs.as<double>(), N, k,
              dS.as<double>(), dsign.as<double>(), dF.as<double>());
  } else {
    // U S = Zc V
    EF_TRY(ef_dgemm_device(N, k, D, 1.0, Z, D, 1, Vt, 1, D, 0.0, dF.as<double>(), k, st));
  }
  EF_CUDA(cudaEventRecord(tm.b, st));

  if (X) {
    EF_CUDA(cudaMemcpyAsync(out->mean_face, dmean.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
    EF_CUDA(cudaMemcpyAsync(out->scaler_mean, dmean.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
    EF_CUDA(cudaMemcpyAsync(out->scaler_var, dvar.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
    EF_CUDA(cudaMemcpyAsync(out->scaler_scale, dscale.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  }
  EF_CUDA(cudaMemcpyAsync(out->pca_mean, dpm.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(out->components, Vt, sizeof(double) * (size_t)k * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(out->features, dF.p, sizeof(double) * (size_t)N * k, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaStreamSynchronize(st));
  for (int i = 0; i < k; ++i) {
    out->explained_variance[i] = ev[i];
    out->explained_variance_ratio[i] = total > 0.0 ? ev[i] / total : 0.0;
    out->singular_values[i] = S[i];
  }
  double noise = 0.0;
  if (k < std::min(N, D)) {
    // sklearn: mean of the discarded explained variances (over min(N, D) - k entries)
    const int m = std::min(N, D);
    for (int i = k; i < m; ++i) noise += (i < n ? ev[i] : 0.0);
    noise /= (double)(m - k);
  }
  *out->noise_variance = noise;
  if (info) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, tm.a, tm.b);
    info->sweeps = sweeps;
    info->branch = snapshot ? 0 : 1;
    info->off_norm = off;
    info->gpu_ms = ms;
  }
  return EF_OK;
}

int ef_fit_gen2_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                     ef_fit_info_t* info) {
  if (!X) return EF_ERR_INVALID;
  return fit_gen2_impl(X, nullptr, ldx, N, D, k, 0, out, info);
}

int ef_fit_manual_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                       ef_fit_info_t* info) {
  if (!X) return EF_ERR_INVALID;
  return fit_gen2_impl(X, nullptr, ldx, N, D, k, 1, out, info);
}

int ef_pca_fit_f64_host(const double* Z, int64_t ldz, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                        ef_fit_info_t* info) {
  if (!Z) return EF_ERR_INVALID;
  return fit_gen2_impl(nullptr, Z, ldz, N, D, k, 0, out, info);
}

int ef_scaler_fit_u8_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t flavour, double* mean, double* var,
                          double* scale) {
  if (!X || !mean || !var || !scale || N < 1 || D <= 0 || ldx < D || flavour < 0 || flavour > 1) return EF_ERR_INVALID;
  cudaStream_t st = nullptr;
  EF_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { cudaStreamDestroy(s); } } guard{st};
  ef::DevBuf dX, dsum, dmean, dvar, dscale;
  const int64_t ldxd = ef::round_up(D, 16);
  EF_TRY(dX.ensure((size_t)N * ldxd));
  EF_TRY(dsum.ensure(sizeof(int64_t) * D));
  EF_TRY(dmean.ensure(sizeof(double) * D));
  EF_TRY(dvar.ensure(sizeof(double) * D));
  EF_TRY(dscale.ensure(sizeof(double) * D));
  EF_CUDA(cudaMemcpy2DAsync(dX.p, ldxd, X, ldx, D, N, cudaMemcpyHostToDevice, st));
  EF_TRY(ef_colsum_u8_device(dX.as<uint8_t>(), ldxd, N, D, dsum.as<int64_t>(), st));
  EF_LAUNCH(mean_from_colsum_kernel, (unsigned)ef::ceil_div(D, 256), 256, 0, st, dsum.as<long long>(), D, (long long)N,
            dmean.as<double>());
  if (flavour == 1) {
    EF_LAUNCH(scaler_var_manual_kernel, (unsigned)ef::ceil_div(D, 128), 128, 0, st, dX.as<uint8_t>(), ldxd, (int64_t)N, D,
              dmean.as<double>(), dvar.as<double>(), dscale.as<double>());
  } else {
    EF_LAUNCH(scaler_var_kernel, (unsigned)ef::ceil_div(D, 128), 128, 0, st, dX.as<uint8_t>(), ldxd, (int64_t)N, D,
              dmean.as<double>(), dvar.as<double>(), dscale.as<double>());
  }
  EF_CUDA(cudaMemcpyAsync(mean, dmean.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(var, dvar.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(scale, dscale.p, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaStreamSynchronize(st));
  return EF_OK;
}

}  // extern "C"
