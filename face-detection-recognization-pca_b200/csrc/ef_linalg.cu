// Float64 / exact-integer linear-algebra building blocks of the PCA fit.
//
//   ef_standardize_u8_device   X -> (X - mean) / scale - shift         (useless/train.py:74, StandardScaler.transform)
//   ef_dgemm_device            strided float64 GEMM                     (np.dot at useless/train.py:84,91,122)
//   ef_eigh_jacobi_device      one-sided Jacobi eigensolver             (np.linalg.eigh at useless/train.py:88,103)
//   ef_colsum_u8_device        exact column sums                        (np.mean at useless/train.py:70, train-v5.py:366)
//   ef_gram_u8_device          exact integer Gram X X^T / X^T X         (useless/train.py:84 / :99 before centring)
//   ef_gram_center_device      integer Gram -> centred float64 matrix
#include <climits>
#include <cstdlib>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

// ------------------------------------------------------------------------------------------ standardize
__global__ void standardize_kernel(const uint8_t* __restrict__ X, int64_t ldx, int64_t N, int D,
                                   const double* __restrict__ mean, const double* __restrict__ scale,
                                   const double* __restrict__ shift, double* __restrict__ Z, int64_t ldz) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const double m = mean ? mean[d] : 0.0;
  const double s = scale ? scale[d] : 1.0;
  const double t = shift ? shift[d] : 0.0;
  for (int64_t n = blockIdx.y; n < N; n += gridDim.y) {
    double v = (double)X[n * ldx + d] - m;   // X -= mean_
    if (scale) v = v / s;                    // X /= scale_   (IEEE division, second rounding like sklearn)
    if (shift) v = v - t;
    Z[n * ldz + d] = v;
  }
}

// ------------------------------------------------------------------------------------------------ dgemm
// C[m][n] = alpha * sum_k A(m,k) B(k,n) + beta * C[m][n]; T x T tile per CTA, (T/16)^2 outputs per thread.
template <int T>
__global__ void __launch_bounds__(256)
dgemm_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t sam, int64_t sak,
             const double* __restrict__ Bm, int64_t sbk, int64_t sbn, double beta, double* __restrict__ C,
             int64_t ldc) {
  constexpr int KC = 16;
  constexpr int R = T / 16;
  __shared__ double As[KC][T + 1];
  __shared__ double Bs[KC][T + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * T, n0 = blockIdx.x * T;
  double acc[R][R];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[i][j] = 0.0;
  const bool a_k_contig = (sak == 1);
  const bool b_n_contig = (sbn == 1);

  for (int k0 = 0; k0 < K; k0 += KC) {
    for (int e = tid; e < T * KC; e += 256) {
      int mm, kk;
      if (a_k_contig) { mm = e / KC; kk = e % KC; } else { mm = e % T; kk = e / T; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.0;
    }
    for (int e = tid; e < T * KC; e += 256) {
      int nn, kk;
      if (b_n_contig) { nn = e % T; kk = e / T; } else { nn = e / KC; kk = e % KC; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? Bm[(int64_t)k * sbk + (int64_t)n * sbn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      double av[R], bv[R];
#pragma unroll
      for (int i = 0; i < R; ++i) av[i] = As[kk][ty * R + i];
#pragma unroll
      for (int j = 0; j < R; ++j) bv[j] = Bs[kk][tx * R + j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty * R + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int n = n0 + tx * R + j;
      if (n >= N) continue;
      double v = alpha * acc[i][j];
      if (beta != 0.0) v += beta * C[(int64_t)m * ldc + n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// Large products (the covariance x block products of the subspace solver, back-projection at scale): 128 x 128 tile,
// 8 x 8 outputs per thread -- one shared-memory read per four DFMAs instead of one per two.  Every output still
// accumulates its K terms in ascending order with one fma each, so the result is bit identical to dgemm_kernel.
template <int RJ>
__global__ void __launch_bounds__(256, 1)
dgemm_big_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t sam, int64_t sak,
                 const double* __restrict__ Bm, int64_t sbk, int64_t sbn, double beta, double* __restrict__ C,
                 int64_t ldc) {
  constexpr int TM = 128, TN = 16 * RJ, KC = 8, R = 8;
  constexpr int LA = TM * KC / 256, LB = (TN * KC + 255) / 256;     // elements per thread and chunk
  __shared__ double As[KC][TM + 2];
  __shared__ double Bs[KC][TN + 2];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  double acc[R][RJ];
#pragma unroll
  for (int i = 0; i < R; ++i)
#pragma unroll
    for (int j = 0; j < RJ; ++j) acc[i][j] = 0.0;
  const bool a_k_contig = (sak == 1);
  const bool b_n_contig = (sbn == 1);
  double ra[LA], rb[LB];
  // the next K chunk is fetched into registers while the current one is multiplied (one memory latency per chunk
  // would otherwise be exposed: there is a single CTA of 8 warps per SM)
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      const int e = l * 256 + tid;
      int mm, kk;
      if (a_k_contig) { mm = e / KC; kk = e % KC; } else { mm = e % TM; kk = e / TM; }
      const int m = m0 + mm, k = k0 + kk;
      ra[l] = (m < M && k < K) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.0;
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      const int e = l * 256 + tid;
      int nn, kk;
      if (b_n_contig) { nn = e % TN; kk = e / TN; } else { nn = e / KC; kk = e % KC; }
      const int n = n0 + nn, k = k0 + kk;
      rb[l] = (e < TN * KC && n < N && k < K) ? Bm[(int64_t)k * sbk + (int64_t)n * sbn] : 0.0;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int l = 0; l < LA; ++l) {
      const int e = l * 256 + tid;
      int mm, kk;
      if (a_k_contig) { mm = e / KC; kk = e % KC; } else { mm = e % TM; kk = e / TM; }
      As[kk][mm] = ra[l];
    }
#pragma unroll
    for (int l = 0; l < LB; ++l) {
      const int e = l * 256 + tid;
      int nn, kk;
      if (b_n_contig) { nn = e % TN; kk = e / TN; } else { nn = e / KC; kk = e % KC; }
      if (e < TN * KC) Bs[kk][nn] = rb[l];
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += KC) {
    stash();
    __syncthreads();
    if (k0 + KC < K) fetch(k0 + KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      double av[R], bv[RJ];
      // rows ty + 16 i, columns tx + 16 j: consecutive lanes read consecutive shared-memory words (no bank conflicts)
#pragma unroll
      for (int i = 0; i < R; ++i) av[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < RJ; ++j) bv[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < RJ; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const int m = m0 + ty + 16 * i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < RJ; ++j) {
      const int n = n0 + tx + 16 * j;
      if (n >= N) continue;
      double v = alpha * acc[i][j];
      if (beta != 0.0) v += beta * C[(int64_t)m * ldc + n];
      C[(int64_t)m * ldc + n] = v;
    }
  }
}

// ----------------------------------------------------------------------------------------------- Jacobi
// One-sided (Hestenes) Jacobi on the symmetric matrix A: the rows of Bt start as the columns of A, the rows of
// Vt as the identity; plane rotations applied to row pairs of both drive the rows of Bt = (A V)^T to mutual
// orthogonality, at which point the rows of Vt are the eigenvectors and lambda_i = v_i . b_i.
// All n/2 disjoint pairs of a round (round-robin tournament) are rotated concurrently, one CTA per pair; a
// grid-wide barrier separates rounds (cooperative launch guarantees co-residency).  The matrices stay L2 resident
// (n <= 4096: 2 x 128 MB worst case, the shipped sizes 178..590 are 0.5-5.6 MB) and are accessed with .cg so that
// no stale L1 line is read after another SM rotated the row.
struct JacobiCtl {
  unsigned int barrier;        // monotonically increasing arrival counter
  unsigned int abort_flag;     // set when a spin timed out
  unsigned long long off_bits; // max |cos| of the current sweep (bits of a non-negative double)
  int sweeps;
  int converged;
  double off_final;
  double fro;                  // Frobenius norm of A (invariant under the rotations): the scale of the stopping test
};

__device__ __forceinline__ bool grid_barrier(JacobiCtl* ctl, unsigned int target) {
  __syncthreads();
  bool ok = true;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&ctl->barrier, 1u);
    long long spins = 0;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(&ctl->barrier));
      if (v >= target) break;
      if (++spins > (1ll << 24)) {
        unsigned int a;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(a) : "l"(&ctl->abort_flag));
        if (a || spins > (1ll << 28)) { atomicExch(&ctl->abort_flag, 1u); ok = false; break; }
      }
    }
    __threadfence();
  }
  ok = __syncthreads_and(ok);
  return ok;
}

__device__ __forceinline__ double block_sum3(double a, double b, double c, double* red, double& ob, double& oc) {
  // reduces three values across the CTA; result broadcast to all threads
  a = ef::warp_sum(a);
  b = ef::warp_sum(b);
  c = ef::warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) { red[warp * 3 + 0] = a; red[warp * 3 + 1] = b; red[warp * 3 + 2] = c; }
  __syncthreads();
  double sa = 0.0, sb = 0.0, sc = 0.0;
  for (int w = 0; w < nw; ++w) { sa += red[w * 3 + 0]; sb += red[w * 3 + 1]; sc += red[w * 3 + 2]; }
  ob = sb;
  oc = sc;
  return sa;
}

__global__ void __launch_bounds__(128)
jacobi_kernel(double* __restrict__ Bt, double* __restrict__ Vt, int n, int ld, JacobiCtl* ctl, int max_sweeps,
              double tol) {
  __shared__ double red[3 * 4];
  const int np = (n + 1) / 2;
  const int players = 2 * np;
  const int tid = threadIdx.x;
  unsigned int target = 0;
  const double rot_tol = tol * 0.25;
  const double fro = ctl->fro;

  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    double local_off = 0.0;
    for (int round = 0; round < players - 1; ++round) {
      for (int pair = blockIdx.x; pair < np; pair += gridDim.x) {
        int a, b;
        if (pair == 0) { a = players - 1; b = round; }
        else { a = (round + pair) % (players - 1); b = (round - pair + players - 1) % (players - 1); }
        if (a >= n || b >= n) continue;   // the dummy player of an odd n
        const int p = min(a, b), q = max(a, b);
        double* bp = Bt + (int64_t)p * ld;
        double* bq = Bt + (int64_t)q * ld;
        double alpha = 0.0, beta = 0.0, gamma = 0.0;
        for (int i = tid; i < n; i += blockDim.x) {
          const double x = __ldcg(bp + i), y = __ldcg(bq + i);
          alpha = fma(x, x, alpha);
          beta = fma(y, y, beta);
          gamma = fma(x, y, gamma);
        }
        alpha = block_sum3(alpha, beta, gamma, red, beta, gamma);
        // gamma = (V^T A^2 V)_pq ~ (lambda_p + lambda_q) E_pq with E = offdiag(V^T A V): the pair is converged when
        // |E_pq| <= tol * |A|_F, the absolute (norm-wise backward stable) criterion LAPACK's eigh meets too.
        // A relative test |gamma| <= tol sqrt(alpha beta) cannot be met for the numerically null directions of a
        // centred Gram matrix, whose columns of B = A V are rounding noise.
        const double denom = (sqrt(alpha) + sqrt(beta)) * fro;
        if (!(denom > 0.0) || gamma == 0.0) continue;
        const double off = fabs(gamma) / denom;
        local_off = fmax(local_off, off);
        if (off <= rot_tol) continue;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t);
        const double s = c * t;
        double* vp = Vt + (int64_t)p * ld;
        double* vq = Vt + (int64_t)q * ld;
        for (int i = tid; i < n; i += blockDim.x) {
          const double x = __ldcg(bp + i), y = __ldcg(bq + i);
          __stcg(bp + i, c * x - s * y);
          __stcg(bq + i, s * x + c * y);
          const double u = __ldcg(vp + i), v = __ldcg(vq + i);
          __stcg(vp + i, c * u - s * v);
          __stcg(vq + i, s * u + c * v);
        }
      }
      target += gridDim.x;
      if (!grid_barrier(ctl, target)) return;
    }
    // sweep-level convergence test (same decision in every CTA)
    if (tid == 0) atomicMax(&ctl->off_bits, (unsigned long long)__double_as_longlong(local_off));
    target += gridDim.x;
    if (!grid_barrier(ctl, target)) return;
    unsigned long long bits;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(bits) : "l"(&ctl->off_bits));
    const double off = __longlong_as_double((long long)bits);
    target += gridDim.x;
    if (!grid_barrier(ctl, target)) return;   // everyone has read off_bits before it is reset
    if (blockIdx.x == 0 && tid == 0) {
      ctl->sweeps = sweep + 1;
      ctl->off_final = off;
      ctl->converged = off <= tol;
      if (off > tol) atomicExch(&ctl->off_bits, 0ull);
    }
    if (off <= tol) return;
    target += gridDim.x;
    if (!grid_barrier(ctl, target)) return;   // reset visible before the next sweep's atomicMax
  }
}

// ------------------------------------------------------------------------------- Jacobi, cluster resident
// Same one-sided Jacobi for n <= 320, entirely ON CHIP: one cluster of 16 CTAs, one warp per pair, both players of a
// pair (row of B + row of V each) live in the REGISTERS of that warp (lane l holds elements l, l + 32, ...).  A round
// is: 3 warp-shuffle dot products, one rotation in registers, then the Brent-Luk shift -- every warp PUSHES its two
// players into the mailboxes of its neighbour warps (st.shared::cluster across CTA boundaries, plain shared stores
// inside a CTA), one hardware cluster barrier, and pulls its new players from its own mailbox.  No global memory and
// no grid-wide software barrier inside a sweep: a round costs ~1 us instead of ~4 us.
namespace jc {
constexpr int kClusterCtas = 16;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_remote_f64(uint32_t addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_remote_u32(uint32_t addr, unsigned v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_remote_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
}  // namespace jc

// Mailbox of one warp: [buffer 2][slot 2 (top, bottom)][B row | V row][EPL][32 lanes] doubles + the player ids.
template <int EPL>
struct JcMailbox {
  double data[2][2][2][EPL][32];
  int id[2][2];
  int pad[4];
};

template <int EPL>
__global__ void __launch_bounds__(320, 1)
jacobi_cluster_kernel(double* __restrict__ Bt, double* __restrict__ Vt, int n, int ld, JacobiCtl* ctl, int max_sweeps,
                      double tol, int warps_per_cta) {
  extern __shared__ __align__(16) unsigned char jc_smem[];
  JcMailbox<EPL>* box = reinterpret_cast<JcMailbox<EPL>*>(jc_smem);
  unsigned long long* off_all = reinterpret_cast<unsigned long long*>(box + warps_per_cta);   // [16] per-CTA maxima
  unsigned long long* off_cta = off_all + jc::kClusterCtas;                                    // this CTA's running max
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rank = jc::cluster_rank();
  const int np = (n + 1) / 2;                      // pairs = active warps of the cluster
  const int g = (int)rank * warps_per_cta + warp;  // position of this warp in the Brent-Luk ring
  const bool active = g < np;
  const double fro = ctl->fro;
  const double fro2 = fro * fro, rot_tol2 = (tol * 0.25) * (tol * 0.25);

  // players: top = 2g, bottom = 2g + 1 (id >= n: the dummy player of an odd n, a zero row that is never rotated)
  int id_t = 2 * g, id_b = 2 * g + 1;
  double bt[EPL], vt[EPL], bb[EPL], vb[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = lane + 32 * e;
    const bool in = i < n;
    bt[e] = (active && in && id_t < n) ? Bt[(int64_t)id_t * ld + i] : 0.0;
    bb[e] = (active && in && id_b < n) ? Bt[(int64_t)id_b * ld + i] : 0.0;
    vt[e] = (active && in && i == id_t) ? 1.0 : 0.0;
    vb[e] = (active && in && i == id_b) ? 1.0 : 0.0;
  }
  if (threadIdx.x == 0) *off_cta = 0ull;
  __syncthreads();
  jc::cluster_barrier();                           // every CTA of the cluster runs: remote mailboxes exist

  // destinations of the Brent-Luk shift (constant over the whole run)
  //   top[0] stays; bottom[0] -> top[1]; top[i] -> top[i+1] (1 <= i <= np-2); top[np-1] -> bottom[np-1];
  //   bottom[i] -> bottom[i-1] (i >= 1)
  int dt_g, dt_slot, db_g, db_slot;
  if (g == 0) { dt_g = 0; dt_slot = 0; } else if (g == np - 1) { dt_g = g; dt_slot = 1; } else { dt_g = g + 1; dt_slot = 0; }
  if (g == 0) { db_g = np > 1 ? 1 : 0; db_slot = np > 1 ? 0 : 1; } else { db_g = g - 1; db_slot = 1; }
  const uint32_t box_local = jc::smem_addr(box);
  const uint32_t dt_base = jc::mapa(box_local + (uint32_t)((dt_g % warps_per_cta) * sizeof(JcMailbox<EPL>)),
                                    (uint32_t)(dt_g / warps_per_cta));
  const uint32_t db_base = jc::mapa(box_local + (uint32_t)((db_g % warps_per_cta) * sizeof(JcMailbox<EPL>)),
                                    (uint32_t)(db_g / warps_per_cta));
  constexpr uint32_t kSlotBytes = 2 * EPL * 32 * 8, kBufBytes = 2 * kSlotBytes;
  constexpr uint32_t kIdOff = 2 * kBufBytes;

  int sweeps_done = 0, converged = 0;
  double off_final = 0.0;
  unsigned step = 0;
  for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
    double local_off2 = 0.0;
    for (int round = 0; round < 2 * np - 1; ++round, ++step) {
      if (active) {
        if (id_t < n && id_b < n) {
          double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
          for (int e = 0; e < EPL; ++e) {
            alpha = fma(bt[e], bt[e], alpha);
            beta = fma(bb[e], bb[e], beta);
            gamma = fma(bt[e], bb[e], gamma);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
            beta += __shfl_xor_sync(0xffffffffu, beta, o);
            gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
          }
          // off = |gamma| / ((sqrt(alpha) + sqrt(beta)) |A|_F), tested and tracked in squared form (one square root)
          const double d2 = (alpha + beta + 2.0 * sqrt(alpha * beta)) * fro2;
          if (d2 > 0.0 && gamma != 0.0) {
            const double g2 = gamma * gamma;
            if (g2 > local_off2 * d2) local_off2 = g2 / d2;
            if (g2 > rot_tol2 * d2) {
              // the lower-numbered player takes the role of row p (same orientation as the global-memory kernel);
              // t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)) with zeta = (a_q - a_p) / (2 gamma), written with one
              // square root and one division: t = 2 gamma / (delta + sign(delta) sqrt(delta^2 + 4 gamma^2))
              const bool t_first = id_t < id_b;
              const double delta = t_first ? beta - alpha : alpha - beta;
              const double root = sqrt(fma(delta, delta, 4.0 * g2));
              const double t = (delta == 0.0) ? copysign(1.0, gamma) : (2.0 * gamma) / (delta + copysign(root, delta));
              const double c = rsqrt(fma(t, t, 1.0));
              const double s = c * t;
#pragma unroll
              for (int e = 0; e < EPL; ++e) {
                const double xp = t_first ? bt[e] : bb[e], xq = t_first ? bb[e] : bt[e];
                const double np_ = c * xp - s * xq, nq_ = s * xp + c * xq;
                bt[e] = t_first ? np_ : nq_;
                bb[e] = t_first ? nq_ : np_;
                const double up = t_first ? vt[e] : vb[e], uq = t_first ? vb[e] : vt[e];
                const double mp_ = c * up - s * uq, mq_ = s * up + c * uq;
                vt[e] = t_first ? mp_ : mq_;
                vb[e] = t_first ? mq_ : mp_;
              }
            }
          }
        }
        // Brent-Luk shift: push both players to their next positions
        const uint32_t buf = step & 1u;
        const uint32_t pt = dt_base + buf * kBufBytes + (uint32_t)dt_slot * kSlotBytes + (uint32_t)lane * 8u;
        const uint32_t pb = db_base + buf * kBufBytes + (uint32_t)db_slot * kSlotBytes + (uint32_t)lane * 8u;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          jc::st_remote_f64(pt + (uint32_t)e * 256u, bt[e]);
          jc::st_remote_f64(pt + (uint32_t)(EPL + e) * 256u, vt[e]);
          jc::st_remote_f64(pb + (uint32_t)e * 256u, bb[e]);
          jc::st_remote_f64(pb + (uint32_t)(EPL + e) * 256u, vb[e]);
        }
        if (lane == 0) {
          jc::st_remote_u32(dt_base + kIdOff + (buf * 2u + (uint32_t)dt_slot) * 4u, (unsigned)id_t);
          jc::st_remote_u32(db_base + kIdOff + (buf * 2u + (uint32_t)db_slot) * 4u, (unsigned)id_b);
        }
      }
      jc::cluster_barrier();
      if (active) {
        const uint32_t buf = step & 1u;
        const JcMailbox<EPL>& mine = box[warp];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          bt[e] = mine.data[buf][0][0][e][lane];
          vt[e] = mine.data[buf][0][1][e][lane];
          bb[e] = mine.data[buf][1][0][e][lane];
          vb[e] = mine.data[buf][1][1][e][lane];
        }
        id_t = mine.id[buf][0];
        id_b = mine.id[buf][1];
      }
    }
    // sweep-level convergence test: per-CTA maximum, pushed to every CTA of the cluster
    if (lane == 0 && active) atomicMax(off_cta, (unsigned long long)__double_as_longlong(sqrt(local_off2)));
    __syncthreads();
    if (threadIdx.x < jc::kClusterCtas)
      jc::st_remote_u64(jc::mapa(jc::smem_addr(off_all + rank), threadIdx.x), *off_cta);
    jc::cluster_barrier();
    unsigned long long m = 0ull;
    for (int c = 0; c < jc::kClusterCtas; ++c) m = max(m, off_all[c]);
    const double off = __longlong_as_double((long long)m);
    sweeps_done = sweep + 1;
    off_final = off;
    converged = off <= tol;
    __syncthreads();
    if (threadIdx.x == 0) *off_cta = 0ull;
    jc::cluster_barrier();                         // everybody has read off_all before the next sweep overwrites it
  }
  // results: row `player id` of Bt / Vt
  if (active) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int i = lane + 32 * e;
      if (i < n) {
        if (id_t < n) { Bt[(int64_t)id_t * ld + i] = bt[e]; Vt[(int64_t)id_t * ld + i] = vt[e]; }
        if (id_b < n) { Bt[(int64_t)id_b * ld + i] = bb[e]; Vt[(int64_t)id_b * ld + i] = vb[e]; }
      }
    }
  }
  if (rank == 0 && threadIdx.x == 0) {
    ctl->sweeps = sweeps_done;
    ctl->off_final = off_final;
    ctl->converged = converged;
  }
}

template <int EPL>
int launch_jacobi_cluster(double* A, double* Vt, int n, JacobiCtl* ctl, int max_sweeps, double tol, cudaStream_t st) {
  const int np = (n + 1) / 2;
  const int wpc = (int)ef::ceil_div(np, jc::kClusterCtas);
  const size_t smem = (size_t)wpc * sizeof(JcMailbox<EPL>) + sizeof(unsigned long long) * (jc::kClusterCtas + 2);
  if (wpc > 10 || smem > 227 * 1024) return EF_ERR_UNSUPPORTED;
  if (EF_FIRST_ON_DEVICE()) {
    EF_CUDA(cudaFuncSetAttribute(jacobi_cluster_kernel<EPL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    EF_CUDA(cudaFuncSetAttribute(jacobi_cluster_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(jc::kClusterCtas);
  cfg.blockDim = dim3((unsigned)wpc * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = jc::kClusterCtas;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  // the 16-CTA cluster needs 16 free SMs of one GPC partition; if the device cannot place it, the caller falls back
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, jacobi_cluster_kernel<EPL>, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return EF_ERR_UNSUPPORTED;
  }
  int ld = n;
  EF_CUDA(cudaLaunchKernelEx(&cfg, jacobi_cluster_kernel<EPL>, A, Vt, n, ld, ctl, max_sweeps, tol, wpc));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  return EF_OK;
}

// EF_ERR_UNSUPPORTED when n is outside the on-chip kernel's coverage (the global-memory kernel takes over)
int jacobi_cluster(double* A, double* Vt, int n, JacobiCtl* ctl, int max_sweeps, double tol, cudaStream_t st) {
  if (n < 2 || n > 320 || getenv("EF_NO_CLUSTER_JACOBI")) return EF_ERR_UNSUPPORTED;
  const int epl = (n + 31) / 32;
  if (epl <= 2) return launch_jacobi_cluster<2>(A, Vt, n, ctl, max_sweeps, tol, st);
  if (epl <= 4) return launch_jacobi_cluster<4>(A, Vt, n, ctl, max_sweeps, tol, st);
  if (epl <= 6) return launch_jacobi_cluster<6>(A, Vt, n, ctl, max_sweeps, tol, st);
  if (epl <= 8) return launch_jacobi_cluster<8>(A, Vt, n, ctl, max_sweeps, tol, st);
  return launch_jacobi_cluster<10>(A, Vt, n, ctl, max_sweeps, tol, st);
}

// ------------------------------------------------------------------------------------- blocked cluster Jacobi
// 320 < n <= 640 (the 512-crop dark model, the 590-face Gen-2 model): the players no longer fit one 16-CTA cluster (two
// rows of n doubles per player in registers), and the grid-barrier kernel pays 4 us per round (58 ms at n = 512).  The
// players are cut into four groups; a sweep is three CROSS launches (two clusters each, one group pair per cluster:
// group a's players stay on top, group b's move one pair further every round -- after s rounds every cross pair has
// met exactly once) and one INTRA launch (four clusters, one Brent-Luk tournament per group).  Every pair of the sweep
// is rotated exactly once, as in the one-cluster kernel, with the same rotation arithmetic; rounds are separated by the
// hardware cluster barrier (~1.5 us), launches by the kernel boundary; rows travel through L2 between launches.
template <int EPL>
struct JbMailbox {                       // cross mode: only the bottom player moves
  double data[2][2][EPL][32];            // [buffer][B row | V row]
  int id[2];
  int pad[2];
};

struct JbGroups {
  int ga[8], gb[8];                      // per cluster: the group pair (cross) or the group (intra, gb unused)
};

template <int EPL>
__device__ __forceinline__ void jb_rotate(double (&bt)[EPL], double (&vt)[EPL], double (&bb)[EPL], double (&vb)[EPL],
                                          int id_t, int id_b, double fro2, double rot_tol2, double& local_off2) {
  double alpha = 0.0, beta = 0.0, gamma = 0.0;
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    alpha = fma(bt[e], bt[e], alpha);
    beta = fma(bb[e], bb[e], beta);
    gamma = fma(bt[e], bb[e], gamma);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
    beta += __shfl_xor_sync(0xffffffffu, beta, o);
    gamma += __shfl_xor_sync(0xffffffffu, gamma, o);
  }
  const double d2 = (alpha + beta + 2.0 * sqrt(alpha * beta)) * fro2;
  if (d2 > 0.0 && gamma != 0.0) {
    const double g2 = gamma * gamma;
    if (g2 > local_off2 * d2) local_off2 = g2 / d2;
    if (g2 > rot_tol2 * d2) {
      const bool t_first = id_t < id_b;            // same orientation as the other Jacobi kernels
      const double delta = t_first ? beta - alpha : alpha - beta;
      const double root = sqrt(fma(delta, delta, 4.0 * g2));
      const double t = (delta == 0.0) ? copysign(1.0, gamma) : (2.0 * gamma) / (delta + copysign(root, delta));
      const double c = rsqrt(fma(t, t, 1.0));
      const double sn = c * t;
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const double xp = t_first ? bt[e] : bb[e], xq = t_first ? bb[e] : bt[e];
        const double np_ = c * xp - sn * xq, nq_ = sn * xp + c * xq;
        bt[e] = t_first ? np_ : nq_;
        bb[e] = t_first ? nq_ : np_;
        const double up = t_first ? vt[e] : vb[e], uq = t_first ? vb[e] : vt[e];
        const double mp_ = c * up - sn * uq, mq_ = sn * up + c * uq;
        vt[e] = t_first ? mp_ : mq_;
        vb[e] = t_first ? mq_ : mp_;
      }
    }
  }
}

template <int EPL, int MODE>
__global__ void __launch_bounds__(320, 1)
jacobi_block_kernel(double* __restrict__ Bt, double* __restrict__ Vt, int n, int ld, JacobiCtl* ctl, double tol, int s,
                    JbGroups groups, int warps_per_cta) {
  extern __shared__ __align__(16) unsigned char jb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rank = jc::cluster_rank();
  const int cl = blockIdx.x / jc::kClusterCtas;
  const int ga = groups.ga[cl], gb = groups.gb[cl];
  const int np = MODE == 0 ? s : (s + 1) / 2;      // pairs of this cluster
  const int g = (int)rank * warps_per_cta + warp;
  const bool active = g < np;
  const double fro = ctl->fro;
  const double fro2 = fro * fro, rot_tol2 = (tol * 0.25) * (tol * 0.25);
  const int hi_a = min(n, (ga + 1) * s), hi_b = min(n, (gb + 1) * s);
  int id_t, id_b;
  if (MODE == 0) { id_t = ga * s + g; id_b = gb * s + g; if (id_t >= hi_a) id_t = INT_MAX; if (id_b >= hi_b) id_b = INT_MAX; }
  else { id_t = ga * s + 2 * g; id_b = ga * s + 2 * g + 1; if (id_t >= hi_a) id_t = INT_MAX; if (id_b >= hi_a) id_b = INT_MAX; }
  if (!active) id_t = id_b = INT_MAX;
  double bt[EPL], vt[EPL], bb[EPL], vb[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int i = lane + 32 * e;
    const bool in = i < n;
    bt[e] = (in && id_t != INT_MAX) ? Bt[(int64_t)id_t * ld + i] : 0.0;
    vt[e] = (in && id_t != INT_MAX) ? Vt[(int64_t)id_t * ld + i] : 0.0;
    bb[e] = (in && id_b != INT_MAX) ? Bt[(int64_t)id_b * ld + i] : 0.0;
    vb[e] = (in && id_b != INT_MAX) ? Vt[(int64_t)id_b * ld + i] : 0.0;
  }
  jc::cluster_barrier();                           // every CTA of the cluster runs: remote mailboxes exist
  double local_off2 = 0.0;
  unsigned step = 0;
  if (MODE == 0) {
    JbMailbox<EPL>* box = reinterpret_cast<JbMailbox<EPL>*>(jb_smem);
    const int gd = (g + 1) % np;                   // the bottom player moves one pair further
    const uint32_t d_base = jc::mapa(jc::smem_addr(box) + (uint32_t)((gd % warps_per_cta) * sizeof(JbMailbox<EPL>)),
                                     (uint32_t)(gd / warps_per_cta));
    constexpr uint32_t kBufBytes = 2 * EPL * 32 * 8, kIdOff = 2 * kBufBytes;
    for (int round = 0; round < np; ++round, ++step) {
      if (active) {
        if (id_t != INT_MAX && id_b != INT_MAX) jb_rotate<EPL>(bt, vt, bb, vb, id_t, id_b, fro2, rot_tol2, local_off2);
        const uint32_t buf = step & 1u;
        const uint32_t pb = d_base + buf * kBufBytes + (uint32_t)lane * 8u;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          jc::st_remote_f64(pb + (uint32_t)e * 256u, bb[e]);
          jc::st_remote_f64(pb + (uint32_t)(EPL + e) * 256u, vb[e]);
        }
        if (lane == 0) jc::st_remote_u32(d_base + kIdOff + buf * 4u, (unsigned)id_b);
      }
      jc::cluster_barrier();
      if (active) {
        const uint32_t buf = step & 1u;
        const JbMailbox<EPL>& mine = box[warp];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          bb[e] = mine.data[buf][0][e][lane];
          vb[e] = mine.data[buf][1][e][lane];
        }
        id_b = mine.id[buf];
      }
    }
  } else {
    JcMailbox<EPL>* box = reinterpret_cast<JcMailbox<EPL>*>(jb_smem);
    int dt_g, dt_slot, db_g, db_slot;              // Brent-Luk shift, as in jacobi_cluster_kernel
    if (g == 0) { dt_g = 0; dt_slot = 0; } else if (g == np - 1) { dt_g = g; dt_slot = 1; } else { dt_g = g + 1; dt_slot = 0; }
    if (g == 0) { db_g = np > 1 ? 1 : 0; db_slot = np > 1 ? 0 : 1; } else { db_g = g - 1; db_slot = 1; }
    const uint32_t box_local = jc::smem_addr(box);
    const uint32_t dt_base = jc::mapa(box_local + (uint32_t)((dt_g % warps_per_cta) * sizeof(JcMailbox<EPL>)),
                                      (uint32_t)(dt_g / warps_per_cta));
    const uint32_t db_base = jc::mapa(box_local + (uint32_t)((db_g % warps_per_cta) * sizeof(JcMailbox<EPL>)),
                                      (uint32_t)(db_g / warps_per_cta));
    constexpr uint32_t kSlotBytes = 2 * EPL * 32 * 8, kBufBytes = 2 * kSlotBytes, kIdOff = 2 * kBufBytes;
    for (int round = 0; round < 2 * np - 1; ++round, ++step) {
      if (active) {
        if (id_t != INT_MAX && id_b != INT_MAX) jb_rotate<EPL>(bt, vt, bb, vb, id_t, id_b, fro2, rot_tol2, local_off2);
        const uint32_t buf = step & 1u;
        const uint32_t pt = dt_base + buf * kBufBytes + (uint32_t)dt_slot * kSlotBytes + (uint32_t)lane * 8u;
        const uint32_t pb = db_base + buf * kBufBytes + (uint32_t)db_slot * kSlotBytes + (uint32_t)lane * 8u;
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          jc::st_remote_f64(pt + (uint32_t)e * 256u, bt[e]);
          jc::st_remote_f64(pt + (uint32_t)(EPL + e) * 256u, vt[e]);
          jc::st_remote_f64(pb + (uint32_t)e * 256u, bb[e]);
          jc::st_remote_f64(pb + (uint32_t)(EPL + e) * 256u, vb[e]);
        }
        if (lane == 0) {
          jc::st_remote_u32(dt_base + kIdOff + (buf * 2u + (uint32_t)dt_slot) * 4u, (unsigned)id_t);
          jc::st_remote_u32(db_base + kIdOff + (buf * 2u + (uint32_t)db_slot) * 4u, (unsigned)id_b);
        }
      }
      jc::cluster_barrier();
      if (active) {
        const uint32_t buf = step & 1u;
        const JcMailbox<EPL>& mine = box[warp];
#pragma unroll
        for (int e = 0; e < EPL; ++e) {
          bt[e] = mine.data[buf][0][0][e][lane];
          vt[e] = mine.data[buf][0][1][e][lane];
          bb[e] = mine.data[buf][1][0][e][lane];
          vb[e] = mine.data[buf][1][1][e][lane];
        }
        id_t = mine.id[buf][0];
        id_b = mine.id[buf][1];
      }
    }
  }
  if (active) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const int i = lane + 32 * e;
      if (i < n) {
        if (id_t != INT_MAX) { Bt[(int64_t)id_t * ld + i] = bt[e]; Vt[(int64_t)id_t * ld + i] = vt[e]; }
        if (id_b != INT_MAX) { Bt[(int64_t)id_b * ld + i] = bb[e]; Vt[(int64_t)id_b * ld + i] = vb[e]; }
      }
    }
    if (lane == 0 && local_off2 > 0.0)
      atomicMax(reinterpret_cast<unsigned long long*>(&ctl->off_bits),
                (unsigned long long)__double_as_longlong(sqrt(local_off2)));
  }
  jc::cluster_barrier();                           // no CTA leaves while a peer may still write into its mailboxes
}

template <int EPL, int MODE>
int launch_jacobi_block(double* A, double* Vt, int n, JacobiCtl* ctl, double tol, int s, const JbGroups& groups,
                        int clusters, cudaStream_t st) {
  const int np = MODE == 0 ? s : (s + 1) / 2;
  const int wpc = (int)ef::ceil_div(np, jc::kClusterCtas);
  const size_t smem = (size_t)wpc * (MODE == 0 ? sizeof(JbMailbox<EPL>) : sizeof(JcMailbox<EPL>));
  if (wpc > 10 || smem > 227 * 1024) return EF_ERR_UNSUPPORTED;
  if (EF_FIRST_ON_DEVICE()) {
    EF_CUDA(cudaFuncSetAttribute(jacobi_block_kernel<EPL, MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    EF_CUDA(cudaFuncSetAttribute(jacobi_block_kernel<EPL, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(clusters * jc::kClusterCtas));
  cfg.blockDim = dim3((unsigned)wpc * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = jc::kClusterCtas;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, jacobi_block_kernel<EPL, MODE>, &cfg) != cudaSuccess ||
      max_clusters < clusters) {
    cudaGetLastError();
    return EF_ERR_UNSUPPORTED;
  }
  int ld = n;
  EF_CUDA(cudaLaunchKernelEx(&cfg, jacobi_block_kernel<EPL, MODE>, A, Vt, n, ld, ctl, tol, s, groups, wpc));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  return EF_OK;
}

__global__ void jacobi_block_finish_kernel(JacobiCtl* ctl, int sweeps, int converged) {
  ctl->sweeps = sweeps;
  ctl->converged = converged;
  ctl->off_final = __longlong_as_double((long long)ctl->off_bits);
}

template <int EPL>
int jacobi_blocked_epl(double* A, double* Vt, int n, JacobiCtl* ctl, int max_sweeps, double tol, cudaStream_t st) {
  const int s = (n + 3) / 4;                        // four groups
  JbGroups cross[3] = {{{0, 2}, {1, 3}}, {{0, 1}, {2, 3}}, {{0, 1}, {3, 2}}};
  JbGroups intra = {{0, 1, 2, 3}, {0, 0, 0, 0}};
  int sweeps = 0, converged = 0;
  for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
    EF_CUDA(cudaMemsetAsync(&ctl->off_bits, 0, sizeof(unsigned long long), st));
    for (int r = 0; r < 3; ++r) EF_TRY((launch_jacobi_block<EPL, 0>(A, Vt, n, ctl, tol, s, cross[r], 2, st)));
    EF_TRY((launch_jacobi_block<EPL, 1>(A, Vt, n, ctl, tol, s, intra, 4, st)));
    unsigned long long bits = 0;
    EF_CUDA(cudaMemcpyAsync(&bits, &ctl->off_bits, sizeof(bits), cudaMemcpyDeviceToHost, st));
    EF_CUDA(cudaStreamSynchronize(st));
    double off;
    memcpy(&off, &bits, sizeof(off));
    sweeps = sweep + 1;
    converged = off <= tol ? 1 : 0;
  }
  EF_LAUNCH(jacobi_block_finish_kernel, 1, 1, 0, st, ctl, sweeps, converged);
  return EF_OK;
}

// EF_ERR_UNSUPPORTED outside 320 < n <= 640 or when the device cannot place four 16-CTA clusters
int jacobi_blocked(double* A, double* Vt, int n, JacobiCtl* ctl, int max_sweeps, double tol, cudaStream_t st) {
  if (n <= 320 || n > 640 || getenv("EF_NO_CLUSTER_JACOBI") || getenv("EF_NO_BLOCK_JACOBI")) return EF_ERR_UNSUPPORTED;
  const int epl = (n + 31) / 32;
  if (epl <= 12) return jacobi_blocked_epl<12>(A, Vt, n, ctl, max_sweeps, tol, st);
  if (epl <= 16) return jacobi_blocked_epl<16>(A, Vt, n, ctl, max_sweeps, tol, st);
  return jacobi_blocked_epl<20>(A, Vt, n, ctl, max_sweeps, tol, st);
}

// |A|_F with a fixed summation order (one CTA), so that the stopping threshold is bit-reproducible
__global__ void jacobi_fro_kernel(const double* __restrict__ A, int64_t count, JacobiCtl* ctl) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < count; i += blockDim.x) s = fma(A[i], A[i], s);
  s = ef::warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    ctl->fro = sqrt(t);
  }
}

__global__ void jacobi_init_kernel(double* __restrict__ Vt, int n, int ld, JacobiCtl* ctl) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) {
    ctl->barrier = 0;
    ctl->abort_flag = 0;
    ctl->off_bits = 0;
    ctl->sweeps = 0;
    ctl->converged = 0;
    ctl->off_final = 0.0;
  }
  if (e >= (int64_t)n * ld) return;
  const int r = (int)(e / ld), c = (int)(e % ld);
  Vt[e] = (r == c) ? 1.0 : 0.0;
}

// lambda_i = v_i . b_i ; one warp per row
__global__ void jacobi_rayleigh_kernel(const double* __restrict__ Bt, const double* __restrict__ Vt, int n, int ld,
                                       double* __restrict__ lam) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int i = lane; i < n; i += 32) s = fma(Vt[(int64_t)row * ld + i], Bt[(int64_t)row * ld + i], s);
  s = ef::warp_sum(s);
  if (lane == 0) lam[row] = s;
}

// rank by counting (descending, ties -> lower original index first) and scatter rows
__global__ void jacobi_sort_kernel(const double* __restrict__ lam, const double* __restrict__ Vt, int n, int ld,
                                   double* __restrict__ evals, double* __restrict__ evecs) {
  const int row = blockIdx.x;
  const double li = lam[row];
  __shared__ int rank_s;
  if (threadIdx.x == 0) rank_s = 0;
  __syncthreads();
  int cnt = 0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double lj = lam[j];
    if (lj > li || (lj == li && j < row)) ++cnt;
  }
  atomicAdd(&rank_s, cnt);
  __syncthreads();
  const int r = rank_s;
  if (threadIdx.x == 0) evals[r] = li;
  for (int i = threadIdx.x; i < n; i += blockDim.x) evecs[(int64_t)r * n + i] = Vt[(int64_t)row * ld + i];
}

// ------------------------------------------------------------------------------------- integer reductions
__global__ void colsum_u8_kernel(const uint8_t* __restrict__ X, int64_t ldx, int64_t N, int D,
                                 unsigned long long* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  unsigned long long s = 0;
  for (int64_t n = blockIdx.y; n < N; n += gridDim.y) s += X[n * ldx + d];
  if (gridDim.y == 1) out[d] = s; else atomicAdd(out + d, s);
}

// Exact integer Gram, 64 x 64 outputs per CTA, 4 x 4 per thread, dp4a on K-packed words.
// side 0: rows of X are the vectors (K = pixels, contiguous); side 1: columns of X are the vectors (K = rows,
// bytes transposed through shared memory).
__global__ void __launch_bounds__(256)
gram_u8_kernel(const uint8_t* __restrict__ X, int64_t ldx, int64_t N, int D, int d0, int d1, int side, int n_out,
               int64_t k_per_split, unsigned long long* __restrict__ G) {
  constexpr int T = 64, KB = 64, LDS = KB + 16;
  __shared__ __align__(16) uint8_t As[T][LDS];
  __shared__ __align__(16) uint8_t Bs[T][LDS];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * T, j0 = blockIdx.x * T;
  const int64_t K = side == 0 ? (int64_t)(d1 - d0) : N;
  const int64_t kb = (int64_t)blockIdx.z * k_per_split;
  const int64_t ke = min(K, kb + k_per_split);
  unsigned long long acc64[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc64[i][j] = 0;

  for (int64_t kc = kb; kc < ke; kc += 16384) {   // u32 partial sums stay below 2^32: 16384 * 255^2 < 2^31
    unsigned acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0;
    const int64_t kce = min(ke, kc + 16384);
    for (int64_t k0 = kc; k0 < kce; k0 += KB) {
      __syncthreads();
      for (int e = tid; e < T * KB; e += 256) {
        int v, kk;
        if (side == 0) { v = e / KB; kk = e % KB; } else { v = e % T; kk = e / T; }
        const int64_t k = k0 + kk;
        uint8_t a = 0, b = 0;
        if (k < kce) {
          if (side == 0) {
            if (i0 + v < n_out) a = X[(int64_t)(i0 + v) * ldx + d0 + k];
            if (j0 + v < n_out) b = X[(int64_t)(j0 + v) * ldx + d0 + k];
          } else {
            if (i0 + v < n_out) a = X[k * ldx + i0 + v];
            if (j0 + v < n_out) b = X[k * ldx + j0 + v];
          }
        }
        As[v][kk] = a;
        Bs[v][kk] = b;
      }
      __syncthreads();
#pragma unroll
      for (int kw = 0; kw < KB / 4; ++kw) {
        unsigned av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = *reinterpret_cast<const unsigned*>(&As[ty * 4 + i][kw * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const unsigned*>(&Bs[tx + 16 * j][kw * 4]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = __dp4a(av[i], bv[j], acc[i][j]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc64[i][j] += acc[i][j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i0 + ty * 4 + i;
    if (r >= n_out) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = j0 + tx + 16 * j;
      if (c < n_out) atomicAdd(G + (int64_t)r * n_out + c, acc64[i][j]);
    }
  }
}

__global__ void gram_rowsum_kernel(const long long* __restrict__ G, int n, long long* __restrict__ rowsum,
                                   unsigned long long* __restrict__ grand) {
  // one warp per row: exact int64 row sums (values < 2^63 for every supported shape)
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  long long s = 0;
  for (int j = lane; j < n; j += 32) s += G[(int64_t)row * n + j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    rowsum[row] = s;
    atomicAdd(grand, (unsigned long long)s);
  }
}

// Centring with an EXACT integer numerator and a single rounding:
//  side 0 (double centring of X X^T):  n^2 Gc[i][j] = n^2 G[i][j] - n (r_i + r_j) + g,  r = row sums, g = grand sum
//  side 1 (X^T X with column sums s): N Gc[a][b]   = N G[a][b] - s_a s_b
__global__ void gram_center_kernel(const long long* __restrict__ G, int n, int side,
                                   const long long* __restrict__ rowsum, const long long* __restrict__ grand_sum,
                                   const long long* __restrict__ colsum, long long Nrows, double alpha,
                                   double* __restrict__ C) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * n) return;
  const int i = (int)(e / n), j = (int)(e % n);
  double v;
  if (side == 0) {
    const long long nn = (long long)n;
    const long long num = nn * nn * G[e] - nn * (rowsum[i] + rowsum[j]) + *grand_sum;
    v = (double)num / ((double)nn * (double)nn);
  } else {
    const long long num = Nrows * G[e] - colsum[i] * colsum[j];
    v = (double)num / (double)Nrows;
  }
  C[e] = v * alpha;
}

}  // namespace

extern "C" {

int ef_standardize_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, const double* mean,
                             const double* scale, const double* shift, double* Z, int64_t ldz, ef_stream_t stream) {
  if (!X || !Z || N < 0 || D <= 0 || ldx < D || ldz < D) return EF_ERR_INVALID;
  if (N == 0) return EF_OK;
  dim3 grid((unsigned)ef::ceil_div(D, 256), (unsigned)std::min<int64_t>(N, 64));
  EF_LAUNCH(standardize_kernel, grid, 256, 0, ef::as_stream(stream), X, ldx, N, D, mean, scale, shift, Z, ldz);
  return EF_OK;
}

int ef_dgemm_device(int32_t M, int32_t N, int32_t K, double alpha, const double* A, int64_t sam, int64_t sak,
                    const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc,
                    ef_stream_t stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0 || ldc < N) return EF_ERR_INVALID;
  if (M == 0 || N == 0) return EF_OK;
  cudaStream_t st = ef::as_stream(stream);
  const int64_t tiles128 = ef::ceil_div(M, 128) * ef::ceil_div(N, 128);
  const int64_t tiles64 = ef::ceil_div(M, 64) * ef::ceil_div(N, 64);
  if (tiles128 >= ef::sm_count() && K >= 64) {
    // 128 x 128 or 128 x 96 tiles, whichever wastes fewer columns (N = 288 = 3 x 96)
    const int64_t waste128 = ef::round_up(N, 128) - N, waste96 = ef::round_up(N, 96) - N;
    if (waste96 < waste128) {
      dim3 grid((unsigned)ef::ceil_div(N, 96), (unsigned)ef::ceil_div(M, 128));
      EF_LAUNCH(dgemm_big_kernel<6>, grid, 256, 0, st, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
    } else {
      dim3 grid((unsigned)ef::ceil_div(N, 128), (unsigned)ef::ceil_div(M, 128));
      EF_LAUNCH(dgemm_big_kernel<8>, grid, 256, 0, st, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
    }
  } else if (tiles64 >= ef::sm_count()) {
    dim3 grid((unsigned)ef::ceil_div(N, 64), (unsigned)ef::ceil_div(M, 64));
    EF_LAUNCH(dgemm_kernel<64>, grid, 256, 0, st, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
  } else {
    dim3 grid((unsigned)ef::ceil_div(N, 32), (unsigned)ef::ceil_div(M, 32));
    EF_LAUNCH(dgemm_kernel<32>, grid, 256, 0, st, M, N, K, alpha, A, sam, sak, B, sbk, sbn, beta, C, ldc);
  }
  return EF_OK;
}

size_t ef_eigh_work_bytes(int32_t n) {
  if (n <= 0) return 0;
  // Vt [n][n] + lambda [n] + control block
  return sizeof(double) * ((size_t)n * n + (size_t)n) + 256;
}

int ef_eigh_jacobi_device(double* A, int32_t n, double* evals, double* evecs, void* work, int32_t max_sweeps,
                          double tol, int32_t* sweeps_used, double* off_norm, ef_stream_t stream) {
  if (!A || !evals || !evecs || !work || n <= 0) return EF_ERR_INVALID;
  if (n > 4096) return EF_ERR_UNSUPPORTED;
  cudaStream_t st = ef::as_stream(stream);
  if (max_sweeps <= 0) max_sweeps = 40;
  if (!(tol > 0.0)) tol = 1e-15;
  double* Vt = reinterpret_cast<double*>(work);
  double* lam = Vt + (size_t)n * n;
  JacobiCtl* ctl = reinterpret_cast<JacobiCtl*>(reinterpret_cast<char*>(work) +
                                                ef::round_up(sizeof(double) * ((size_t)n * n + n), 128));
  EF_LAUNCH(jacobi_init_kernel, (unsigned)ef::ceil_div((int64_t)n * n, 256), 256, 0, st, Vt, n, n, ctl);
  EF_LAUNCH(jacobi_fro_kernel, 1, 1024, 0, st, A, (int64_t)n * n, ctl);
  int st_cluster = EF_ERR_UNSUPPORTED;
  if (n > 1) st_cluster = jacobi_cluster(A, Vt, n, ctl, max_sweeps, tol, st);
  if (st_cluster == EF_ERR_UNSUPPORTED && n > 320) st_cluster = jacobi_blocked(A, Vt, n, ctl, max_sweeps, tol, st);
  if (st_cluster != EF_OK && st_cluster != EF_ERR_UNSUPPORTED) return st_cluster;
  if (n > 1 && st_cluster != EF_OK) {
    int per_sm = 0;
    EF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jacobi_kernel, 128, 0));
    if (per_sm < 1) return EF_ERR_UNSUPPORTED;
    const int np = (n + 1) / 2;
    int grid = std::min(np, per_sm * ef::sm_count());
    int ld = n;
    void* args[] = {&A, &Vt, &n, &ld, &ctl, &max_sweeps, &tol};
    EF_CUDA(cudaLaunchCooperativeKernel((const void*)jacobi_kernel, dim3(grid), dim3(128), args, 0, st));
    ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  }
  EF_LAUNCH(jacobi_rayleigh_kernel, (unsigned)ef::ceil_div((int64_t)n * 32, 256), 256, 0, st, A, Vt, n, n, lam);
  EF_LAUNCH(jacobi_sort_kernel, n, 128, 0, st, lam, Vt, n, n, evals, evecs);
  if (sweeps_used || off_norm) {
    JacobiCtl h;
    EF_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    EF_CUDA(cudaStreamSynchronize(st));
    if (sweeps_used) *sweeps_used = h.sweeps;
    if (off_norm) *off_norm = h.off_final;
    if (h.abort_flag) return EF_ERR_CUDA;
    if (n > 1 && !h.converged) return EF_ERR_NOCONVERGE;
  }
  return EF_OK;
}

int ef_colsum_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int64_t* out, ef_stream_t stream) {
  if (!X || !out || N < 0 || D <= 0 || ldx < D) return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  EF_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * (size_t)D, st));
  if (N == 0) return EF_OK;
  dim3 grid((unsigned)ef::ceil_div(D, 128), (unsigned)std::min<int64_t>(ef::ceil_div(N, 64), 64));
  EF_LAUNCH(colsum_u8_kernel, grid, 128, 0, st, X, ldx, N, D, reinterpret_cast<unsigned long long*>(out));
  return EF_OK;
}

int ef_gram_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                      int64_t* G, ef_stream_t stream) {
  if (!X || !G || N < 0 || D <= 0 || ldx < D || d0 < 0 || d1 > D || d0 > d1 || (side != 0 && side != 1))
    return EF_ERR_INVALID;
  const int64_t n_out = side == 0 ? N : D;
  const int64_t K = side == 0 ? (d1 - d0) : N;
  if (n_out == 0 || K == 0) return EF_OK;
  if (n_out > 65536) return EF_ERR_UNSUPPORTED;
  const int64_t tiles = ef::ceil_div(n_out, 64) * ef::ceil_div(n_out, 64);
  int64_t splits = ef::ceil_div(2 * (int64_t)ef::sm_count(), tiles);
  const int64_t max_splits = ef::ceil_div(K, 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int64_t per = ef::round_up(ef::ceil_div(K, splits), 64);
  splits = ef::ceil_div(K, per);
  dim3 grid((unsigned)ef::ceil_div(n_out, 64), (unsigned)ef::ceil_div(n_out, 64), (unsigned)splits);
  EF_LAUNCH(gram_u8_kernel, grid, 256, 0, ef::as_stream(stream), X, ldx, N, D, d0, d1, side, (int)n_out, per,
            reinterpret_cast<unsigned long long*>(G));
  return EF_OK;
}

int ef_gram_center_device(const int64_t* G, int32_t n, int32_t side, const int64_t* colsum, int64_t N, double alpha,
                          double* C, void* work, ef_stream_t stream) {
  if (!G || !C || n <= 0 || (side != 0 && side != 1)) return EF_ERR_INVALID;
  if (side == 1 && (!colsum || N <= 0)) return EF_ERR_INVALID;
  if (side == 0 && !work) return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  long long* rowsum = reinterpret_cast<long long*>(work);
  long long* grand = rowsum ? rowsum + n : nullptr;
  if (side == 0) {
    EF_CUDA(cudaMemsetAsync(grand, 0, sizeof(long long), st));
    EF_LAUNCH(gram_rowsum_kernel, (unsigned)ef::ceil_div((int64_t)n * 32, 256), 256, 0, st,
              reinterpret_cast<const long long*>(G), n, rowsum, reinterpret_cast<unsigned long long*>(grand));
  }
  EF_LAUNCH(gram_center_kernel, (unsigned)ef::ceil_div((int64_t)n * n, 256), 256, 0, st,
            reinterpret_cast<const long long*>(G), n, side, rowsum, grand, reinterpret_cast<const long long*>(colsum),
            (long long)N, alpha, C);
  return EF_OK;
}

size_t ef_gram_center_work_bytes(int32_t n) { return n > 0 ? sizeof(int64_t) * ((size_t)n + 2) : 0; }

}  // extern "C"
