// K2a: exact-integer projection of uint8 crops on a digit-sliced basis (CUDA-core dp4a path).
//
// Replaces np.dot(face - mean_face, eigenfaces) (useless/scan.py:93-96) and scaler.transform +
// pca.transform (scan-template-v4.py:265-266) for a batch.
//
// The float64 basis W_eff[d][c] (= eigenfaces, or components_/scale_ for the sklearn models) is split on the
// host into S signed 7-bit digit planes per column (see ef_model.cu):
//     W~[d][c] = 2^e_c * sum_s q_s[d][c] * 2^-(7 s + 6),   q_s in [-64, 64],  |W - W~| <= 2^(e_c - 7 S)
// so that  x . W~[:,c]  =  2^e_c * sum_s 2^-(7s+6) * (x . q_s[:,c])  where every (x . q_s) is an EXACT int32
// dot product of uint8 pixels with int8 digits.  Integer accumulation is order independent, hence split-K with
// atomics and the multi-GPU shards are bit-reproducible, and with S = 8 the result carries 56 bits of the
// basis -- float64-equivalent -- which is what makes bit-exact identity labels against the float64 reference
// possible.  The tensor-core kernel (ef_project_tc.cu) computes the very same integers with tcgen05 kind::i8.
#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

constexpr int BM = 128;      // crops per CTA tile
constexpr int BK = 128;      // bytes of K per stage
constexpr int LDS_ROW = 144; // padded shared-memory row (bytes): conflict-free LDS.128 for 8 consecutive rows
constexpr int kThreads = 256;
constexpr int kStages = 3;

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// acc[b][n] += sum_{d in this CTA's K range} X[b][d] * Wq[n][d]
template <int CN>
__global__ void __launch_bounds__(kThreads)
project_dp4a_kernel(const uint8_t* __restrict__ X, int64_t ldx, int B, int D, const int8_t* __restrict__ Wq,
                    int64_t ldw, int NC, int k_tiles_per_split, int32_t* __restrict__ acc, int ld_acc) {
  constexpr int BN = 16 * CN;
  extern __shared__ __align__(16) uint8_t smem[];
  uint8_t* xs = smem;                                   // [kStages][BM][LDS_ROW]
  uint8_t* ws = smem + kStages * BM * LDS_ROW;          // [kStages][BN][LDS_ROW]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int total_k_tiles = (D + BK - 1) / BK;
  const int kt0 = blockIdx.z * k_tiles_per_split;
  const int kt1 = min(total_k_tiles, kt0 + k_tiles_per_split);
  const int nkt = kt1 - kt0;
  if (nkt <= 0) return;

  auto load_stage = [&](int stage, int kt) {
    const int kbase = kt * BK;
    uint8_t* xd = xs + stage * BM * LDS_ROW;
    uint8_t* wd = ws + stage * BN * LDS_ROW;
#pragma unroll
    for (int i = 0; i < (BM * (BK / 16)) / kThreads; ++i) {
      const int c = tid + i * kThreads;
      const int r = c >> 3, ck = c & 7;
      const int row = m0 + r, kk = kbase + ck * 16;
      int valid = (row < B) ? min(max(D - kk, 0), 16) : 0;
      const uint8_t* src = X + (int64_t)(row < B ? row : 0) * ldx + (valid > 0 ? kk : 0);
      cp_async16(xd + r * LDS_ROW + ck * 16, src, valid);
    }
    for (int c = tid; c < BN * (BK / 16); c += kThreads) {
      const int r = c >> 3, ck = c & 7;
      const int col = n0 + r, kk = kbase + ck * 16;
      const int valid = (col < NC && kk < ldw) ? 16 : 0;
      const int8_t* src = Wq + (int64_t)(col < NC ? col : 0) * ldw + (valid > 0 ? kk : 0);
      cp_async16(wd + r * LDS_ROW + ck * 16, src, valid);
    }
  };

  int accr[8][CN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) accr[i][j] = 0;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < nkt) load_stage(s, kt0 + s);
    cp_async_commit();
  }

  for (int it = 0; it < nkt; ++it) {
    cp_async_wait<kStages - 2>();
    __syncthreads();
    {
      const int nxt = it + kStages - 1;
      if (nxt < nkt) load_stage(nxt % kStages, kt0 + nxt);
      cp_async_commit();
    }
    const uint8_t* xd = xs + (it % kStages) * BM * LDS_ROW + (ty * 8) * LDS_ROW;
    const uint8_t* wd = ws + (it % kStages) * BN * LDS_ROW + tx * LDS_ROW;
#pragma unroll
    for (int kc = 0; kc < BK / 16; ++kc) {
      uint4 wv[CN];
#pragma unroll
      for (int j = 0; j < CN; ++j) wv[j] = *reinterpret_cast<const uint4*>(wd + (j * 16) * LDS_ROW + kc * 16);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xd + i * LDS_ROW + kc * 16);
#pragma unroll
        for (int j = 0; j < CN; ++j) {
          int a = accr[i][j];
          a = dp4a_us(xv.x, (int)wv[j].x, a);
          a = dp4a_us(xv.y, (int)wv[j].y, a);
          a = dp4a_us(xv.z, (int)wv[j].z, a);
          a = dp4a_us(xv.w, (int)wv[j].w, a);
          accr[i][j] = a;
        }
      }
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= B) continue;
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      const int col = n0 + j * 16 + tx;
      if (col < NC && accr[i][j] != 0) atomicAdd(acc + (int64_t)col * ld_acc + row, accr[i][j]);   // plane-major
    }
  }
}

// Combine the digit planes: P[b][c] = 2^e_c * sum_s acc[b][s*kq + c] * 2^-(7s+6) - bias[c]   (float64)
// and, when the residual column is present (column k of every plane),
//   resid2[b] = sumsq[b] - 2 * (x . u~) + c0 - |P_b|^2
// (the combination kernels live in ef_epilogue.cu)

// sumsq[b] = sum_d x^2 * qq[d]  (qq == NULL: exact integer sum of squares)
__global__ void rowsumsq_kernel(const uint8_t* __restrict__ X, int64_t ldx, int B, int D,
                                const double* __restrict__ qq, double* __restrict__ out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const uint8_t* x = X + (int64_t)warp * ldx;
  if (!qq) {
    unsigned long long s = 0;
    const int nv = ((ldx & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) ? D / 16 : 0;
    for (int i = lane; i < nv; i += 32) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
      unsigned t = __dp4a(v.x, v.x, 0u);
      t = __dp4a(v.y, v.y, t);
      t = __dp4a(v.z, v.z, t);
      t = __dp4a(v.w, v.w, t);
      s += t;
    }
    for (int d = nv * 16 + lane; d < D; d += 32) s += (unsigned)x[d] * (unsigned)x[d];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[warp] = (double)s;
  } else {
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
      const double v = (double)x[d];
      s += v * v * qq[d];
    }
    s = ef::warp_sum(s);
    if (lane == 0) out[warp] = s;
  }
}

// Weighted form for aligned rows with D % 16 == 0 (the standardised 64 x 64 models).  The warp-per-row kernel above is
// latency bound there (one dependent chain of D / 32 steps per warp, 22 us for 4096 x 4096).  Here a thread owns 16
// consecutive pixels of a 4096-pixel strip with their 16 weights in registers and the CTA walks over kRowsW rows, eight
// at a time: the eight 16-byte crop loads of a thread are independent and in flight together, x^2 is formed in integers
// (exact) and enters one fma per pixel (two interleaved chains), the eight warp reductions interleave.  Fixed order:
// warp tree, strips in sequence, warps in sequence -- every path of the library gets its weighted sum of squares here.
constexpr int kRowsW = 16;
// exact uint32 -> double without the conversion pipe: 2^52 + v has v in its low mantissa word
__device__ __forceinline__ double u32_to_double(unsigned v) {
  return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0;
}
__global__ void __launch_bounds__(256, 2)
rowsumsq_weighted_kernel(const uint8_t* __restrict__ X, int64_t ldx, int B, int D, const double* __restrict__ qq,
                         double* __restrict__ out) {
  __shared__ double part[kRowsW][8];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // finalize_slabs_kernel may be scheduled behind it
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * kRowsW;
  for (int base = 0; base < D; base += 4096) {
    const int d0 = base + 16 * tid;
    const bool live = d0 < D;                                   // D % 16 == 0: a live thread has all 16 pixels
    double w[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double2 t = live ? __ldg(reinterpret_cast<const double2*>(qq + d0) + j) : make_double2(0.0, 0.0);
      w[2 * j] = t.x;
      w[2 * j + 1] = t.y;
    }
#pragma unroll
    for (int r0 = 0; r0 < kRowsW; r0 += 8) {
      uint4 v[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {                              // rows past the batch re-read the last row (discarded)
        const int row = min(row0 + r0 + r, B - 1);
        v[r] = live ? __ldg(reinterpret_cast<const uint4*>(X + (int64_t)row * ldx + d0)) : make_uint4(0u, 0u, 0u, 0u);
      }
      double s[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const unsigned word[4] = {v[r].x, v[r].y, v[r].z, v[r].w};
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned b0 = word[j] & 255u, b1 = (word[j] >> 8) & 255u, b2 = (word[j] >> 16) & 255u, b3 = word[j] >> 24;
          s0 = fma(u32_to_double(b0 * b0), w[4 * j], s0);
          s1 = fma(u32_to_double(b1 * b1), w[4 * j + 1], s1);
          s0 = fma(u32_to_double(b2 * b2), w[4 * j + 2], s0);
          s1 = fma(u32_to_double(b3 * b3), w[4 * j + 3], s1);
        }
        s[r] = s0 + s1;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int r = 0; r < 8; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) part[r0 + r][warp] = base == 0 ? s[r] : part[r0 + r][warp] + s[r];
      }
    }
  }
  __syncthreads();
  if (tid < kRowsW && row0 + tid < B) {
    double s = part[tid][0];
#pragma unroll
    for (int w8 = 1; w8 < 8; ++w8) s += part[tid][w8];
    out[row0 + tid] = s;
  }
  // Launched with programmatic stream serialization behind the projection (row_sumsq(..., overlap)) this grid RUNS
  // CONCURRENTLY with it -- both only read the crops -- and waits for it here, at its end: a later kernel that waits
  // for this grid then also has the projection's results.  (A no-op for a normal launch.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <int CN>
int launch_dp4a(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int32_t* acc,
                int ld_acc, cudaStream_t stream) {
  constexpr int BN = 16 * CN;
  const size_t smem = (size_t)kStages * (BM + BN) * LDS_ROW;
  EF_ENSURE_SMEM(project_dp4a_kernel<CN>, smem);
  const int mt = (B + BM - 1) / BM, nt = (NC + BN - 1) / BN;
  const int total_k_tiles = (D + BK - 1) / BK;
  int ksplit = (int)ef::ceil_div(2 * (int64_t)ef::sm_count(), (int64_t)mt * nt);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > total_k_tiles) ksplit = total_k_tiles;
  const int per = (int)ef::ceil_div(total_k_tiles, ksplit);
  ksplit = (int)ef::ceil_div(total_k_tiles, per);
  dim3 grid(mt, nt, ksplit);
  EF_LAUNCH(project_dp4a_kernel<CN>, grid, kThreads, smem, stream, X, ldx, B, D, Wq, ldw, NC, per, acc, ld_acc);
  return EF_OK;
}

}  // namespace

namespace ef {

int project_dp4a(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int32_t* acc,
                 int ld_acc, cudaStream_t stream) {
  if (B <= 0) return EF_OK;   // acc (plane-major [NC][ld_acc]) must be zero on entry: the epilogue kernels clear it
  const int cn = (int)std::min<int64_t>(8, ceil_div(NC, 16));
  switch (cn) {
    case 1: return launch_dp4a<1>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 2: return launch_dp4a<2>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 3: return launch_dp4a<3>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 4: return launch_dp4a<4>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 5: return launch_dp4a<5>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 6: return launch_dp4a<6>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    case 7: return launch_dp4a<7>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
    default: return launch_dp4a<8>(X, ldx, B, D, Wq, ldw, NC, acc, ld_acc, stream);
  }
}

int row_sumsq(const uint8_t* X, int64_t ldx, int B, int D, const double* qq, double* out, cudaStream_t stream,
              bool overlap) {
  if (B <= 0) return EF_OK;
  const int threads = 256;
  const int grid = (int)ceil_div((int64_t)B * 32, threads);
  if (qq && D % 16 == 0 && (ldx & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(qq) & 15) == 0) {
    if (overlap)
      EF_LAUNCH_PDL(rowsumsq_weighted_kernel, (unsigned)ceil_div(B, kRowsW), 256, 0, stream, X, ldx, B, D, qq, out);
    else
      EF_LAUNCH(rowsumsq_weighted_kernel, (unsigned)ceil_div(B, kRowsW), 256, 0, stream, X, ldx, B, D, qq, out);
    return EF_OK;
  }
  EF_LAUNCH(rowsumsq_kernel, grid, threads, 0, stream, X, ldx, B, D, qq, out);
  return EF_OK;
}

}  // namespace ef
