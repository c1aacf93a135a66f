// K3 on tensor cores: the EXACT integer Gram matrix of uint8 vectors as a tcgen05 kind::i8 SYRK.
//
//   G[i][j] += sum_k A[i][k] * A[j][k]        A: uint8 [n][lda], K-major (one vector per row), G: int64 [n][ldg]
//
// This is the contraction behind manual_pca (useless/train.py:84 `np.dot(Xc, Xc.T)` for N < D, :99 `np.cov(Xc.T)` for
// N >= D) and sklearn's PCA on the same crops, taken BEFORE centring: pixels are 8-bit, so X X^T (or X^T X) is an exact
// integer, and centring is applied afterwards to the small matrix with an exact integer numerator
// (ef_gram_center_device).  Exact integers make the result independent of the K split, of the CTA schedule and of the
// number of GPUs the rows are sharded over (the all-reduce adds int64).
//
// Blackwell mapping (sm_100a):
//   * both operands are tiles of the SAME matrix, staged by TMA (SWIZZLE_128B) into a 4-stage shared-memory ring;
//     X X^T takes K-major tiles (a vector per row); X^T X takes the SAME row-major X as MN-major operands (TMA boxes
//     of 128 samples x 128 pixel bytes, UMMA descriptors with the transpose bits set) -- no transposed copy of X;
//     tcgen05.mma kind::i8 multiplies u8 x u8 into s32 TMEM accumulators, tile 128 x 256 (UMMA M = 128, N = 256);
//   * an s32 accumulator holds at most 256 K blocks (32768 x 255^2 < 2^31): longer K runs are cut into segments, each
//     flushed into the int64 result;
//   * two TMEM accumulator stages (2 x 256 columns): the four epilogue warps drain segment s while the MMA warp
//     already issues segment s + 1;
//   * only tiles that touch the upper triangle are computed (SYRK); a mirror kernel fills the lower triangle;
//   * schedule: many tiles (>= 2 per SM) -> whole tiles per CTA, exclusive ownership, plain coalesced int64
//     read-add-write; few tiles (the snapshot Gram of a few hundred crops) -> stream-K over all SMs with int64 RED
//     atomics (bit reproducible, integer addition is associative);
//   * the epilogue transposes each 32 x 32 block through shared memory so that a warp writes 256 contiguous bytes.
#include <cuda.h>

#include <cstdlib>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kThreads = 192;          // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue
constexpr int kSegKb = 256;            // K blocks (of 128 bytes) per s32 segment: 32768 * 255 * 255 < 2^31
constexpr int kAccStages = 2;

struct GramArgs {
  int n, kb_total, block_n, m_tiles, n_tiles, stages, whole, tmem_cols;
  int overwrite;                       // G = A A^T instead of +=: the first K segment of a tile stores (both triangles)
  int mn_major;                        // operand = X [K samples][n pixels] as it lies in memory (MN-major UMMA operands)
  long long tiles;                     // tiles touching the upper triangle
  unsigned long long* G;
  long long ldg;
  int* status;
};

struct GramShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long tmem_full_bar[kAccStages];
  unsigned long long tmem_empty_bar[kAccStages];
  uint32_t tmem_base;
  int failed;
};

// Instruction descriptor: D = s32, A = u8, B = u8, both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t umma_idesc_u8u8(int n) {
  return (2u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

// MN-major SWIZZLE_128B operand (the pixel axis contiguous, as X lies in memory): a TMA box of 128 samples x 128 pixel
// bytes is 16 swizzle atoms of (8 samples x 128 pixels), 1024 bytes apart along K (stride byte offset); a second block of
// 128 pixels (N = 256) starts `lbo` bytes further (leading byte offset).
__device__ __forceinline__ uint64_t umma_desc_sw128_mn(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Upper-triangle tiles in column-block-major order: column block tj holds row tiles 0 .. min(m_tiles, w tj + w) - 1
// with w = block_n / 128 rounded up (row tile ti intersects c >= r iff 128 ti <= block_n tj + block_n - 1).
__device__ __forceinline__ int rows_in_col(const GramArgs& a, int tj) {
  const long long last_col = (long long)a.block_n * tj + a.block_n - 1;
  const long long cnt = last_col / BLOCK_M + 1;
  return (int)(cnt < a.m_tiles ? cnt : a.m_tiles);
}
__device__ __forceinline__ void decode_tile(const GramArgs& a, long long t, int& ti, int& tj) {
  int j = 0;
  for (; j < a.n_tiles - 1; ++j) {
    const int c = rows_in_col(a, j);
    if (t < c) break;
    t -= c;
  }
  tj = j;
  ti = (int)t;
}

// The sequence of (tile, K-block range) segments of this CTA; identical in the three roles.
struct SegIter {
  long long u, u_end;          // stream-K: unit cursor over tiles * kb_total
  long long t, t_step;         // whole tiles: tile cursor
  int kb;
  __device__ SegIter(const GramArgs& a) {
    if (a.whole) {
      t = blockIdx.x; t_step = gridDim.x; kb = 0; u = u_end = 0;
    } else {
      const long long total = a.tiles * a.kb_total;
      u = total * blockIdx.x / gridDim.x;
      u_end = total * (blockIdx.x + 1) / gridDim.x;
      t = t_step = 0; kb = 0;
    }
  }
  __device__ bool next(const GramArgs& a, long long& tile, int& kb0, int& kb1) {
    if (a.whole) {
      if (t >= a.tiles) return false;
      tile = t; kb0 = kb;
      kb1 = min(a.kb_total, kb + kSegKb);
      kb = kb1;
      if (kb >= a.kb_total) { kb = 0; t += t_step; }
      return true;
    }
    if (u >= u_end) return false;
    tile = u / a.kb_total;
    kb0 = (int)(u % a.kb_total);
    const long long lim = min((long long)a.kb_total, kb0 + (u_end - u));
    kb1 = (int)min(lim, (long long)kb0 + kSegKb);
    u += kb1 - kb0;
    return true;
  }
};

__global__ void __launch_bounds__(kThreads, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GramArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  const int b_stage_bytes = a.block_n * BLOCK_K;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)a.stages * A_STAGE_BYTES;
  uint32_t* stage_out = reinterpret_cast<uint32_t*>(sB + (size_t)a.stages * b_stage_bytes);   // [4 warps][32][33]
  GramShared* sh = reinterpret_cast<GramShared*>(stage_out + 4 * 32 * 33);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&sh->tmem_full_bar[s], 1);
      mbar_init(&sh->tmem_empty_bar[s], 4);
    }
    sh->failed = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"((uint32_t)a.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_bytes = (uint32_t)(A_STAGE_BYTES + b_stage_bytes);
      SegIter it(a);
      long long tile;
      int kb0, kb1;
      bool ok = true;
      while (ok && it.next(a, tile, kb0, kb1)) {
        int ti, tj;
        decode_tile(a, tile, ti, tj);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->empty_bar[stage], phase ^ 1, failed)) { ok = false; break; }
          mbar_arrive_expect_tx(&sh->full_bar[stage], stage_bytes);
          if (a.mn_major) {
            // boxes of 128 samples x 128 pixel bytes straight from X: no transposed copy
            tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, &tmap_a, &sh->full_bar[stage], ti * BLOCK_M, kb * BLOCK_K);
            for (int h = 0; h < a.block_n / BLOCK_M; ++h)
              tma_load_2d(sB + (size_t)stage * b_stage_bytes + (size_t)h * A_STAGE_BYTES, &tmap_a, &sh->full_bar[stage],
                          tj * a.block_n + h * BLOCK_M, kb * BLOCK_K);
          } else {
            tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, &tmap_a, &sh->full_bar[stage], kb * BLOCK_K, ti * BLOCK_M);
            tma_load_2d(sB + (size_t)stage * b_stage_bytes, &tmap_b, &sh->full_bar[stage], kb * BLOCK_K, tj * a.block_n);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t seg = 0;
      const uint32_t idesc = umma_idesc_u8u8(a.block_n) | (a.mn_major ? ((1u << 15) | (1u << 16)) : 0u);
      SegIter it(a);
      long long tile;
      int kb0, kb1;
      bool ok = true;
      while (ok && it.next(a, tile, kb0, kb1)) {
        const uint32_t as = seg % kAccStages, use = seg / kAccStages;
        if (!mbar_wait(&sh->tmem_empty_bar[as], (use & 1) ^ 1, failed)) { ok = false; break; }
        tc_fence_after();
        const uint32_t d_addr = tmem_base + as * (uint32_t)a.block_n;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
          if (a.mn_major) {
            // K advances by 32 samples = 32 rows of 128 bytes per MMA
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_i8(d_addr, umma_desc_sw128_mn(a_addr + k * UMMA_K * 128, A_STAGE_BYTES),
                      umma_desc_sw128_mn(b_addr + k * UMMA_K * 128, A_STAGE_BYTES), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_i8(d_addr, umma_desc_sw128(a_addr + k * UMMA_K), umma_desc_sw128(b_addr + k * UMMA_K), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&sh->empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        umma_commit(&sh->tmem_full_bar[as]);
        ++seg;
      }
    }
  } else {
    // ===================================================================== epilogue warps
    const int lane_group = warp & 3;                     // TMEM lanes [32 g, 32 g + 32)
    uint32_t* my_stage = stage_out + (warp - 2) * 32 * 33;
    uint32_t seg = 0;
    SegIter it(a);
    long long tile;
    int kb0, kb1;
    bool ok = true;
    while (ok && it.next(a, tile, kb0, kb1)) {
      int ti, tj;
      decode_tile(a, tile, ti, tj);
      const uint32_t as = seg % kAccStages, use = seg / kAccStages;
      if (!mbar_wait(&sh->tmem_full_bar[as], use & 1, failed)) { ok = false; break; }
      tc_fence_after();
      const int r0 = ti * BLOCK_M + lane_group * 32;     // first row of this warp's 32-row slab
      for (int c0 = 0; c0 < a.block_n; c0 += 32) {
        const int cg0 = tj * a.block_n + c0;             // first global column of this 32-column chunk
        // warp-uniform skip: nothing of this 32 x 32 block lies in the upper triangle, or it is out of range
        const bool dead = (cg0 + 31 < r0) || (r0 >= a.n) || (cg0 >= a.n);
        if (dead) continue;
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(lane_group * 32) << 16) + as * (uint32_t)a.block_n + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) my_stage[lane * 33 + j] = v[j];       // row = lane, column = j
        __syncwarp();
        const int c = cg0 + lane;
        if (a.whole && a.overwrite && kb0 == 0) {
          // fresh result: plain stores, no read of G; the block goes out twice -- as it is (upper triangle, a warp writes
          // 256 contiguous bytes of a row) and transposed (lower triangle: row cg0 + j, columns r0 .. r0 + 31), so that
          // no mirror pass has to read the upper triangle back
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const int r = r0 + rr;
            if (c < a.n && r < a.n && c >= r) __stcg(a.G + (long long)r * a.ldg + c, (unsigned long long)my_stage[rr * 33 + lane]);
          }
          const int cm = r0 + lane;
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int rm = cg0 + j;
            if (rm < a.n && cm < a.n && rm > cm) __stcg(a.G + (long long)rm * a.ldg + cm, (unsigned long long)my_stage[lane * 33 + j]);
          }
        } else if (a.whole) {
          // exclusive owner of this tile: read-add-write, the 16 loads of a half block in flight together
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            unsigned long long old[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int r = r0 + half * 16 + q;
              const bool live = c < a.n && r < a.n && c >= r;
              old[q] = live ? __ldcg(a.G + (long long)r * a.ldg + c) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const int rr = half * 16 + q, r = r0 + rr;
              if (c < a.n && r < a.n && c >= r)
                __stcg(a.G + (long long)r * a.ldg + c, old[q] + (unsigned long long)my_stage[rr * 33 + lane]);
            }
          }
        } else {
          for (int rr = 0; rr < 32; ++rr) {
            const int r = r0 + rr;
            if (r >= a.n) break;
            const uint32_t val = my_stage[rr * 33 + lane];
            if (c < a.n && c >= r && val != 0u) atomicAdd(a.G + (long long)r * a.ldg + c, (unsigned long long)val);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty_bar[as]);
      ++seg;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
  if (threadIdx.x == 0 && sh->failed) atomicExch(a.status, 1);
}

// lower triangle := upper triangle (32 x 32 blocks through shared memory, both sides coalesced)
__global__ void gram_mirror_kernel(unsigned long long* __restrict__ G, int n, long long ldg) {
  __shared__ unsigned long long t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;            // block row / block column of the SOURCE (upper) block
  if (bj < bi) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
  for (int y = ty; y < 32; y += 8) {
    const int r = bi * 32 + y, c = bj * 32 + tx;
    t[y][tx] = (r < n && c < n) ? G[(long long)r * ldg + c] : 0ull;
  }
  __syncthreads();
  for (int y = ty; y < 32; y += 8) {
    const int r = bj * 32 + y, c = bi * 32 + tx;         // destination (r, c) = source (c, r)
    if (r < n && c < n && r > c) G[(long long)r * ldg + c] = t[tx][y];
  }
}

// out[c][r] = in[r][c] for uint8, 64 x 64 tiles
__global__ void transpose_u8_kernel(const uint8_t* __restrict__ in, long long ldi, long long rows, int cols,
                                    uint8_t* __restrict__ out, long long ldo) {
  __shared__ uint8_t t[64][65];
  const long long r0 = (long long)blockIdx.y * 64;
  const int c0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;    // 64 x 4 threads
  for (int y = ty; y < 64; y += 4) {
    const long long r = r0 + y;
    const int c = c0 + tx;
    t[y][tx] = (r < rows && c < cols) ? in[r * ldi + c] : (uint8_t)0;
  }
  __syncthreads();
  for (int y = ty; y < 64; y += 4) {
    const int c = c0 + y;
    const long long r = r0 + tx;
    if (c < cols && r < rows) out[(long long)c * ldo + r] = t[tx][y];
  }
}

}  // namespace

namespace ef {

using namespace ef_tc;

// G[n][ldg] (int64) += A A^T restricted to the upper triangle, then mirrored.  A: uint8 [n][lda] K-major with K valid
// bytes per row.  EF_ERR_UNSUPPORTED when the TMA alignment rules are not met.
int gram_tc(const uint8_t* A, int64_t lda, int64_t n, int64_t K, int64_t* G, int64_t ldg, int* status,
            cudaStream_t stream, bool overwrite, bool mn_major) {
  if (n <= 0 || K <= 0) return EF_OK;
  if (mn_major && n < 256) return EF_ERR_UNSUPPORTED;      // MN-major tiles are whole 128-pixel blocks (block_n = 256)
  if ((lda & 15) || (reinterpret_cast<uintptr_t>(A) & 15) || n > (1 << 20) || K > (1ll << 31) - 256)
    return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  GramArgs a{};
  a.n = (int)n;
  a.kb_total = (int)ceil_div(K, BLOCK_K);
  a.block_n = (int)std::min<int64_t>(256, round_up(n, 16));
  a.m_tiles = (int)ceil_div(n, BLOCK_M);
  a.n_tiles = (int)ceil_div(n, a.block_n);
  long long tiles = 0;
  for (int tj = 0; tj < a.n_tiles; ++tj) {
    const long long cnt = ((long long)a.block_n * tj + a.block_n - 1) / BLOCK_M + 1;
    tiles += std::min<long long>(cnt, a.m_tiles);
  }
  a.tiles = tiles;
  a.tmem_cols = 32;
  while (a.tmem_cols < kAccStages * a.block_n) a.tmem_cols *= 2;
  const int stage_bytes = A_STAGE_BYTES + a.block_n * BLOCK_K;
  const size_t fixed = 4 * 32 * 33 * sizeof(uint32_t) + sizeof(GramShared) + 64;
  a.stages = (int)std::min<size_t>(kMaxStages, ((size_t)kSmemLimit - fixed) / stage_bytes);
  if (a.stages < 2) return EF_ERR_UNSUPPORTED;
  const int sms = sm_count();
  a.whole = tiles >= 2ll * sms ? 1 : 0;
  a.overwrite = overwrite ? 1 : 0;
  a.mn_major = mn_major ? 1 : 0;
  if (overwrite && !a.whole) {
    // stream-K (small n) accumulates partial tiles with atomics: a fresh result starts from zero
    EF_CUDA(cudaMemsetAsync(G, 0, sizeof(int64_t) * (size_t)n * (size_t)ldg, stream));
    a.overwrite = 0;
  }
  a.G = reinterpret_cast<unsigned long long*>(G);
  a.ldg = ldg;
  a.status = status;
  int grid;
  if (a.whole) {
    grid = sms;
  } else {
    // stream-K: at least 8 K blocks per CTA so that the atomic flush of a partial tile stays a small share
    const long long units = tiles * a.kb_total;
    grid = (int)std::max<long long>(1, std::min<long long>(sms, units / 8));
  }
  CUtensorMap ma, mb;
  if (mn_major) {
    // A = X [K samples][lda], n pixels per row: one map, boxes of 128 pixel bytes x 128 samples
    if (!make_map(&ma, A, (uint64_t)n, (uint64_t)K, (uint64_t)lda, BLOCK_K)) return EF_ERR_UNSUPPORTED;
    mb = ma;
  } else {
    if (!make_map(&ma, A, (uint64_t)K, (uint64_t)n, (uint64_t)lda, BLOCK_M)) return EF_ERR_UNSUPPORTED;
    if (!make_map(&mb, A, (uint64_t)K, (uint64_t)n, (uint64_t)lda, (uint32_t)a.block_n)) return EF_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)a.stages * stage_bytes + fixed;
  EF_ENSURE_SMEM(gram_tc_kernel, smem);
  EF_LAUNCH(gram_tc_kernel, grid, kThreads, smem, stream, ma, mb, a);
  if (a.overwrite && a.kb_total <= kSegKb) return EF_OK;      // every tile had ONE segment: both triangles are written
  dim3 mg((unsigned)ceil_div(n, 32), (unsigned)ceil_div(n, 32));
  EF_LAUNCH(gram_mirror_kernel, mg, 256, 0, stream, a.G, a.n, (long long)ldg);
  return EF_OK;
}

int transpose_u8(const uint8_t* in, int64_t ldi, int64_t rows, int cols, uint8_t* out, int64_t ldo,
                 cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return EF_OK;
  dim3 grid((unsigned)ceil_div(cols, 64), (unsigned)ceil_div(rows, 64));
  if (grid.y > 65535) return EF_ERR_UNSUPPORTED;
  EF_LAUNCH(transpose_u8_kernel, grid, 256, 0, stream, in, (long long)ldi, (long long)rows, cols, out, (long long)ldo);
  return EF_OK;
}

}  // namespace ef

extern "C" {

size_t ef_gram_u8_tc_work_bytes(int64_t N, int32_t D, int32_t side) {
  if (N <= 0 || D <= 0) return 256;
  return 256 + (side == 1 ? (size_t)D * (size_t)ef::round_up(N, 128) : 0);
}

static int gram_u8_tc_impl(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                           int64_t* G, void* work, size_t work_bytes, ef_stream_t stream, bool overwrite) {
  if (!X || !G || !work || N < 0 || D <= 0 || ldx < D || d0 < 0 || d1 > D || d0 > d1 || (side != 0 && side != 1))
    return EF_ERR_INVALID;
  if (work_bytes < ef_gram_u8_tc_work_bytes(N, D, side) || (reinterpret_cast<uintptr_t>(work) & 255))
    return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  int* status = reinterpret_cast<int*>(work);
  EF_CUDA(cudaMemsetAsync(status, 0, 16, st));
  const int64_t n_out = side == 0 ? N : D;
  if (N == 0 || (side == 0 && d1 == d0)) {
    if (overwrite && n_out > 0) EF_CUDA(cudaMemsetAsync(G, 0, sizeof(int64_t) * (size_t)n_out * (size_t)n_out, st));
    return EF_OK;
  }
  if (side == 0) {
    if (d0 & 15) return EF_ERR_UNSUPPORTED;
    return ef::gram_tc(X + d0, ldx, N, d1 - d0, G, N, status, st, overwrite, false);
  }
  if (d0 != 0 || d1 != D) return EF_ERR_UNSUPPORTED;
  // X^T X straight from the row-major X: the pixel axis is the contiguous one, i.e. MN-major UMMA operands (TMA boxes of
  // 128 samples x 128 pixel bytes).  The transposed copy (one more pass over X) is only the fallback for D < 256 or an
  // unaligned X.
  if (D >= 256 && (ldx & 15) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && !getenv("EF_GRAM_TRANSPOSE")) {
    const int rc = ef::gram_tc(X, ldx, D, N, G, D, status, st, overwrite, true);
    if (rc != EF_ERR_UNSUPPORTED) return rc;
  }
  uint8_t* XT = reinterpret_cast<uint8_t*>(work) + 256;
  const int64_t ldt = ef::round_up(N, 128);
  EF_TRY(ef::transpose_u8(X, ldx, N, D, XT, ldt, st));
  return ef::gram_tc(XT, ldt, D, N, G, D, status, st, overwrite, false);
}

int ef_gram_u8_tc_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                         int64_t* G, void* work, size_t work_bytes, ef_stream_t stream) {
  return gram_u8_tc_impl(X, ldx, N, D, d0, d1, side, G, work, work_bytes, stream, false);
}

int ef_gram_u8_tc_store_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                               int64_t* G, void* work, size_t work_bytes, ef_stream_t stream) {
  return gram_u8_tc_impl(X, ldx, N, D, d0, d1, side, G, work, work_bytes, stream, true);
}

}  // extern "C"
