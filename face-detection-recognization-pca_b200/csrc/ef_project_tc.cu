// K2a on tensor cores: the exact-integer digit-plane projection as a tcgen05 kind::i8 GEMM.
//
//   acc[n][b] += sum_d X[b][d] * Wq[n][d]      X: uint8 crops (A, K-major), Wq: int8 digit planes (B, K-major)
//
// Blackwell mapping (sm_100a):
//   * operands are staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) straight from HBM/L2 into shared memory --
//     the uint8 crops need no conversion pass, kind::i8 multiplies u8 x s8 into s32 exactly;
//   * one elected thread issues tcgen05.mma (M = 128 crops, N = all digit-plane columns of the tile, K = 32 bytes per
//     instruction), accumulators live in TMEM (up to 512 columns = 512 plane columns per pass over the crops);
//   * stream-K: the (crop tile, k block) space is cut into one contiguous range per SM, so all 148 SMs stream HBM
//     even though a 4096-crop batch has only 32 crop tiles; partial tiles are merged with int32 RED atomics, which
//     is bit-reproducible because integer addition is associative;
//   * the four epilogue warps are idle during the main loop, so they walk the same pipeline stages and compute the
//     per-crop sum of squares (needed for the reconstruction error) from the crop tile already in shared memory --
//     the crops are read from HBM exactly once;
//   * accumulators are stored plane-major (acc[n][b]) so that the 32 lanes of an epilogue warp (32 consecutive crops)
//     hit one 128-byte line per RED.
// Every mbarrier wait is bounded (2 s): on a timeout the kernel raises a status flag and drains instead of hanging.
#include <cuda.h>

#include <cstdlib>
#include <vector>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;
constexpr int kThreads = 192;                       // warp 0 TMA, warp 1 MMA + TMEM, warps 2..5 epilogue

struct Args {
  int B, NC, block_n, n_tiles, m_tiles, kb_total, stages, tmem_cols;
  int box_rows, n_loads;     // TMA boxes per B stage
  int umma_n, n_umma;        // tcgen05.mma instructions per K step
  int ld_acc;
  int32_t* acc_t;
  // split-K mode (part != null): CTA = (tile, split) with one contiguous K range of ONE tile, its partial tile is
  // STORED row-major into slab `split` (part[split][row][ld_part]) -- no atomics, no zero-initialised accumulators
  int32_t* part;
  int splits, ld_part;
  // combine != 0 (feature-major basis: the eight digit planes of a component in adjacent columns): the epilogue folds
  // them into the exact (hi, lo) int64 pair of ef::planes_to_hilo before the store -- 16 bytes per component instead of
  // 32, in slabs of long long [split][row][ld_part / 4]
  int combine;
  long long slab_stride;
  // tail split (tail_first >= 0; combined slabs, more tiles than SMs): tiles [0, tail_first) are whole waves of one CTA
  // per tile (full K range, slab 0); the tiles of the last, partial wave are split tail_splits ways along K so that the
  // wave takes 1 / tail_splits of a tile's time.  Their partial tiles go to a compact region behind slab 0:
  // [tile - tail_first][split][128 rows][block_n]
  int tail_first, tail_splits;
  // split-K mode with mcast > 1: clusters of `mcast` consecutive crop tiles of the same (column tile, K range) share the
  // basis tile -- every CTA fetches 1/mcast of it and TMA-multicasts that part into the shared memory of all of them
  int mcast, m_tiles_pad;
  double* sumsq;             // may be null
  int* status;
  unsigned long long* probe;   // optional [grid][8] timestamps (ns), debugging aid enabled by EF_TC_PROBE=1
};

__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* map, unsigned long long* bar, int c_inner,
                                               int c_outer, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer), "h"(cta_mask)
      : "memory");
}
// arrives on the barrier at the same shared-memory offset in every CTA of the mask once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_mc(unsigned long long* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

struct Seg {
  int n_tile, m_tile, kb0, kb1;
};

__global__ void __launch_bounds__(kThreads, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const Args a) {
  // 1024-byte alignment is required by the 128-byte swizzle atoms.  The array is used directly (no pointer rounding
  // through integers) so that the compiler keeps the shared address space and emits LDS/STS instead of generic loads.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  // the kernel that consumes the partial tiles (finalize_slabs_kernel) may be scheduled while this one runs: its CTAs
  // wait for this grid's completion before they read (programmatic dependent launch)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int b_stage_bytes = a.block_n * BLOCK_K;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)a.stages * A_STAGE_BYTES;
  Shared* sh = reinterpret_cast<Shared*>(sB + (size_t)a.stages * b_stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool do_sumsq = a.sumsq != nullptr;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      // MMA commit of every CTA that writes into this stage (+ one arrive per local epilogue warp)
      mbar_init(&sh->empty_bar[s], (do_sumsq ? 4 : 0) + (a.part ? a.mcast : 1));
    }
    mbar_init(&sh->tmem_full_bar, 1);
    mbar_init(&sh->tmem_empty_bar, 4);
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
  const bool clustered = a.part && a.mcast > 1;
  if (clustered) cluster_sync_all();                  // the peers' barriers exist before anything is multicast to them
  const uint32_t tmem_base = sh->tmem_base;
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 8 : nullptr;
  if (probe && threadIdx.x == 0) { probe[0] = a.probe[(size_t)gridDim.x * 8]; probe[1] = globaltimer(); }
  volatile int* failed = &sh->failed;

  // this CTA's contiguous range of (n tile, m tile, k block) units
  const long long total_units = (long long)a.n_tiles * a.m_tiles * a.kb_total;
  long long u_begin, u_end;
  int split = 0, tail_tile = -1;
  Seg fixed{0, 0, 0, 0};
  if (a.part) {
    // blockIdx = (column tile, K range, crop tile): the `mcast` CTAs of a cluster are consecutive crop tiles
    if (a.tail_first >= 0) {
      int t = (int)blockIdx.x, ways = 1;
      if (t >= a.tail_first) {
        const int i = t - a.tail_first;
        t = a.tail_first + i / a.tail_splits;
        split = i - (t - a.tail_first) * a.tail_splits;
        ways = a.tail_splits;
        tail_tile = t - a.tail_first;
      }
      fixed.n_tile = t / a.m_tiles_pad;
      fixed.m_tile = t - fixed.n_tile * a.m_tiles_pad;
      fixed.kb0 = (int)((long long)a.kb_total * split / ways);
      fixed.kb1 = (int)((long long)a.kb_total * (split + 1) / ways);
    } else {
      const int grp = blockIdx.x / a.m_tiles_pad;
      fixed.m_tile = blockIdx.x - grp * a.m_tiles_pad;       // may lie past the batch (cluster padding): loads zero fill
      fixed.n_tile = grp / a.splits;
      split = grp - fixed.n_tile * a.splits;
      fixed.kb0 = (int)((long long)a.kb_total * split / a.splits);
      fixed.kb1 = (int)((long long)a.kb_total * (split + 1) / a.splits);
    }
    u_begin = 0;
    u_end = fixed.kb1 - fixed.kb0;
  } else {
    u_begin = total_units * blockIdx.x / gridDim.x;
    u_end = total_units * (blockIdx.x + 1) / gridDim.x;
  }
  auto seg_at = [&](long long u) -> Seg {
    if (a.part) return fixed;
    const long long tile = u / a.kb_total;
    const int kb0 = (int)(u % a.kb_total);
    return Seg{(int)(tile / a.m_tiles), (int)(tile % a.m_tiles), kb0, (int)min((long long)a.kb_total, kb0 + (u_end - u))};
  };
  const uint16_t cta_mask = (uint16_t)((1u << a.mcast) - 1u);
  const int rank = clustered ? (int)cluster_ctarank() : 0;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t stage_bytes = (uint32_t)(A_STAGE_BYTES + b_stage_bytes);
      bool ok = true;
      for (long long u = u_begin; u < u_end && ok;) {
        const Seg sg = seg_at(u);
        for (int kb = sg.kb0; kb < sg.kb1; ++kb) {
          if (!mbar_wait(&sh->empty_bar[stage], phase ^ 1, failed)) { ok = false; break; }
          mbar_arrive_expect_tx(&sh->full_bar[stage], stage_bytes);
          tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, &tmap_x, &sh->full_bar[stage], kb * BLOCK_K, sg.m_tile * BLOCK_M);
          if (clustered) {
            // this CTA's share of the basis tile, delivered to every CTA of the cluster (their full barriers count it)
            tma_load_2d_mc(sB + (size_t)stage * b_stage_bytes + (size_t)rank * a.box_rows * BLOCK_K, &tmap_w,
                           &sh->full_bar[stage], kb * BLOCK_K, sg.n_tile * a.block_n + rank * a.box_rows, cta_mask);
          } else {
            for (int l = 0; l < a.n_loads; ++l)
              tma_load_2d(sB + (size_t)stage * b_stage_bytes + (size_t)l * a.box_rows * BLOCK_K, &tmap_w,
                          &sh->full_bar[stage], kb * BLOCK_K, sg.n_tile * a.block_n + l * a.box_rows);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        u += sg.kb1 - sg.kb0;
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t seg = 0;
      const uint32_t idesc = umma_idesc_i8(a.umma_n);
      bool ok = true;
      for (long long u = u_begin; u < u_end && ok; ++seg) {
        const Seg sg = seg_at(u);
        const int kb0 = sg.kb0, kb1 = sg.kb1;
        if (!mbar_wait(&sh->tmem_empty_bar, (seg & 1) ^ 1, failed)) { ok = false; break; }
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
          tc_fence_after();
          if (probe && probe[2] == 0) probe[2] = globaltimer();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = umma_desc_sw128(a_addr + k * UMMA_K);
            for (int h = 0; h < a.n_umma; ++h) {
              const uint64_t db = umma_desc_sw128(b_addr + h * a.umma_n * BLOCK_K + k * UMMA_K);
              umma_i8(tmem_base + h * a.umma_n, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          // frees the stage once these MMAs have read it -- in every CTA that multicasts into it
          if (clustered) umma_commit_mc(&sh->empty_bar[stage], cta_mask);
          else umma_commit(&sh->empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
        umma_commit(&sh->tmem_full_bar);           // accumulator of this segment is complete
        if (probe) probe[3] = globaltimer();
        u += kb1 - kb0;
      }
    }
  } else {
    // ===================================================================== epilogue warps (+ sum of squares)
    const int lane_group = warp & 3;                       // TMEM lanes [32 g, 32 g + 32) are visible to this warp
    const int row_in_tile = lane_group * 32 + lane;
    int stage = 0;
    uint32_t phase = 0;
    uint32_t seg = 0;
    bool ok = true;
    for (long long u = u_begin; u < u_end && ok; ++seg) {
      const Seg sg = seg_at(u);
      const int kb0 = sg.kb0, kb1 = sg.kb1, n_tile = sg.n_tile, m_tile = sg.m_tile;
      const int row = m_tile * BLOCK_M + row_in_tile;
      unsigned long long ssq = 0;
      if (do_sumsq) {
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
          if (n_tile == 0) {
            // the crop's 128 bytes of this K block sit in one swizzled 128-byte line; summing is order free.
            const uint4* line = reinterpret_cast<const uint4*>(sA + (size_t)stage * A_STAGE_BYTES + row_in_tile * BLOCK_K);
            unsigned int part = 0;                              // <= 128 * 255^2 per K block
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = line[(j + row_in_tile) & 7];      // rotate the chunk order: conflict-free banks
              part = __dp4a(v.x, v.x, part);
              part = __dp4a(v.y, v.y, part);
              part = __dp4a(v.z, v.z, part);
              part = __dp4a(v.w, v.w, part);
            }
            ssq += part;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh->empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (!ok) break;
      }
      if (!mbar_wait(&sh->tmem_full_bar, seg & 1, failed)) { ok = false; break; }
      tc_fence_after();
      // Launched with programmatic stream serialization, the main loop above (it only READS crops and basis) may have
      // run while the previous kernel of the stream was still draining; everything this kernel WRITES (slabs,
      // accumulators, sums of squares) waits for that kernel here.
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int c0 = 0; c0 < a.block_n; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(lane_group * 32) << 16) + (uint32_t)c0, v);
        if (a.part && a.combine) {
          if (row < a.B) {
            int32_t pl[8];
            long long h0, l0, h1, l1;
#pragma unroll
            for (int j = 0; j < 8; ++j) pl[j] = (int32_t)v[j];
            ef::planes_to_hilo(pl, h0, l0);
#pragma unroll
            for (int j = 0; j < 8; ++j) pl[j] = (int32_t)v[8 + j];
            ef::planes_to_hilo(pl, h1, l1);
            // int32-unit offset of (row, column c0) -- a multiple of 16 -- in slab `split`, or in the compact tail region
            const size_t off = tail_tile < 0
                ? (size_t)split * a.slab_stride + (size_t)row * a.ld_part + (size_t)(n_tile * a.block_n + c0)
                : (size_t)a.slab_stride + ((size_t)tail_tile * a.tail_splits + split) * ((size_t)BLOCK_M * a.block_n) +
                      (size_t)(row - m_tile * BLOCK_M) * a.block_n + (size_t)c0;
            longlong2* dst = reinterpret_cast<longlong2*>(reinterpret_cast<long long*>(a.part) + off / 4);
            __stcg(dst, make_longlong2(h0, l0));
            __stcg(dst + 1, make_longlong2(h1, l1));
          }
        } else if (a.part) {
          if (row < a.B) {
            uint4* dst = reinterpret_cast<uint4*>(a.part + (size_t)split * a.slab_stride + (size_t)row * a.ld_part +
                                                  n_tile * a.block_n + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) __stcg(dst + j, make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          }
        } else if (row < a.B) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = n_tile * a.block_n + c0 + j;
            if (col < a.NC && v[j] != 0u) atomicAdd(a.acc_t + (size_t)col * a.ld_acc + row, (int)v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->tmem_empty_bar);
      if (probe && threadIdx.x == 64) probe[4] = globaltimer();
      if (do_sumsq && n_tile == 0 && row < a.B && ssq != 0ull) atomicAdd(a.sumsq + row, (double)ssq);
      u += kb1 - kb0;
    }
  }

  // ----------------------------------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (clustered) cluster_sync_all();                  // no peer still multicasts into, or arrives on, this CTA's memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.tmem_cols)
                 : "memory");
  }
  if (threadIdx.x == 0 && sh->failed) atomicExch(a.status, 1);
  if (probe && threadIdx.x == 0) probe[5] = globaltimer();
}

}  // namespace

namespace ef_tc {

// --------------------------------------------------------------------------------------------- host side
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D byte tensor [rows][pitch] with box [box_rows][128 bytes], 128-byte swizzle, zero fill out of bounds.
bool make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t pitch, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {pitch};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace ef_tc

namespace ef {

using namespace ef_tc;

// Tile shape shared by the two schedules.
static void tile_shape(int NC, int* block_n, int* n_tiles) {
  const int nc16 = (int)round_up(NC, 16);
  if (nc16 <= 256) {
    *block_n = nc16;
    *n_tiles = 1;
  } else {
    *n_tiles = (int)ceil_div(NC, 512);
    *block_n = (int)round_up(ceil_div(NC, *n_tiles), 32);
  }
}

// Split-K schedule: clusters of `mcast` crop tiles share the basis tile; as many K ranges per tile as whole waves of one
// CTA per SM allow.
static void split_shape(int B, int D, int NC, int* splits, int* ld_part, int* mcast, int* m_tiles_pad) {
  int block_n, n_tiles;
  tile_shape(NC, &block_n, &n_tiles);
  const int m_tiles = (int)ceil_div(B, BLOCK_M);
  int c = 1;
  // measured on B200 (4096 x 10 000 x 416 columns): the main loop is bound by the kind::i8 tensor pipe, not by L2 -> SM
  // traffic, and sharing the basis tile changes nothing (72.7 us against 71.2 us per batch): off unless asked for
  if (getenv("EF_TC_MULTICAST")) {
    if (m_tiles >= 3 && block_n % 32 == 0) c = 4;
    else if (m_tiles >= 2 && block_n % 16 == 0) c = 2;
  }
  *mcast = c;
  *m_tiles_pad = (int)round_up(m_tiles, c);
  const long long tiles = (long long)n_tiles * *m_tiles_pad;
  const int kb_total = (int)ceil_div(D, BLOCK_K);
  long long s = sm_count() / tiles;
  if (s < 1) s = 1;
  if (s > kb_total) s = kb_total;
  if (s > 8) s = 8;
  *splits = (int)s;
  *ld_part = n_tiles * block_n;
}

// Tail split of the combined split-K slabs (see Args::tail_first): does it apply to this shape, and how.
bool project_tc_tail_shape(int B, int D, int NC, TcTail* t) {
  int splits, ld_part, mcast, m_tiles_pad, block_n, n_tiles;
  split_shape(B, D, NC, &splits, &ld_part, &mcast, &m_tiles_pad);
  tile_shape(NC, &block_n, &n_tiles);
  const int sms = sm_count();
  const int tiles = n_tiles * m_tiles_pad;
  TcTail r{-1, 1, block_n, m_tiles_pad, 0};
  if (t) *t = r;
  if (mcast != 1 || tiles <= sms || getenv("EF_TC_NO_TAIL_SPLIT")) return false;
  const int full = (tiles / sms) * sms, tail = tiles - full;
  if (tail == 0) return false;
  const int kb_total = (int)ceil_div(D, BLOCK_K);
  const int ways = std::min(std::min(8, sms / tail), kb_total);
  if (ways < 2) return false;
  r.first = full;
  r.splits = ways;
  r.region = (long long)B * ld_part;                  // int32 units: the tail region starts behind slab 0
  if (t) *t = r;
  return true;
}

// int32 elements of the slab buffer for batches of up to B crops (slabs + the largest tail region any batch can ask for)
size_t project_tc_part_elems(int B, int D, int NC) {
  int splits, ld_part;
  project_tc_split_shape(B, D, NC, &splits, &ld_part);
  const size_t rows = std::max<size_t>((size_t)B, (size_t)sm_count() * BLOCK_M);
  return rows * (size_t)ld_part + (size_t)sm_count() * BLOCK_M * 512;
}

void project_tc_split_shape(int B, int D, int NC, int* splits, int* ld_part) {
  int mcast, m_tiles_pad;
  split_shape(B, D, NC, splits, ld_part, &mcast, &m_tiles_pad);
}

size_t project_tc_part_bytes(int B, int D, int NC) {
  int splits, ld_part;
  project_tc_split_shape(B, D, NC, &splits, &ld_part);
  return sizeof(int32_t) * (size_t)splits * (size_t)B * ld_part;
}

int project_tc(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows,
               int32_t* acc_t, int ld_acc, double* sumsq, int* status, cudaStream_t stream, int32_t* part,
               bool combine) {
  if (B <= 0) return EF_OK;
  if (combine && !part) return EF_ERR_INVALID;
  if ((ldx & 15) || (reinterpret_cast<uintptr_t>(X) & 15) || (ldw & 15) || (reinterpret_cast<uintptr_t>(Wq) & 15))
    return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;

  Args a{};
  a.B = B;
  a.NC = NC;
  tile_shape(NC, &a.block_n, &a.n_tiles);
  if (a.block_n <= 256) {
    a.box_rows = a.block_n; a.n_loads = 1; a.umma_n = a.block_n; a.n_umma = 1;
  } else {
    a.box_rows = a.block_n / 2; a.n_loads = 2; a.umma_n = a.block_n / 2; a.n_umma = 2;
  }
  a.tmem_cols = 32;
  while (a.tmem_cols < a.block_n) a.tmem_cols *= 2;
  a.m_tiles = (int)ceil_div(B, BLOCK_M);
  a.kb_total = (int)ceil_div(D, BLOCK_K);
  const int stage_bytes = A_STAGE_BYTES + a.block_n * BLOCK_K;
  a.stages = (kSmemLimit - 1024 - (int)sizeof(Shared)) / stage_bytes;
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  if (const char* e = getenv("EF_TC_STAGES")) { const int v = atoi(e); if (v >= 2 && v < a.stages) a.stages = v; }
  if (a.stages < 2) return EF_ERR_UNSUPPORTED;
  a.ld_acc = ld_acc;
  a.acc_t = acc_t;
  a.part = part;
  a.combine = combine ? 1 : 0;
  a.mcast = 1;
  if (part) {
    split_shape(B, D, NC, &a.splits, &a.ld_part, &a.mcast, &a.m_tiles_pad);
    a.slab_stride = (long long)B * a.ld_part;
    if (a.mcast > 1) {                       // one box per CTA: its share of the basis tile
      a.box_rows = a.block_n / a.mcast;
      a.n_loads = 1;
    }
  }
  a.tail_first = -1;
  a.tail_splits = 1;
  TcTail tail{};
  if (part && combine && project_tc_tail_shape(B, D, NC, &tail)) {
    a.tail_first = tail.first;
    a.tail_splits = tail.splits;
  }
  a.sumsq = sumsq;
  a.status = status;
  a.probe = nullptr;

  CUtensorMap mx, mw;
  if (!make_map(&mx, X, (uint64_t)D, (uint64_t)B, (uint64_t)ldx, BLOCK_M)) return EF_ERR_UNSUPPORTED;
  if (!make_map(&mw, Wq, (uint64_t)ldw, (uint64_t)wq_rows, (uint64_t)ldw, (uint32_t)a.box_rows)) return EF_ERR_UNSUPPORTED;

  const size_t smem = (size_t)a.stages * stage_bytes + sizeof(Shared) + 1024;
  EF_ENSURE_SMEM(project_tc_kernel, smem);
  const long long total_units = (long long)a.n_tiles * a.m_tiles * a.kb_total;
  int grid = sm_count();
  if (const char* e = getenv("EF_TC_GRID")) { const int v = atoi(e); if (v >= 1) grid = v; }
  if (grid > total_units) grid = (int)total_units;
  if (part) grid = a.n_tiles * a.m_tiles_pad * a.splits;
  if (a.tail_first >= 0) grid = a.tail_first + (a.n_tiles * a.m_tiles_pad - a.tail_first) * a.tail_splits;
  static unsigned long long* probe_buf = nullptr;
  static int probe_grid = 0;
  const bool probing = getenv("EF_TC_PROBE") != nullptr;
  if (probing) {
    if (!probe_buf) EF_CUDA(cudaMalloc(&probe_buf, sizeof(unsigned long long) * 8 * 1025));
    EF_CUDA(cudaMemsetAsync(probe_buf, 0, sizeof(unsigned long long) * 8 * 1025, stream));
    a.probe = probe_buf;
    probe_grid = grid;
  }
  if (part && a.mcast > 1) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = (unsigned)a.mcast;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = getenv("EF_NO_PDL") ? 1 : 2;
    EF_CUDA(cudaLaunchKernelEx(&cfg, project_tc_kernel, mx, mw, a));
    ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  } else {
    EF_LAUNCH_PDL(project_tc_kernel, grid, kThreads, smem, stream, mx, mw, a);
  }
  if (probing) {
    // debugging aid: per-CTA phase timestamps relative to the first CTA entering the kernel
    std::vector<unsigned long long> h(8 * (size_t)(probe_grid + 1));
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;                       // the first CTA's setup stamp
    for (int c = 0; c < probe_grid; ++c)
      if (h[(size_t)c * 8 + 1] && h[(size_t)c * 8 + 1] < t0) t0 = h[(size_t)c * 8 + 1];
    double mx_[6] = {0}, sum_[6] = {0};
    for (int c = 0; c < probe_grid; ++c)
      for (int i = 1; i < 6; ++i) {
        const double v = h[(size_t)c * 8 + i] ? (double)(h[(size_t)c * 8 + i] - t0) * 1e-3 : 0.0;
        sum_[i] += v;
        if (v > mx_[i]) mx_[i] = v;
      }
    fprintf(stderr, "[ef_tc_probe] grid %d stages %d us since first CTA (mean/max): setup %.2f/%.2f first_full %.2f/%.2f last_mma %.2f/%.2f flush %.2f/%.2f end %.2f/%.2f\n",
            probe_grid, a.stages, sum_[1] / probe_grid, mx_[1], sum_[2] / probe_grid, mx_[2], sum_[3] / probe_grid, mx_[3],
            sum_[4] / probe_grid, mx_[4], sum_[5] / probe_grid, mx_[5]);
  }
  return EF_OK;
}

}  // namespace ef
