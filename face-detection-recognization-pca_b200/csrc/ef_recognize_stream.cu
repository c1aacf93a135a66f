// K2s, the serving kernel: ONE persistent launch recognises a QUEUE of batches (ef_model_submit_device x n, then
// ef_model_flush_device), streaming them back to back.
//
// recognize_pipe_kernel overlaps the match of batch i-1 with the stream of batch i, but every launch still pays its
// prologue, the cluster exchange + float64 combine after its stream (HBM idle) and the launch gap: 16.6 us per 41 MB
// batch against 6.3 us of HBM time.  Here the roles of a CTA are decoupled by mbarriers instead of CTA / cluster-wide
// barriers, so the loads never stop between batches:
//   warp 0        TMA producer: crop tile + basis tile per 128-byte K block, batch after batch through one stage ring
//   warp 1        tcgen05.mma kind::i8 issuer; TWO TMEM accumulators (item it -> buffer it & 1): the MMAs of batch i+1
//                 start while batch i is still being drained
//   warps 2..5    exact sum of squares of every crop from the staged tiles (dp4a), pushed to the owning CTA per item
//   warp 6        float16 [g_hi | g_lo | g_hi] image of the gallery -> shared memory, once per launch (resident)
//   warp 7        tcgen05.mma kind::f16 issuer of the nearest-row filter
//   warps 8..11   drain + combine: TMEM -> registers -> digit planes combined to TWO exact int64 per column
//                 (ef::planes_to_hilo: 2.2 x fewer bytes than the eight int32 planes) -> 16-byte st.async into the
//                 receive buffer of the CTA that owns those 32 crops; the stores themselves complete the BYTE count of
//                 the owner's mbarrier (a release.cluster arrive is a MEMBAR.ALL.GPU: measured 1.5-2 us per item on the
//                 sum-of-squares warps, which stalled the stage ring once per batch);
//                 then, for MY 32 crops: wait for the four partial slabs (acquire.cluster), exact integer sum, float64
//                 features (+ reconstruction error), exact-scorer vectors and the float16 [hi|hi|lo] filter operand
//                 (double buffered)
//   warps 12..15  match: after ONE round trip to the tensor pipe per item the scan of the filter scores in TMEM
//                 (approximate maximum per crop, then the rows inside the error band), the exact float64 re-score of
//                 those rows, best row per crop through shared-memory atomics, score / index / label
// -- three pipeline stages per CTA (stream, combine, match) working on three consecutive batches at a time.
// The filter is laid out the other way round than in recognize_cluster_kernel / recognize_pipe_kernel: the GALLERY rows
// are the M operand (128 rows per MMA, straight from the resident image) and the CTA's 32 crops the N operand, so the
// scores of one item against up to 1024 gallery rows are 8 MMA blocks of 32 TMEM columns = 256 columns, issued back to
// back and committed once.  Measured reasons: (1) with the crops on M (replicated 4 x to fill 128 lanes) every 128-row
// gallery tile was its own MMA -> commit -> scan -> release round trip behind the queued projection MMAs of the
// following batches, 16 of them per item (13-15 us per batch); (2) the replicated scores cost 4 x the TMEM read bandwidth
// (64 B / clk / SM); (3) a float32 CUDA-core filter from shared memory starves: TMA writes, MMA operand reads and the
// sum-of-squares reads of the stream already keep the shared-memory port ~90 % busy (20 us per item for the scan alone).
// A cluster of 4 CTAs owns crop tile `m` of EVERY queued batch (CTA r streams K quarter r and finishes crops 32r..32r+31),
// so the per-CTA state machines of a cluster advance through the same item sequence and the cross-CTA barriers need no
// item tags.  All arithmetic is the one of recognize_cluster_kernel / recognize_pipe_kernel (same integers, same
// combination, same filter band, same fma order): every output is bit identical to them.
//
// The basis is read in FEATURE-MAJOR plane order (row c * PS + s, PS = 4 or 8 planes per column; ef_model keeps this
// second copy) so that one 16-column tcgen05.ld holds all planes of its columns.
//
// Replaces, like the other K2 kernels, project_face_to_eigenspace + recognize_face (useless/scan.py:80-132) and
// scaler.transform + pca.transform + recognize_face_with_model (scan-template-v4.py:265-287) for a stream of batches.
#include <climits>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kCluster = 4;
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int QB = BLOCK_M / kCluster;      // crops finished by each CTA (one per lane)
static_assert(QB == 32, "one crop per lane");
constexpr int kGalTile = 128;               // gallery rows per filter MMA
constexpr int kMaxGalBlocks = 8;            // 128-row blocks of the resident gallery image: 8 x 32 score columns of TMEM
constexpr int kListCap = 128;
constexpr int kFinishWarps = 4;
constexpr int kAccCols = 128;               // TMEM columns per accumulator buffer (nc_pad <= 128)
constexpr float kFilterEps = 5e-5f;         // same bound as recognize_cluster_kernel (derived there)
constexpr int kScoreCol0 = 2 * kAccCols;    // TMEM columns 256..511: 8 blocks (128 gallery rows each) x 32 crops

struct StreamBatch {
  CUtensorMap map;             // crops of this batch: [B][ldx] bytes, box 128 rows x 128 bytes, SWIZZLE_128B
  int B, pad_;
  const double* sumsq_ext;     // weighted sum of squares (standardised models) or null
  double* out_proj;
  double* out_resid;
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;
  double threshold;
};

struct StreamArgs {
  StreamBatch batch[ef::kStreamMaxBatches];
  CUtensorMap map_w;           // feature-major digit planes [nc_pad][ldw]
  int nb, D, nc_pad, k, kq, S, PS, kb_total, stages, recv_bufs;
  const int32_t* col_exp;
  const double* bias;
  double c0;
  const double* gp;
  const double* gnorm;
  const double* ginv;
  const int32_t* labels;
  int n, kf, g_tiles, debug, prefetch, xbox;   // debug bit 0: EF_STREAM_NO_SSQ (measurement only, the residual is wrong)
  const __half* gimg;          // float16 [g_hi | g_lo | g_hi] image of the prepared gallery (gallery_image)
  int* status;
  int off_recv, off_ps, off_pe, off_bop, off_gal, off_sh;
  unsigned long long* probe;   // debugging aid (EF_TC_PROBE): [grid][8] globaltimer stamps
};

struct StreamShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long acc_full[2];
  unsigned long long acc_empty[2];
  unsigned long long recv_full[2];            // completed by the bytes of the drain + sum-of-squares pushes of all four CTAs
  unsigned long long push_ok[2][kCluster];    // [receive buffer][owner]: owner consumed that buffer (remote arrive)
  unsigned long long gal_full;                // the resident gallery image has landed (once per launch)
  unsigned long long bop_ready[2];            // combine warps: the filter operand of item it is in place ([it & 1])
  unsigned long long scores_full;             // filter MMAs of the item committed
  unsigned long long scores_free;             // match warps: the score columns have been read
  unsigned long long feat_ready[2];           // combine warps: exact-scorer vectors + norms of item it ([it & 1])
  unsigned long long feat_free[2];            // match warps: done with feature buffer [it & 1]
  unsigned long long list_ready[2];           // scan warps: the re-score list [it & 1] of item it is complete
  unsigned long long list_free[2];            // re-score warp: done with list [it & 1]
  unsigned long long best_key[QB];            // per crop: ordered-integer image of the best exact key of the item
  int best_j[QB];
  uint32_t tmem_base;
  int failed;
  int list_cnt[2], overflow[2];
  double pn[2][QB];
  double xu[QB];
  unsigned long long ssq_recv[2][kCluster][QB];
  alignas(16) float fmax_s[kFinishWarps][QB];
  int list_L[2][kListCap], list_j[2][kListCap];
};

template <int METRIC>
__device__ __forceinline__ bool better(double s, int i, double bs, int bi) {
  return s > bs || (s == bs && i < bi);
}

// Exact float64 score of gallery row j for the crop in column L of pe; same fma order as the full float64 scan.
template <int METRIC, int KR>
__device__ __forceinline__ void exact_entry(const double* __restrict__ gp, const double* __restrict__ ginv,
                                            const double* __restrict__ gnorm, const int32_t* __restrict__ labels,
                                            const double* pe, int L, int j, double pn, double& key, double& score,
                                            int& label) {
  const double2* g = reinterpret_cast<const double2*>(gp + (size_t)j * KR);
  double d = 0.0;
#pragma unroll
  for (int c = 0; c < KR; c += 2) {
    const double2 gv = __ldg(g + (c >> 1));
    d = fma(pe[c * QB + L], gv.x, d);
    d = fma(pe[(c + 1) * QB + L], gv.y, d);
  }
  if (METRIC == EF_METRIC_COSINE_G1) {
    const double gi = __ldg(ginv + j), gn = __ldg(gnorm + j);
    key = d * gi;
    score = (pn == 0.0 || gn == 0.0) ? 0.0 : d / (pn * gn);       // useless/scan.py:70-77
  } else {
    key = d;
    score = d;
  }
  label = labels ? __ldg(labels + j) : j;
}

__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// order-preserving map float -> unsigned (for redux.sync max)
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// order-preserving map double -> unsigned 64-bit (for atomicMax on the exact keys)
__device__ __forceinline__ unsigned long long d2ord(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// mbar_wait with the time spent waiting added to a probe slot (debugging aid; slot == nullptr: plain wait)
__device__ __forceinline__ bool timed_wait(unsigned long long* bar, uint32_t parity, volatile int* failed,
                                           unsigned long long* slot, bool cluster_scope = false) {
  if (!slot) return cluster_scope ? mbar_wait_cluster(bar, parity, failed) : mbar_wait(bar, parity, failed);
  const unsigned long long t0 = globaltimer();
  const bool ok = cluster_scope ? mbar_wait_cluster(bar, parity, failed) : mbar_wait(bar, parity, failed);
  *slot += globaltimer() - t0;
  return ok;
}

__device__ __forceinline__ void bar_finish() { asm volatile("bar.sync 5, 128;" ::: "memory"); }   // match warps
__device__ __forceinline__ void bar_front() { asm volatile("bar.sync 6, 128;" ::: "memory"); }    // drain + combine warps

template <int PS>
__device__ __forceinline__ void push_chunk(const uint32_t (&v)[16], int c0, int kq, uint32_t dst, uint32_t dst_bar) {
#pragma unroll
  for (int f = 0; f < 16 / PS; ++f) {
    const int c = c0 / PS + f;
    if (c < kq) {
      int32_t plane[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) plane[s] = s < PS ? (int32_t)v[f * PS + s] : 0;
      long long hi, lo;
      ef::planes_to_hilo(plane, hi, lo);
      st_async_v2_u64(dst + (uint32_t)c * (QB * 16u), (unsigned long long)hi, (unsigned long long)lo, dst_bar);
    }
  }
}

template <int METRIC, int KR>
__global__ void __launch_bounds__(kThreads, 1)
recognize_stream_kernel(const __grid_constant__ StreamArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  const int b_stage_bytes = a.nc_pad * BLOCK_K;
  const int stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)a.stages * A_STAGE_BYTES;
  uint8_t* recv = smem + a.off_recv;                              // [recv_bufs][4 sources][kq][32 crops] (hi, lo) int64
  double* ps = reinterpret_cast<double*>(smem + a.off_ps);        // [KR][QB] features of the item being finished
  double* pe = reinterpret_cast<double*>(smem + a.off_pe);        // [KR][QB] the same as the exact scorer uses them
  uint8_t* bop = smem + a.off_bop;                                // filter N operand: my 32 crops x kf float16 [hi|hi|lo]
  uint8_t* gal = smem + a.off_gal;                                // resident gallery image: g_tiles x 128 rows x kf float16
  StreamShared* sh = reinterpret_cast<StreamShared*>(smem + a.off_sh);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int m_tile = blockIdx.x / kCluster;
  const int row0 = m_tile * BLOCK_M;
  const int kb0 = (int)((long long)a.kb_total * rank / kCluster);
  const int kb1 = (int)((long long)a.kb_total * (rank + 1) / kCluster);
  const int row_bytes = a.kf * 2;
  const uint32_t gal_tile_bytes = (uint32_t)kGalTile * (uint32_t)row_bytes;
  const uint32_t bop_bytes = ((uint32_t)(QB * row_bytes) + 1023u) & ~1023u;     // one of the two filter-operand buffers
  const uint32_t recv_buf_bytes = (uint32_t)(kCluster * a.kq * QB * 16);
  // bytes one use of a receive buffer collects: four sources x (kq columns x 32 crops x (hi, lo) + 32 sums of squares)
  const uint32_t recv_tx_bytes = (uint32_t)(kCluster * (a.kq * QB * 16 + QB * 8));

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 5);             // MMA commit + the four sum-of-squares warps
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->acc_full[s], 1);
      mbar_init(&sh->acc_empty[s], 4);
      mbar_init(&sh->recv_full[s], 1);             // armed by the owner; completed by the BYTES of the async pushes
      for (int q = 0; q < kCluster; ++q) mbar_init(&sh->push_ok[s][q], 1);
    }
    mbar_init(&sh->gal_full, 1);
    mbar_init(&sh->bop_ready[0], 1);
    mbar_init(&sh->bop_ready[1], 1);
    mbar_init(&sh->scores_full, 1);
    mbar_init(&sh->scores_free, 1);
    mbar_init(&sh->feat_ready[0], 1);
    mbar_init(&sh->feat_ready[1], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sh->feat_free[i], 1);
      mbar_init(&sh->list_ready[i], 1);
      mbar_init(&sh->list_free[i], 1);
    }
    sh->failed = 0;
    for (int s = 0; s < a.recv_bufs; ++s) mbar_arrive_expect_tx(&sh->recv_full[s], recv_tx_bytes);   // first use of each buffer
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_sync_all();                               // every CTA of the cluster runs: its barriers and buffers exist
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory"); // the crops may have been written by the previous kernel
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 32 : nullptr;
  // per-stage trace of CTA 0 (first 64 stages): [0..63] TMA issue, [64..127] data landed (MMA thread), behind all CTAs' slots
  unsigned long long* trace = (a.probe && blockIdx.x == 0) ? a.probe + (size_t)gridDim.x * 32 : nullptr;
  if (probe && tid == 0) probe[0] = globaltimer();

  if (warp == 0) {
    // =================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_w) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      // Optional L2 prefetch cursor (EF_STREAM_PREFETCH = K blocks ahead of the loads, cp.async.bulk.prefetch.tensor).
      // Off by default.  Measured: a stage lands 1.0-1.4 us after its issue, four stages in flight give one stage per
      // ~350 ns (81 GB/s per SM); prefetching the crop boxes into L2 did not shorten that (the same holds with the whole
      // batch L2 resident, and with half the SMs idle) and costs 4-30 % through the extra requests.
      int n_issued = 0;
      const int pf_dist = a.prefetch;
      int pg = 0, pkb = kb0;
      auto pf_valid = [&]() {
        while (pg < a.nb && row0 >= a.batch[pg].B) ++pg;
        return pg < a.nb;
      };
      auto pf_step = [&]() {
        if (!pf_valid()) return;
        tma_prefetch_2d(&a.batch[pg].map, pkb * BLOCK_K, row0);
        if (++pkb == kb1) { pkb = kb0; ++pg; }
      };
      for (int i = 0; i < pf_dist; ++i) pf_step();
      for (int g = 0; g < a.nb && ok; ++g) {
        if (row0 >= a.batch[g].B) continue;
        const CUtensorMap* mx = &a.batch[g].map;
        asm volatile("prefetch.tensormap [%0];" ::"l"(mx) : "memory");
        for (int kb = kb0; kb < kb1; ++kb) {
          if (pf_dist > 0) pf_step();
          if (!timed_wait(&sh->empty_bar[stage], phase ^ 1, failed, probe ? probe + 16 : nullptr)) { ok = false; break; }
          mbar_arrive_expect_tx(&sh->full_bar[stage], (uint32_t)stage_bytes);
          if (trace && n_issued < 64) trace[n_issued++] = globaltimer();
          // (EF_STREAM_XBOX splits the crop tile into several boxes.  Measured: every extra TMA instruction costs ~40 ns
          // of the producer's time and nothing is gained, 128-row boxes are the fastest)
          for (int r = 0; r < BLOCK_M; r += a.xbox)
            tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES + (size_t)r * BLOCK_K, mx, &sh->full_bar[stage], kb * BLOCK_K,
                        row0 + r);
          tma_load_2d(sB + (size_t)stage * b_stage_bytes, &a.map_w, &sh->full_bar[stage], kb * BLOCK_K, 0);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================================================== projection MMA issuer
    if (lane == 0) {
      int stage = 0, it = 0, n_landed = 0;
      uint32_t phase = 0;
      const uint32_t idesc = umma_idesc_i8(a.nc_pad);
      bool ok = true;
      for (int g = 0; g < a.nb && ok; ++g) {
        if (row0 >= a.batch[g].B) continue;
        const int buf = it & 1;
        if (!timed_wait(&sh->acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), failed, probe ? probe + 18 : nullptr)) break;
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * kAccCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!timed_wait(&sh->full_bar[stage], phase, failed, probe ? probe + 17 : nullptr)) { ok = false; break; }
          if (trace && n_landed < 64) trace[64 + n_landed++] = globaltimer();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + (size_t)stage * A_STAGE_BYTES);
          const uint32_t b_addr = smem_u32(sB + (size_t)stage * b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_i8(d_addr, umma_desc_sw128(a_addr + k * UMMA_K), umma_desc_sw128(b_addr + k * UMMA_K), idesc,
                    (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&sh->empty_bar[stage]);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (ok) umma_commit(&sh->acc_full[buf]);
        if (probe) probe[it == 0 ? 1 : 2] = globaltimer();
        ++it;
      }
    }
    __syncwarp();
  } else if (warp < 6) {
    // =================================================================== exact sum of squares of the staged crop rows
    const int owner = warp & 3;                      // rows 32 owner .. 32 owner + 31 of the tile belong to CTA `owner`
    const int row_in_tile = owner * 32 + lane;
    int stage = 0, it = 0, n_ssq = 0;
    uint32_t phase = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      if (row0 >= a.batch[g].B) continue;
      unsigned long long ssq = 0;
      for (int kb = kb0; kb < kb1 && ok; ++kb) {
        ok = __all_sync(0xffffffffu, mbar_wait(&sh->full_bar[stage], phase, failed));
        if (!ok) break;
        const uint4* line = reinterpret_cast<const uint4*>(sA + (size_t)stage * A_STAGE_BYTES + row_in_tile * BLOCK_K);
        unsigned int partial = 0;
        if (!(a.debug & 1))                          // EF_STREAM_NO_SSQ (measurement only: the residual is then wrong)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = line[(j + row_in_tile) & 7];
          partial = __dp4a(v.x, v.x, partial);
          partial = __dp4a(v.y, v.y, partial);
          partial = __dp4a(v.z, v.z, partial);
          partial = __dp4a(v.w, v.w, partial);
        }
        ssq += partial;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->empty_bar[stage]);
        if (trace && warp == 2 && lane == 0 && n_ssq < 64) trace[128 + n_ssq++] = globaltimer();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (!ok) break;
      const int rb = it % a.recv_bufs, use = it / a.recv_bufs;
      ok = __all_sync(0xffffffffu, mbar_wait_cluster(&sh->push_ok[rb][owner], (uint32_t)((use & 1) ^ 1), failed));
      if (!ok) break;
      st_async_u64(map_to_cta(smem_u32(&sh->ssq_recv[rb][rank][lane]), (uint32_t)owner), ssq,
                   map_to_cta(smem_u32(&sh->recv_full[rb]), (uint32_t)owner));
      ++it;
    }
  } else if (warp == 6) {
    // =================================================================== warp 6: gallery image -> shared memory (once),
    // then the last pipeline stage of every item: exact float64 re-score of the listed rows, best row per crop, outputs
    {
      int n_items = 0;
      for (int g = 0; g < a.nb; ++g) n_items += row0 < a.batch[g].B ? 1 : 0;
      if (lane == 0 && n_items > 0) {
        const uint32_t bytes = (uint32_t)a.g_tiles * gal_tile_bytes;
        mbar_arrive_expect_tx(&sh->gal_full, bytes);
        bulk_load(gal, a.gimg, bytes, &sh->gal_full);
      }
      __syncwarp();
    }
    constexpr int kRounds = kListCap / 32;
    int it = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      const StreamBatch& bt = a.batch[g];
      if (row0 >= bt.B) continue;
      const int fb = it & 1;
      const uint32_t par = (uint32_t)((it >> 1) & 1);
      const double* pe_b = pe + (size_t)fb * KR * QB;
      sh->best_key[lane] = 0ull;
      sh->best_j[lane] = INT_MAX;
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->list_ready[fb], par, failed, (probe && lane == 0) ? probe + 25 : nullptr));
      ok = __all_sync(0xffffffffu, ok && mbar_wait(&sh->feat_ready[fb], par, failed));
      const bool overflow = *reinterpret_cast<volatile int*>(&sh->overflow[fb]) != 0;
      const int total = overflow ? 0 : min(*reinterpret_cast<volatile int*>(&sh->list_cnt[fb]), kListCap);
      // highest exact key first, then the lowest gallery row (np.argmax's first maximum); the winner writes the outputs
      int eL[kRounds], ej[kRounds], elab[kRounds];
      double ekey[kRounds], escore[kRounds];
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        const int e = r * 32 + lane;
        eL[r] = -1; ej[r] = INT_MAX; elab[r] = -1; ekey[r] = 0.0; escore[r] = 0.0;
        if (ok && e < total) {
          eL[r] = sh->list_L[fb][e];
          ej[r] = sh->list_j[fb][e];
          exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe_b, eL[r], ej[r], sh->pn[fb][eL[r]], ekey[r], escore[r],
                                  elab[r]);
        }
      }
      if (ok && overflow) {
        // degenerate gallery (more rows inside the band than list entries): exact scan of every row, lane = crop
        const double pn = sh->pn[fb][lane];
        double best = -CUDART_INF;
        for (int jj = 0; jj < a.n; ++jj) {
          double kk, ss; int ll;
          exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe_b, lane, jj, pn, kk, ss, ll);
          if (better<METRIC>(kk, jj, best, ej[0])) { best = kk; ekey[0] = kk; escore[0] = ss; elab[0] = ll; ej[0] = jj; }
        }
        if (ej[0] != INT_MAX) eL[0] = lane;
      }
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        if (ekey[r] == 0.0) ekey[r] = 0.0;           // -0.0 -> +0.0: equal keys must compare equal as ordered integers
        if (eL[r] >= 0) atomicMax(&sh->best_key[eL[r]], d2ord(ekey[r]));
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < kRounds; ++r)
        if (eL[r] >= 0 && d2ord(ekey[r]) == sh->best_key[eL[r]]) atomicMin(&sh->best_j[eL[r]], ej[r]);
      __syncwarp();
      if (probe && lane == 0 && (it == 0 || it == 4)) probe[it == 0 ? 10 : 14] = globaltimer();
#pragma unroll
      for (int r = 0; r < kRounds; ++r) {
        if (eL[r] >= 0 && d2ord(ekey[r]) == sh->best_key[eL[r]] && ej[r] == sh->best_j[eL[r]]) {
          const int b = row0 + (int)rank * QB + eL[r];
          if (b < bt.B) {
            bt.out_score[b] = escore[r];
            bt.out_index[b] = ej[r];
            if (bt.out_label) bt.out_label[b] = escore[r] >= bt.threshold ? elab[r] : -1;
          }
        }
      }
      if (sh->best_j[lane] == INT_MAX) {             // only after a pipeline failure (the status flag is raised below)
        const int b = row0 + (int)rank * QB + lane;
        if (b < bt.B) {
          bt.out_score[b] = 0.0;
          bt.out_index[b] = 0;
          if (bt.out_label) bt.out_label[b] = -1;
        }
      }
      if (probe && lane == 0) probe[it == 0 ? 4 : 5] = globaltimer();
      if (probe && lane == 0 && it == 4) probe[15] = globaltimer();
      __syncwarp();
      if (lane == 0 && ok) {
        mbar_arrive(&sh->list_free[fb]);
        mbar_arrive(&sh->feat_free[fb]);
      }
      ++it;
    }
  } else if (warp == 7) {
    // =================================================================== filter MMA issuer: one burst per item
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(QB);     // M = 128 gallery rows, N = 32 crops
      const int n_ks = a.kf >> 4;
      const int swb = row_bytes < 128 ? row_bytes : 128;
      const int pa_mask = (swb >> 5) - 1;            // k-steps per swizzle atom - 1 (row_bytes <= 128: one atom)
      const uint64_t gdesc0 = umma_desc_swz(smem_u32(gal), 0, row_bytes, kGalTile);
      const uint64_t cdesc0 = umma_desc_swz(smem_u32(bop), 0, row_bytes, QB);
      const uint64_t tile_step = (uint64_t)(gal_tile_bytes >> 4);
      int it = 0;
      bool ok = true;
      for (int g = 0; g < a.nb && ok; ++g) {
        if (row0 >= a.batch[g].B) continue;
        if (it == 0 && !mbar_wait(&sh->gal_full, 0u, failed)) break;
        if (!timed_wait(&sh->bop_ready[it & 1], (uint32_t)((it >> 1) & 1), failed, probe ? probe + 26 : nullptr)) break;
        if (!timed_wait(&sh->scores_free, (uint32_t)((it & 1) ^ 1), failed, probe ? probe + 27 : nullptr)) break;   // scan warps read item it - 1
        tc_fence_after();
        const uint64_t cdesc = cdesc0 + (uint64_t)((it & 1) * (bop_bytes >> 4));
        for (int blk = 0; blk < a.g_tiles; ++blk) {
          const uint32_t d_addr = tmem_base + (uint32_t)(kScoreCol0 + blk * QB);
          const uint64_t gd = gdesc0 + (uint64_t)blk * tile_step;
#pragma unroll 1
          for (int ks = 0; ks < n_ks; ++ks) {
            const uint64_t koff = (uint64_t)((ks & pa_mask) << 1);
            umma_f16(d_addr, gd + koff, cdesc + koff, idesc, ks > 0 ? 1u : 0u);
          }
        }
        umma_commit(&sh->scores_full);
        ++it;
      }
    }
    __syncwarp();
  } else if (warp < 12) {
    // =================================================================== drain + combine (warps 8..11, lane = crop)
    // (1) TMEM -> (hi, lo) -> receive buffer of the CTA that owns the 32 crops of my TMEM lane quarter;
    // (2) for MY CTA's 32 crops: exact integer sum of the four K quarters, float64 features, reconstruction error,
    //     exact-scorer vector + norm and the float16 filter operand, double buffered for the match warps.
    const int q = warp & 3;                          // TMEM lane quarter = crops 32 q .. 32 q + 31 = CTA q's crops
    const int fw = q;                                // this warp combines columns fw, fw + 4, ...
    constexpr int kColIters = 6;                     // columns per thread: kq <= 16, KR <= 24
    double col_scale[kColIters], col_bias[kColIters];
#pragma unroll
    for (int i = 0; i < kColIters; ++i) {
      const int c = fw + 4 * i;
      // 2^e as a double: v * 2^e == ldexp(v, e) exactly (no overflow / underflow at these magnitudes)
      col_scale[i] = c < a.kq ? __longlong_as_double((long long)(1023 + __ldg(a.col_exp + c)) << 52) : 0.0;
      col_bias[i] = c < a.k ? __ldg(a.bias + c) : 0.0;
    }
    int it = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      const StreamBatch& bt = a.batch[g];
      if (row0 >= bt.B) continue;
      const int buf = it & 1;
      // (after a failed wait the warp keeps walking through the items without touching the pipeline, so that the four
      // warps still meet at every named barrier: a timeout must end in the status flag, never in a hang)
      unsigned long long* wslot = (probe && tid == 8 * 32) ? probe + 19 : nullptr;
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->acc_full[buf], (uint32_t)((it >> 1) & 1), failed, wslot));
      tc_fence_after();
      const int rb = it % a.recv_bufs, use = it / a.recv_bufs;
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->push_ok[rb][q], (uint32_t)((use & 1) ^ 1), failed,
                                                   wslot ? wslot + 1 : nullptr, true));
      if (trace && tid == 8 * 32 && it < 16) trace[192 + 2 * it] = globaltimer();
      if (ok) {
        const uint32_t dst = map_to_cta(smem_u32(recv) + (uint32_t)rb * recv_buf_bytes +
                                            (uint32_t)((int)rank * a.kq * QB + lane) * 16u,
                                        (uint32_t)q);
        const uint32_t src = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kAccCols);
        const uint32_t dst_bar = map_to_cta(smem_u32(&sh->recv_full[rb]), (uint32_t)q);
        for (int c0 = 0; c0 < a.nc_pad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(src + (uint32_t)c0, v);
          if (a.PS == 8) push_chunk<8>(v, c0, a.kq, dst, dst_bar); else push_chunk<4>(v, c0, a.kq, dst, dst_bar);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sh->acc_empty[buf]);
        if (probe && q == 0 && lane == 0) probe[3] = globaltimer();
        if (trace && tid == 8 * 32 && it < 16) trace[193 + 2 * it] = globaltimer();
      }
      // ---- my CTA's 32 crops of this item
      const int b = row0 + (int)rank * QB + lane;    // the crop this lane finishes
      const bool live = b < bt.B;
      const int fb = it & 1;                         // feature buffer handed to the match warps
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->recv_full[rb], (uint32_t)(use & 1), failed,
                                                   wslot ? wslot + 2 : nullptr, true));
      if (probe && tid == 8 * 32 && (it == 0 || it == 4)) probe[it == 0 ? 7 : 11] = globaltimer();
      const longlong2* rbase = reinterpret_cast<const longlong2*>(recv + (size_t)rb * recv_buf_bytes) + lane;
#pragma unroll
      for (int i = 0; i < kColIters; ++i) {
        const int c = fw + 4 * i;
        if (c < max(a.kq, KR)) {
          double v = 0.0;
          if (c < a.kq && ok) {
            long long hi = 0, lo = 0;
#pragma unroll
            for (int src = 0; src < kCluster; ++src) {
              const longlong2 p = rbase[(src * a.kq + c) * QB];
              hi += p.x;
              lo += p.y;
            }
            v = ef::hilo_to_double(hi, lo) * col_scale[i];
            if (c < a.k) {
              v -= col_bias[i];
              if (bt.out_proj && live) bt.out_proj[(size_t)b * a.k + c] = v;
            } else {
              sh->xu[lane] = v;                      // residual column x . u~
            }
          }
          if (c < KR) ps[c * QB + lane] = c < a.k ? v : 0.0;   // padding columns (k .. KR) must be exact zeros
        }
      }
      unsigned long long ssq_total = 0;
      if (fw == 0) {
#pragma unroll
        for (int src = 0; src < kCluster; ++src) ssq_total += sh->ssq_recv[rb][src][lane];
      }
      bar_front();                                   // receive buffer consumed; features complete
      if (fw == 1) {                                 // re-arm the byte count, then hand the buffer back to the four sources
        if (lane == 0 && ok) mbar_arrive_expect_tx(&sh->recv_full[rb], recv_tx_bytes);
        __syncwarp();
        if (lane < kCluster) mbar_arrive_remote_relaxed(map_to_cta(smem_u32(&sh->push_ok[rb][rank]), (uint32_t)lane));
      }
      // the match warps are done with feature buffer fb (item it - 2)?
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->feat_free[fb], (uint32_t)(((it >> 1) & 1) ^ 1), failed,
                                                   wslot ? wslot + 3 : nullptr));
      double pv[KR];
#pragma unroll
      for (int c = 0; c < KR; ++c) pv[c] = ps[c * QB + lane];     // independent loads, then the ordered fma chain
      double n2 = 0.0;
#pragma unroll
      for (int c = 0; c < KR; ++c) n2 = fma(pv[c], pv[c], n2);    // columns >= k are exact zeros: same sum as c < k
      double pn = sqrt(n2);
      if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
      double* pe_b = pe + (size_t)fb * KR * QB;
#pragma unroll
      for (int c = 0; c < KR; ++c)
        if ((c & 3) == fw) pe_b[c * QB + lane] = METRIC == EF_METRIC_COSINE_SK ? pv[c] / pn : pv[c];
      if (fw == 0) {
        sh->pn[fb][lane] = pn;
        if (bt.out_resid && live) {
          const double sq = bt.sumsq_ext ? bt.sumsq_ext[b] : (double)ssq_total;
          const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
          bt.out_resid[b] = r > 0.0 ? r : 0.0;
        }
      }
      {
        // filter N operand: row `lane` = my crop, K = [hi | hi | lo] of the unit feature vector (float32 is plenty: the
        // filter is approximate by construction); chunks of 8 halfs, warp fw writes chunks fw, fw + 4, ...
        const float rinv = n2 > 0.0 ? rsqrtf((float)n2) : 0.f;
        const int KC = a.kf >> 3;
        uint8_t* bop_b = bop + (size_t)fb * bop_bytes;
        for (int kc = fw; kc < KC; kc += 4) {
          __align__(16) __half h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int kk = kc * 8 + i;
            const int seg = kk >= 3 * a.k ? 3 : (kk >= 2 * a.k ? 2 : (kk >= a.k ? 1 : 0));
            __half hi = __float2half_rn(0.f), lo = hi;
            if (seg < 3) split_half((float)ps[(kk - seg * a.k) * QB + lane] * rinv, hi, lo);
            h[i] = seg < 2 ? hi : lo;
          }
          *reinterpret_cast<uint4*>(bop_b + swz_chunk_offset(lane, kc, row_bytes, QB)) = *reinterpret_cast<const uint4*>(h);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      bar_front();                                   // pe / pn / filter operand of buffer fb complete; ps free again
      if (tid == 8 * 32 && ok) {
        mbar_arrive(&sh->bop_ready[fb]);             // -> filter MMA issuer
        mbar_arrive(&sh->feat_ready[fb]);            // -> match warps (exact-scorer vectors)
      }
      if (probe && tid == 8 * 32 && (it == 0 || it == 4)) probe[it == 0 ? 8 : 12] = globaltimer();
      ++it;
    }
  } else {
    // =================================================================== scan (warps 12..15): filter scores -> re-score list
    const int fw = warp - (kWarps - kFinishWarps), ftid = tid - (kWarps - kFinishWarps) * 32;
    int it = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      const StreamBatch& bt = a.batch[g];
      if (row0 >= bt.B) continue;
      const int lb = it & 1;
      // the re-score warp is done with list lb (item it - 2)?
      unsigned long long* wslot = (probe && ftid == 0) ? probe + 23 : nullptr;
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->list_free[lb], (uint32_t)(((it >> 1) & 1) ^ 1), failed,
                                                   wslot ? wslot + 1 : nullptr));
      if (ftid == 0) { sh->list_cnt[lb] = 0; sh->overflow[lb] = 0; }
      // ---- scores[gallery row][crop] are in TMEM: lane = gallery row 32 fw + lane of every 128-row block
      ok = __all_sync(0xffffffffu, ok && timed_wait(&sh->scores_full, (uint32_t)(it & 1), failed, wslot));
      tc_fence_after();
      const uint32_t sc_addr = tmem_base + ((uint32_t)(fw * 32) << 16) + (uint32_t)kScoreCol0;
      {
        // pass A: approximate maximum per crop over my rows, then over the warp (redux) and the four warps
        float mx[QB];
#pragma unroll
        for (int c = 0; c < QB; ++c) mx[c] = -CUDART_INF_F;
        for (int blk = 0; blk < a.g_tiles && ok; ++blk) {
          uint32_t v[32];
          tmem_ld32(sc_addr + (uint32_t)(blk * QB), v);
          if (blk * kGalTile + fw * 32 + lane < a.n) {
#pragma unroll
            for (int c = 0; c < QB; ++c) mx[c] = fmaxf(mx[c], __uint_as_float(v[c]));
          }
        }
        float mine = -CUDART_INF_F;
#pragma unroll
        for (int c = 0; c < QB; ++c) {
          const unsigned r = __reduce_max_sync(0xffffffffu, f2ord(mx[c]));
          if (c == lane) mine = ord2f(r);
        }
        sh->fmax_s[fw][lane] = mine;                 // maximum of crop `lane` over this warp's rows
      }
      bar_finish();
      if (probe && ftid == 0 && (it == 0 || it == 4)) probe[it == 0 ? 9 : 13] = globaltimer();
      {
        // pass B: rows inside the band of the maximum -> re-score list
        // every thread needs the thresholds of all 32 crops: broadcast 16-byte loads of the four warps' maxima
        float thr[QB];
#pragma unroll
        for (int c4 = 0; c4 < QB / 4; ++c4) {
          float4 M = reinterpret_cast<const float4*>(sh->fmax_s[0])[c4];
#pragma unroll
          for (int w = 1; w < kFinishWarps; ++w) {
            const float4 o = reinterpret_cast<const float4*>(sh->fmax_s[w])[c4];
            M.x = fmaxf(M.x, o.x); M.y = fmaxf(M.y, o.y); M.z = fmaxf(M.z, o.z); M.w = fmaxf(M.w, o.w);
          }
          thr[4 * c4] = M.x - 2.f * kFilterEps;
          thr[4 * c4 + 1] = M.y - 2.f * kFilterEps;
          thr[4 * c4 + 2] = M.z - 2.f * kFilterEps;
          thr[4 * c4 + 3] = M.w - 2.f * kFilterEps;
        }
        for (int blk = 0; blk < a.g_tiles && ok; ++blk) {
          uint32_t v[32];
          tmem_ld32(sc_addr + (uint32_t)(blk * QB), v);
          const int j = blk * kGalTile + fw * 32 + lane;
          unsigned mask = 0u;
#pragma unroll
          for (int c = 0; c < QB; ++c) mask |= (__uint_as_float(v[c]) >= thr[c] ? 1u : 0u) << c;
          if (j >= a.n) mask = 0u;
          while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int slot = atomicAdd(&sh->list_cnt[lb], 1);
            if (slot < kListCap) {
              sh->list_L[lb][slot] = c;
              sh->list_j[lb][slot] = j;
            } else {
              sh->overflow[lb] = 1;
            }
          }
        }
      }
      tc_fence_before();                             // the score columns have been read: the next item's MMAs may land
      __threadfence_block();
      bar_finish();                                  // list complete; fmax_s free for the next item
      if (ftid == 0 && ok) {
        mbar_arrive(&sh->scores_free);
        mbar_arrive(&sh->list_ready[lb]);
      }
      ++it;
    }
  }

  // ======================================================================= teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // no CTA leaves while a peer may still write or arrive into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
  if (probe && tid == 0) probe[6] = globaltimer();
}

struct StreamLayout {
  int stages, recv_bufs;
  int off_recv, off_ps, off_pe, off_bop, off_gal, off_sh;
  size_t smem;
};

// Shared-memory plan: the deepest stage ring that leaves room for the resident gallery image.
bool plan_layout(int nc_pad, int kq, int kr, int kf, int g_tiles, StreamLayout* out) {
  const size_t stage_bytes = (size_t)A_STAGE_BYTES + (size_t)nc_pad * BLOCK_K;
  const size_t tile_bytes = (size_t)kGalTile * kf * 2;
  const size_t bop_bytes = (size_t)ef::round_up((int64_t)QB * kf * 2, 1024);
  struct Cand { int stages, recv_bufs; };
  const Cand cands[] = {{5, 2}, {4, 2}, {4, 1}, {3, 2}, {3, 1}, {2, 1}};
  const char* e_st = getenv("EF_STREAM_STAGES");
  const char* e_rb = getenv("EF_STREAM_RECV_BUFS");
  for (const Cand& c : cands) {
    if (e_st && atoi(e_st) != c.stages) continue;
    if (e_rb && atoi(e_rb) != c.recv_bufs) continue;
    size_t off = (size_t)c.stages * stage_bytes;
    StreamLayout L{};
    L.stages = c.stages; L.recv_bufs = c.recv_bufs;
    L.off_recv = (int)off; off += (size_t)c.recv_bufs * kCluster * kq * QB * 16;
    L.off_ps = (int)off;   off += sizeof(double) * kr * QB;
    L.off_pe = (int)off;   off += 2 * sizeof(double) * kr * QB;
    off = (size_t)ef::round_up((int64_t)off, 1024);
    L.off_bop = (int)off;  off += 2 * bop_bytes;
    L.off_gal = (int)off;  off += (size_t)g_tiles * tile_bytes;
    off = (size_t)ef::round_up((int64_t)off, 128);
    L.off_sh = (int)off;
    L.smem = off + sizeof(StreamShared);
    if (L.smem <= (size_t)kSmemLimit) { *out = L; return true; }
  }
  return false;
}

template <int METRIC, int KR>
int launch_stream(StreamArgs& a, const StreamLayout& L, int m_tiles, cudaStream_t stream) {
  // one attribute call per (device, instantiation): a second device in the same process gets its own
  static size_t attr[64] = {0};
  int dev = 0;
  EF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return EF_ERR_UNSUPPORTED;
  if (L.smem > attr[dev]) {
    EF_CUDA(cudaFuncSetAttribute(recognize_stream_kernel<METRIC, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)L.smem));
    attr[dev] = L.smem;
  }
  static unsigned long long* probe_buf[64] = {nullptr};
  const bool probing = getenv("EF_TC_PROBE") != nullptr;
  const int grid_n = m_tiles * kCluster;
  a.probe = nullptr;
  if (probing && grid_n <= 4096) {
    if (!probe_buf[dev]) EF_CUDA(cudaMalloc(&probe_buf[dev], sizeof(unsigned long long) * 32 * 4096));
    EF_CUDA(cudaMemsetAsync(probe_buf[dev], 0, sizeof(unsigned long long) * 32 * 4096, stream));
    a.probe = probe_buf[dev];
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid_n);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCluster;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = getenv("EF_NO_PDL") ? 1 : 2;
  EF_CUDA(cudaLaunchKernelEx(&cfg, recognize_stream_kernel<METRIC, KR>, a));
  ef::g_launches.fetch_add(1, std::memory_order_relaxed);
  if (a.probe) {
    std::vector<unsigned long long> h((size_t)grid_n * 32 + 256);
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf[dev], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid_n; ++c) if (h[(size_t)c * 32] && h[(size_t)c * 32] < t0) t0 = h[(size_t)c * 32];
    const char* names[16] = {"start", "mma_first_item", "mma_last_item", "last_push", "first_item_out", "last_item_out",
                             "end", "i0_recv", "i0_features", "i0_scanned", "i0_rescored", "i4_recv", "i4_features",
                             "i4_scanned", "i4_rescored", "i4_out"};
    fprintf(stderr, "[ef_stream_probe] grid %d batches %d stages %d recv_bufs %d gallery blocks %d smem %zu; us since first CTA start (mean/max):",
            grid_n, a.nb, a.stages, a.recv_bufs, a.g_tiles, L.smem);
    for (int i = 0; i < 16; ++i) {
      if (names[i][0] == '-') continue;
      double sum = 0, mx = 0;
      int cnt = 0;
      for (int c = 0; c < grid_n; ++c) {
        if (!h[(size_t)c * 32 + i]) continue;
        const double v = (double)(h[(size_t)c * 32 + i] - t0) * 1e-3;
        sum += v;
        ++cnt;
        if (v > mx) mx = v;
      }
      fprintf(stderr, " %s %.2f/%.2f", names[i], cnt ? sum / cnt : 0.0, mx);
    }
    {
      const unsigned long long* tr = h.data() + (size_t)grid_n * 32;
      fprintf(stderr, "\n[ef_stream_probe] CTA 0, first stages: issue / landed / sum-of-squares arrive (ns since first issue):");
      for (int i = 0; i < 64; ++i)
        if (tr[i] && tr[64 + i]) fprintf(stderr, " %llu/%llu/%llu", tr[i] - tr[0], tr[64 + i] - tr[0], tr[128 + i] ? tr[128 + i] - tr[0] : 0ull);
      fprintf(stderr, "\n[ef_stream_probe] CTA 0, drain start/end per item (ns since first issue):");
      for (int i = 0; i < 16; ++i)
        if (tr[192 + 2 * i]) fprintf(stderr, " %llu-%llu", tr[192 + 2 * i] - tr[0], tr[193 + 2 * i] - tr[0]);
    }
    // accumulated wait times of the roles (slots 16..): where each pipeline stage spends its idle time
    const char* wnames[12] = {"tma:empty", "mma:full", "mma:acc_empty", "drain:acc_full", "drain:push_ok", "comb:recv_full",
                              "comb:feat_free", "scan:scores_full", "scan:list_free", "rescore:list_ready", "fmma:bop_ready",
                              "fmma:scores_free"};
    fprintf(stderr, "\n[ef_stream_probe] waits, us per launch (mean/max over CTAs):");
    for (int i = 0; i < 12; ++i) {
      double sum = 0, mx = 0;
      for (int c = 0; c < grid_n; ++c) {
        const double v = (double)h[(size_t)c * 32 + 16 + i] * 1e-3;
        sum += v;
        if (v > mx) mx = v;
      }
      fprintf(stderr, " %s %.1f/%.1f", wnames[i], sum / grid_n, mx);
    }
    fprintf(stderr, "\n");
  }
  return EF_OK;
}

template <int METRIC>
int dispatch_kr(StreamArgs& a, const StreamLayout& L, int kr, int m_tiles, cudaStream_t st) {
  switch (kr) {
    case 4: return launch_stream<METRIC, 4>(a, L, m_tiles, st);
    case 8: return launch_stream<METRIC, 8>(a, L, m_tiles, st);
    case 12: return launch_stream<METRIC, 12>(a, L, m_tiles, st);
    case 16: return launch_stream<METRIC, 16>(a, L, m_tiles, st);
    default: return launch_stream<METRIC, 24>(a, L, m_tiles, st);
  }
}

}  // namespace

namespace ef {

int stream_plane_stride(int S) { return S <= 4 ? 4 : 8; }

bool stream_supported(int D, int k, int kq, int S, int metric, int64_t n) {
  const int nc_pad = (int)round_up((int64_t)kq * stream_plane_stride(S), 16);
  if (metric == EF_METRIC_L2 || k > 21 || nc_pad > kAccCols || n <= 0 || n > (int64_t)kMaxGalBlocks * kGalTile) return false;
  if (filter_kf(k) > 64 || ceil_div(D, BLOCK_K) < kCluster) return false;
  StreamLayout L;
  return plan_layout(nc_pad, kq, fused_epilogue_kpad(k), filter_kf(k), (int)ceil_div(n, kGalTile), &L);
}

int stream_encode_batch(StreamBatchDesc* d, int D) {
  using namespace ef_tc;
  if (!d || d->B <= 0 || !d->x || (d->ldx & 15) || (reinterpret_cast<uintptr_t>(d->x) & 15)) return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  CUtensorMap m;
  if (!make_map(&m, d->x, (uint64_t)D, (uint64_t)d->B, (uint64_t)d->ldx, BLOCK_M)) return EF_ERR_UNSUPPORTED;
  memcpy(d->tmap, &m, sizeof(m));
  return EF_OK;
}

// One persistent launch over nb <= kStreamMaxBatches queued batches.  EF_ERR_UNSUPPORTED outside the kernel's coverage.
int recognize_stream(const StreamBatchDesc* batches, int nb, int D, const int8_t* Wfm, int64_t ldw, int wfm_rows, int k,
                     int kq, int S, const int32_t* col_exp, const double* bias, double c0, const double* gp_padded,
                     int kpad, const double* gnorm, const double* ginv, const void* gimg, int64_t n,
                     const int32_t* labels, int metric, int* status, cudaStream_t stream) {
  using namespace ef_tc;
  if (nb <= 0) return EF_OK;
  if (nb > kStreamMaxBatches) return EF_ERR_INVALID;
  if (!stream_supported(D, k, kq, S, metric, n) || kpad != fused_epilogue_kpad(k) || kpad > 24) return EF_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(gimg) & 15) return EF_ERR_UNSUPPORTED;
  const int PS = stream_plane_stride(S);
  const int nc_pad = (int)round_up((int64_t)kq * PS, 16);
  if (nc_pad > wfm_rows || (ldw & 15) || (reinterpret_cast<uintptr_t>(Wfm) & 15)) return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  StreamArgs a{};
  a.xbox = BLOCK_M;                                  // rows per TMA box of the crop tile (multiple of 8, divides 128)
  if (const char* e = getenv("EF_STREAM_XBOX")) {
    const int v = atoi(e);
    if (v == 8 || v == 16 || v == 32 || v == 64 || v == 128) a.xbox = v;
  }
  int max_B = 0;
  for (int g = 0; g < nb; ++g) {
    const StreamBatchDesc& d = batches[g];
    if (d.B <= 0 || !d.x || (d.ldx & 15) || (reinterpret_cast<uintptr_t>(d.x) & 15)) return EF_ERR_UNSUPPORTED;
    StreamBatch& b = a.batch[g];
    static_assert(sizeof(CUtensorMap) == sizeof(d.tmap), "tensor map blob");
    if (a.xbox == BLOCK_M) {
      memcpy(&b.map, d.tmap, sizeof(CUtensorMap));   // encoded at submit time
    } else if (!make_map(&b.map, d.x, (uint64_t)D, (uint64_t)d.B, (uint64_t)d.ldx, (uint32_t)a.xbox)) {
      return EF_ERR_UNSUPPORTED;
    }
    b.B = d.B;
    b.sumsq_ext = d.sumsq_ext;
    b.out_proj = d.out_proj; b.out_resid = d.out_resid; b.out_score = d.out_score; b.out_index = d.out_index;
    b.out_label = d.out_label; b.threshold = d.threshold;
    max_B = std::max(max_B, d.B);
  }
  if (!make_map(&a.map_w, Wfm, (uint64_t)ldw, (uint64_t)wfm_rows, (uint64_t)ldw, (uint32_t)nc_pad)) return EF_ERR_UNSUPPORTED;
  StreamLayout L;
  const int g_tiles = (int)ceil_div(n, kGalTile);
  if (!plan_layout(nc_pad, kq, kpad, filter_kf(k), g_tiles, &L)) return EF_ERR_UNSUPPORTED;
  a.nb = nb; a.D = D; a.nc_pad = nc_pad; a.k = k; a.kq = kq; a.S = S; a.PS = PS;
  a.kb_total = (int)ceil_div(D, BLOCK_K);
  a.stages = L.stages; a.recv_bufs = L.recv_bufs;
  a.col_exp = col_exp; a.bias = bias; a.c0 = c0;
  a.gp = gp_padded; a.gnorm = gnorm; a.ginv = ginv; a.labels = labels; a.n = (int)n;
  a.kf = filter_kf(k); a.g_tiles = g_tiles;
  a.debug = getenv("EF_STREAM_NO_SSQ") ? 1 : 0;
  a.prefetch = 0;                                    // K blocks (16 KB crop boxes) the L2 prefetch cursor runs ahead
  if (const char* e = getenv("EF_STREAM_PREFETCH")) a.prefetch = std::max(0, std::min(64, atoi(e)));
  a.gimg = reinterpret_cast<const __half*>(gimg);
  a.status = status;
  a.off_recv = L.off_recv; a.off_ps = L.off_ps; a.off_pe = L.off_pe; a.off_bop = L.off_bop; a.off_gal = L.off_gal;
  a.off_sh = L.off_sh;
  const int m_tiles = (int)ceil_div(max_B, BLOCK_M);
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(a, L, kpad, m_tiles, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(a, L, kpad, m_tiles, stream);
    default: return EF_ERR_UNSUPPORTED;
  }
}

}  // namespace ef
