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
//   warp 6        float32 image of the normalised gallery -> shared memory (resident when it fits, else a
//                 cp.async.bulk ring)
//   warps 8..11   drain: TMEM -> registers -> digit planes combined to TWO exact int64 per column (ef::planes_to_hilo:
//                 2.2 x fewer bytes than the eight int32 planes) -> 16-byte st.shared::cluster into the receive buffer of
//                 the CTA that owns those 32 crops -> remote mbarrier arrive (release.cluster)
//   warps 12..15  finish: wait for the four partial slabs of MY 32 crops (acquire.cluster), exact integer sum, float64
//                 features (+ reconstruction error), ONE pass of a float32 CUDA-core filter over the gallery (lane =
//                 crop, features in registers, gallery rows broadcast from shared memory, running maximum + the rows
//                 within the error band of it), the exact float64 re-score of those rows and score / index / label --
//                 all while warps 0..11 already work on the next batch.
// (The first version ran the filter as tcgen05 kind::f16 MMAs like recognize_pipe_kernel.  Measured: every filter MMA
// queues behind the projection MMAs of the following batches in the one tensor pipe, 16 dependent round trips of ~0.8 us
// per item made the finish warps the bottleneck at 13-15 us per batch.  The filter is 0.3 MFLOP per crop: on the idle
// FP32 pipe it needs no round trips at all.)
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
constexpr int kMaxRing = 16;                // gallery tiles in shared memory (resident gallery: up to 2048 rows)
constexpr int kListCap = 128;
constexpr int kFinishWarps = 4;
constexpr int kFinishThreads = kFinishWarps * 32;
constexpr int kAccCols = 128;               // TMEM columns per accumulator buffer (nc_pad <= 128)
constexpr int kCandCap = 8;                 // filter candidates a thread keeps (pruned against the running maximum)
// Filter error bound: features and gallery rows are unit vectors rounded to float32 (2^-24 relative per component), the
// k <= 24 products are accumulated with fmaf (2^-24 relative per step), the feature normalisation uses rsqrtf (one
// common factor 1 + 2^-22 on all scores of a crop): |s~ - cos| < 4e-6.  5e-5 (the band of the tensor-core filters) covers
// it with a wide margin; a too-large value only costs extra float64 re-scores, never correctness.
constexpr float kFilterEps = 5e-5f;

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
  int n, ring, g_tiles, resident;
  const float* gimg;           // float32 image of the normalised gallery [g_tiles * 128][KR]
  int* status;
  int off_recv, off_ps, off_pe, off_gal, off_sh;
  unsigned long long* probe;   // debugging aid (EF_TC_PROBE): [grid][8] globaltimer stamps
};

struct StreamShared {
  unsigned long long full_bar[kMaxStages];
  unsigned long long empty_bar[kMaxStages];
  unsigned long long acc_full[2];
  unsigned long long acc_empty[2];
  unsigned long long recv_full[2];            // arrived on by the drain + sum-of-squares lanes of all four CTAs
  unsigned long long push_ok[2][kCluster];    // [receive buffer][owner]: owner consumed that buffer (remote arrive)
  unsigned long long gal_full[kMaxRing];
  unsigned long long gal_empty[kMaxRing];
  uint32_t tmem_base;
  int failed;
  int list_cnt, overflow;
  double pn[QB];
  double xu[QB];
  unsigned long long ssq_recv[2][kCluster][QB];
  float fmax_s[kFinishWarps][QB];
  int cand_j[kCandCap][kFinishThreads];          // per-thread candidate rows of the filter pass ([slot][thread]: no conflicts)
  float cand_s[kCandCap][kFinishThreads];
  int list_L[kListCap], list_j[kListCap], list_label[kListCap];
  double list_key[kListCap], list_score[kListCap];
  int red_l[kFinishWarps][QB];
  double red_s[kFinishWarps][QB];
  double red_d[kFinishWarps][QB];
  int red_i[kFinishWarps][QB];
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

__device__ __forceinline__ void bar_finish() { asm volatile("bar.sync 5, 128;" ::: "memory"); }

template <int PS>
__device__ __forceinline__ void push_chunk(const uint32_t (&v)[16], int c0, int kq, uint32_t dst) {
#pragma unroll
  for (int f = 0; f < 16 / PS; ++f) {
    const int c = c0 / PS + f;
    if (c < kq) {
      int32_t plane[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) plane[s] = s < PS ? (int32_t)v[f * PS + s] : 0;
      long long hi, lo;
      ef::planes_to_hilo(plane, hi, lo);
      st_cluster_v2_u64(dst + (uint32_t)c * (QB * 16u), (unsigned long long)hi, (unsigned long long)lo);
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
  uint8_t* gal = smem + a.off_gal;
  StreamShared* sh = reinterpret_cast<StreamShared*>(smem + a.off_sh);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int m_tile = blockIdx.x / kCluster;
  const int row0 = m_tile * BLOCK_M;
  const int kb0 = (int)((long long)a.kb_total * rank / kCluster);
  const int kb1 = (int)((long long)a.kb_total * (rank + 1) / kCluster);
  const uint32_t gal_tile_bytes = (uint32_t)(kGalTile * KR * sizeof(float));
  const int n_seq = a.g_tiles;                     // one filter pass over the gallery per item
  const uint32_t recv_buf_bytes = (uint32_t)(kCluster * a.kq * QB * 16);

  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(&sh->full_bar[s], 1);
      mbar_init(&sh->empty_bar[s], 5);             // MMA commit + the four sum-of-squares warps
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&sh->acc_full[s], 1);
      mbar_init(&sh->acc_empty[s], 4);
      mbar_init(&sh->recv_full[s], kCluster * 2 * QB);   // 4 sources x (32 drain lanes + 32 sum-of-squares lanes)
      for (int q = 0; q < kCluster; ++q) mbar_init(&sh->push_ok[s][q], 1);
    }
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&sh->gal_full[s], 1);
      mbar_init(&sh->gal_empty[s], kFinishWarps);
    }
    sh->failed = 0;
    sh->list_cnt = 0;
    sh->overflow = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(2u * kAccCols)
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
  unsigned long long* probe = a.probe ? a.probe + (size_t)blockIdx.x * 8 : nullptr;
  if (probe && tid == 0) probe[0] = globaltimer();

  if (warp == 0) {
    // =================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&a.map_w) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int g = 0; g < a.nb && ok; ++g) {
        if (row0 >= a.batch[g].B) continue;
        const CUtensorMap* mx = &a.batch[g].map;
        asm volatile("prefetch.tensormap [%0];" ::"l"(mx) : "memory");
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->empty_bar[stage], phase ^ 1, failed)) { ok = false; break; }
          mbar_arrive_expect_tx(&sh->full_bar[stage], (uint32_t)stage_bytes);
          tma_load_2d(sA + (size_t)stage * A_STAGE_BYTES, mx, &sh->full_bar[stage], kb * BLOCK_K, row0);
          tma_load_2d(sB + (size_t)stage * b_stage_bytes, &a.map_w, &sh->full_bar[stage], kb * BLOCK_K, 0);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =================================================================== projection MMA issuer
    if (lane == 0) {
      int stage = 0, it = 0;
      uint32_t phase = 0;
      const uint32_t idesc = umma_idesc_i8(a.nc_pad);
      bool ok = true;
      for (int g = 0; g < a.nb && ok; ++g) {
        if (row0 >= a.batch[g].B) continue;
        const int buf = it & 1;
        if (!mbar_wait(&sh->acc_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1), failed)) break;
        tc_fence_after();
        const uint32_t d_addr = tmem_base + (uint32_t)(buf * kAccCols);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (!mbar_wait(&sh->full_bar[stage], phase, failed)) { ok = false; break; }
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
    int stage = 0, it = 0;
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
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      if (!ok) break;
      const int rb = it % a.recv_bufs, use = it / a.recv_bufs;
      ok = __all_sync(0xffffffffu, mbar_wait_cluster(&sh->push_ok[rb][owner], (uint32_t)((use & 1) ^ 1), failed));
      if (!ok) break;
      st_cluster_u64(map_to_cta(smem_u32(&sh->ssq_recv[rb][rank][lane]), (uint32_t)owner), ssq);
      mbar_arrive_remote(map_to_cta(smem_u32(&sh->recv_full[rb]), (uint32_t)owner));
      ++it;
    }
  } else if (warp == 6) {
    // =================================================================== float16 gallery image -> shared memory
    if (lane == 0) {
      int n_items = 0;
      for (int g = 0; g < a.nb; ++g) n_items += row0 < a.batch[g].B ? 1 : 0;
      const uint8_t* img = reinterpret_cast<const uint8_t*>(a.gimg);   // [g_tiles] tiles of 128 rows x KR floats
      if (a.resident) {
        if (n_items > 0)
          for (int t = 0; t < a.g_tiles; ++t) {
            mbar_arrive_expect_tx(&sh->gal_full[t], gal_tile_bytes);
            bulk_load(gal + (size_t)t * gal_tile_bytes, img + (size_t)t * gal_tile_bytes, gal_tile_bytes, &sh->gal_full[t]);
          }
      } else {
        const long long total = (long long)n_items * n_seq;
        int slot = 0, t = 0;
        uint32_t use_parity = 1;                     // parity of the (use - 1)-th release; a fresh barrier passes parity 1
        for (long long s = 0; s < total; ++s) {
          if (!mbar_wait(&sh->gal_empty[slot], use_parity, failed)) break;
          mbar_arrive_expect_tx(&sh->gal_full[slot], gal_tile_bytes);
          bulk_load(gal + (size_t)slot * gal_tile_bytes, img + (size_t)t * gal_tile_bytes, gal_tile_bytes, &sh->gal_full[slot]);
          if (++slot == a.ring) { slot = 0; use_parity ^= 1; }
          if (++t == a.g_tiles) t = 0;
        }
      }
    }
    __syncwarp();
  } else if (warp == 7) {
    // spare warp
  } else if (warp < 12) {
    // =================================================================== drain: TMEM -> (hi, lo) -> owner's receive buffer
    const int q = warp & 3;                          // TMEM lane quarter = crops 32 q .. 32 q + 31 = CTA q's crops
    int it = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      if (row0 >= a.batch[g].B) continue;
      const int buf = it & 1;
      ok = __all_sync(0xffffffffu, mbar_wait(&sh->acc_full[buf], (uint32_t)((it >> 1) & 1), failed));
      if (!ok) break;
      tc_fence_after();
      const int rb = it % a.recv_bufs, use = it / a.recv_bufs;
      ok = __all_sync(0xffffffffu, mbar_wait_cluster(&sh->push_ok[rb][q], (uint32_t)((use & 1) ^ 1), failed));
      if (!ok) break;
      const uint32_t dst = map_to_cta(smem_u32(recv) + (uint32_t)rb * recv_buf_bytes +
                                          (uint32_t)((int)rank * a.kq * QB + lane) * 16u,
                                      (uint32_t)q);
      const uint32_t src = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kAccCols);
      for (int c0 = 0; c0 < a.nc_pad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(src + (uint32_t)c0, v);
        if (a.PS == 8) push_chunk<8>(v, c0, a.kq, dst); else push_chunk<4>(v, c0, a.kq, dst);
      }
      mbar_arrive_remote(map_to_cta(smem_u32(&sh->recv_full[rb]), (uint32_t)q));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sh->acc_empty[buf]);
      if (probe && q == 0 && lane == 0) probe[3] = globaltimer();
      ++it;
    }
  } else {
    // =================================================================== finish: features, filter scan, exact re-score
    const int fw = warp - (kWarps - kFinishWarps), ftid = tid - (kWarps - kFinishWarps) * 32;
    int it = 0;
    unsigned int gs = 0;
    bool ok = true;
    for (int g = 0; g < a.nb; ++g) {
      const StreamBatch& bt = a.batch[g];
      if (row0 >= bt.B) continue;
      const int b = row0 + (int)rank * QB + lane;    // the crop this lane finishes
      const bool live = b < bt.B;
      const bool want_resid = bt.out_resid != nullptr;
      const int rb = it % a.recv_bufs, use = it / a.recv_bufs;
      ok = __all_sync(0xffffffffu, ok && mbar_wait_cluster(&sh->recv_full[rb], (uint32_t)(use & 1), failed));
      // ---- exact integer sum of the four K quarters, float64 features (columns fw, fw + 4, ...)
      const longlong2* rbase = reinterpret_cast<const longlong2*>(recv + (size_t)rb * recv_buf_bytes) + lane;
      for (int c = fw; c < max(a.kq, KR); c += kFinishWarps) {
        double v = 0.0;
        if (c < a.kq && ok) {
          long long hi = 0, lo = 0;
#pragma unroll
          for (int src = 0; src < kCluster; ++src) {
            const longlong2 p = rbase[(src * a.kq + c) * QB];
            hi += p.x;
            lo += p.y;
          }
          v = ldexp(ef::hilo_to_double(hi, lo), __ldg(a.col_exp + c));
          if (c < a.k) {
            v -= __ldg(a.bias + c);
            if (bt.out_proj && live) bt.out_proj[(size_t)b * a.k + c] = v;
          } else {
            sh->xu[lane] = v;                        // residual column x . u~
          }
        }
        if (c < KR) ps[c * QB + lane] = c < a.k ? v : 0.0;   // padding columns (k .. KR) must be exact zeros
      }
      unsigned long long ssq_total = 0;
      if (fw == 0) {
#pragma unroll
        for (int src = 0; src < kCluster; ++src) ssq_total += sh->ssq_recv[rb][src][lane];
        if (lane == 0) { sh->list_cnt = 0; sh->overflow = 0; }
      }
      bar_finish();                                  // receive buffer consumed; features complete
      if (fw == 1 && lane < kCluster)                // hand the buffer back to the four sources
        mbar_arrive_remote(map_to_cta(smem_u32(&sh->push_ok[rb][rank]), (uint32_t)lane));
      double n2 = 0.0;
      for (int c = 0; c < a.k; ++c) n2 = fma(ps[c * QB + lane], ps[c * QB + lane], n2);
      double pn = sqrt(n2);
      if (METRIC == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
      for (int c = fw; c < KR; c += kFinishWarps) {
        double v = ps[c * QB + lane];
        if (METRIC == EF_METRIC_COSINE_SK) v = v / pn;
        pe[c * QB + lane] = v;
      }
      if (fw == 0) {
        sh->pn[lane] = pn;
        if (want_resid && live) {
          const double sq = bt.sumsq_ext ? bt.sumsq_ext[b] : (double)ssq_total;
          const double r = sq - 2.0 * sh->xu[lane] + a.c0 - n2;
          bt.out_resid[b] = r > 0.0 ? r : 0.0;
        }
      }
      double best = -CUDART_INF, best_score = 0.0;
      int best_i = INT_MAX, best_label = -1;
      auto consider = [&](double key, double score, int label, int j) {
        if (better<METRIC>(key, j, best, best_i)) { best = key; best_score = score; best_label = label; best_i = j; }
      };
      // ---- float32 filter, one pass: running maximum of this warp's rows and the rows within the band of it
      float ph[KR];
      {
        const float rinv = n2 > 0.0 ? rsqrtf((float)n2) : 0.f;
#pragma unroll
        for (int c = 0; c < KR; ++c) ph[c] = (float)ps[c * QB + lane] * rinv;
      }
      const float band = 2.f * kFilterEps;
      float m = -CUDART_INF_F;
      int cnt = 0;
      bool spill = false;
      for (int t = 0; t < n_seq; ++t, ++gs) {
        int slot;
        uint32_t par;
        if (a.resident) { slot = t; par = 0u; } else { slot = (int)(gs % (unsigned)a.ring); par = (gs / (unsigned)a.ring) & 1u; }
        ok = __all_sync(0xffffffffu, ok && mbar_wait(&sh->gal_full[slot], par, failed));
        if (!ok) continue;
        const float* gt = reinterpret_cast<const float*>(gal + (size_t)slot * gal_tile_bytes) + (size_t)(fw * 32) * KR;
        const int j0 = t * kGalTile + fw * 32;
        // groups of 8 rows: the 8 scores are formed first (independent fmaf chains, no side effects, so the loads and
        // the arithmetic of a group pipeline freely), one comparison of their maximum against the band decides whether
        // the rare candidate bookkeeping (shared-memory stores the compiler must order against the gallery loads) runs
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float sc[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4* g4 = reinterpret_cast<const float4*>(gt + (r0 + u) * KR);
            float acc = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < KR / 4; ++c4) {
              const float4 gv = g4[c4];              // all lanes read the same row: shared-memory broadcast
              acc = fmaf(ph[4 * c4], gv.x, acc);
              acc = fmaf(ph[4 * c4 + 1], gv.y, acc);
              acc = fmaf(ph[4 * c4 + 2], gv.z, acc);
              acc = fmaf(ph[4 * c4 + 3], gv.w, acc);
            }
            sc[u] = acc;
          }
          const float gm = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])), fmaxf(fmaxf(sc[4], sc[5]), fmaxf(sc[6], sc[7])));
          if (gm >= m - band) {                      // rare after the first rows
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int j = j0 + r0 + u;
              if (j < a.n && sc[u] >= m - band) {
                if (sc[u] > m) m = sc[u];
                if (cnt == kCandCap) {               // drop what fell out of the band of the running maximum
                  const float lim = m - band;
                  int w = 0;
                  for (int i = 0; i < kCandCap; ++i) {
                    const float cs = sh->cand_s[i][ftid];
                    if (cs >= lim) {
                      sh->cand_s[w][ftid] = cs;
                      sh->cand_j[w][ftid] = sh->cand_j[i][ftid];
                      ++w;
                    }
                  }
                  cnt = w;
                }
                if (cnt < kCandCap) {
                  sh->cand_s[cnt][ftid] = sc[u];
                  sh->cand_j[cnt][ftid] = j;
                  ++cnt;
                } else {
                  spill = true;                      // more than kCandCap rows of one crop inside the band: exact scan
                }
              }
            }
          }
        }
        if (!a.resident) {
          __syncwarp();
          if (lane == 0) mbar_arrive(&sh->gal_empty[slot]);
        }
      }
      sh->fmax_s[fw][lane] = m;
      if (spill) sh->overflow = 1;
      bar_finish();
      {
        float M = sh->fmax_s[0][lane];
#pragma unroll
        for (int w = 1; w < kFinishWarps; ++w) M = fmaxf(M, sh->fmax_s[w][lane]);
        const float thr = M - band;
        for (int i = 0; i < cnt; ++i) {
          if (sh->cand_s[i][ftid] >= thr) {
            const int slot = atomicAdd(&sh->list_cnt, 1);
            if (slot < kListCap) {
              sh->list_L[slot] = lane;
              sh->list_j[slot] = sh->cand_j[i][ftid];
            } else {
              sh->overflow = 1;
            }
          }
        }
      }
      __threadfence_block();
      bar_finish();
      const bool overflow = *reinterpret_cast<volatile int*>(&sh->overflow) != 0;
      const int total = overflow ? 0 : min(*reinterpret_cast<volatile int*>(&sh->list_cnt), kListCap);
      for (int e = ftid; e < total; e += kFinishThreads) {
        const int L = sh->list_L[e];
        double key, score; int label;
        exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe, L, sh->list_j[e], sh->pn[L], key, score, label);
        sh->list_key[e] = key;
        sh->list_score[e] = score;
        sh->list_label[e] = label;
      }
      bar_finish();
      for (int e = fw; e < total; e += kFinishWarps)
        if (sh->list_L[e] == lane) consider(sh->list_key[e], sh->list_score[e], sh->list_label[e], sh->list_j[e]);
      if (overflow) {
        for (int j = fw; j < a.n; j += kFinishWarps) {
          double key, score; int label;
          exact_entry<METRIC, KR>(a.gp, a.ginv, a.gnorm, a.labels, pe, lane, j, pn, key, score, label);
          consider(key, score, label, j);
        }
      }
      sh->red_s[fw][lane] = best;
      sh->red_d[fw][lane] = best_score;
      sh->red_i[fw][lane] = best_i;
      sh->red_l[fw][lane] = best_label;
      bar_finish();
      if (fw == 0 && live) {
        double bs = sh->red_s[0][lane], score = sh->red_d[0][lane];
        int bi = sh->red_i[0][lane], bl = sh->red_l[0][lane];
        for (int w = 1; w < kFinishWarps; ++w)
          if (better<METRIC>(sh->red_s[w][lane], sh->red_i[w][lane], bs, bi)) {
            bs = sh->red_s[w][lane];
            score = sh->red_d[w][lane];
            bi = sh->red_i[w][lane];
            bl = sh->red_l[w][lane];
          }
        if (bi == INT_MAX) { bi = 0; bl = -1; }      // only after a pipeline failure (the status flag is raised below)
        bt.out_score[b] = score;
        bt.out_index[b] = bi;
        if (bt.out_label) bt.out_label[b] = score >= bt.threshold ? bl : -1;
      }
      if (probe && ftid == 0) probe[it == 0 ? 4 : 5] = globaltimer();
      bar_finish();                                  // red_* / list_* / pn are rewritten by the next item
      ++it;
    }
  }

  // ======================================================================= teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // no CTA leaves while a peer may still write or arrive into it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kAccCols) : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
  if (probe && tid == 0) probe[6] = globaltimer();
}

struct StreamLayout {
  int stages, recv_bufs, ring, resident;
  int off_recv, off_ps, off_pe, off_gal, off_sh;
  size_t smem;
};

// Shared-memory plan: the deepest stage ring with the gallery image resident comes first.
bool plan_layout(int nc_pad, int kq, int kr, int g_tiles, StreamLayout* out) {
  const size_t stage_bytes = (size_t)A_STAGE_BYTES + (size_t)nc_pad * BLOCK_K;
  const size_t tile_bytes = (size_t)kGalTile * kr * sizeof(float);
  struct Cand { int stages, recv_bufs, resident, ring; };
  const Cand cands[] = {{5, 2, 1, 0}, {4, 2, 1, 0}, {4, 1, 1, 0}, {3, 2, 1, 0}, {3, 1, 1, 0}, {4, 2, 0, 4},
                        {4, 1, 0, 4}, {3, 2, 0, 4}, {3, 1, 0, 3}, {3, 1, 0, 2}, {2, 1, 0, 2}};
  const char* e_st = getenv("EF_STREAM_STAGES");
  const char* e_res = getenv("EF_STREAM_RESIDENT");
  const char* e_rb = getenv("EF_STREAM_RECV_BUFS");
  for (const Cand& c : cands) {
    if (e_st && atoi(e_st) != c.stages) continue;
    if (e_res && atoi(e_res) != c.resident) continue;
    if (e_rb && atoi(e_rb) != c.recv_bufs) continue;
    int ring = c.resident ? g_tiles : std::min(c.ring, g_tiles);
    if (ring > kMaxRing || ring < 1) continue;
    if (!c.resident && ring < 2 && g_tiles >= 2) continue;
    size_t off = (size_t)c.stages * stage_bytes;
    StreamLayout L{};
    L.stages = c.stages; L.recv_bufs = c.recv_bufs; L.ring = ring; L.resident = c.resident;
    L.off_recv = (int)off; off += (size_t)c.recv_bufs * kCluster * kq * QB * 16;
    L.off_ps = (int)off;   off += sizeof(double) * kr * QB;
    L.off_pe = (int)off;   off += sizeof(double) * kr * QB;
    off = (size_t)ef::round_up((int64_t)off, 128);
    L.off_gal = (int)off;  off += (size_t)ring * tile_bytes;
    off = (size_t)ef::round_up((int64_t)off, 128);
    L.off_sh = (int)off;
    L.smem = off + sizeof(StreamShared);
    if (L.smem <= (size_t)kSmemLimit) { *out = L; return true; }
  }
  return false;
}

// prepared gallery -> float32 unit rows [n_pad][kr] (zero rows / zero padding stay zero)
__global__ void stream_gallery_kernel(const double* __restrict__ gp, int kr, const double* __restrict__ ginv, int n, int k,
                                      int metric, float* __restrict__ img) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * kr) return;
  const int j = (int)(e / kr), c = (int)(e - (long long)j * kr);
  const double scale = metric == EF_METRIC_COSINE_G1 ? ginv[j] : 1.0;
  img[e] = c < k ? (float)(gp[e] * scale) : 0.f;
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
    if (!probe_buf[dev]) EF_CUDA(cudaMalloc(&probe_buf[dev], sizeof(unsigned long long) * 8 * 4096));
    EF_CUDA(cudaMemsetAsync(probe_buf[dev], 0, sizeof(unsigned long long) * 8 * 4096, stream));
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
    std::vector<unsigned long long> h((size_t)grid_n * 8);
    EF_CUDA(cudaStreamSynchronize(stream));
    EF_CUDA(cudaMemcpy(h.data(), probe_buf[dev], h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < grid_n; ++c) if (h[(size_t)c * 8] && h[(size_t)c * 8] < t0) t0 = h[(size_t)c * 8];
    const char* names[8] = {"start", "mma_first_item", "mma_last_item", "last_push", "first_item_out", "last_item_out",
                            "end", "-"};
    fprintf(stderr, "[ef_stream_probe] grid %d batches %d stages %d recv_bufs %d ring %d resident %d smem %zu; us since first CTA start (mean/max):",
            grid_n, a.nb, a.stages, a.recv_bufs, a.ring, a.resident, L.smem);
    for (int i = 0; i < 8; ++i) {
      if (names[i][0] == '-') continue;
      double sum = 0, mx = 0;
      int cnt = 0;
      for (int c = 0; c < grid_n; ++c) {
        if (!h[(size_t)c * 8 + i]) continue;
        const double v = (double)(h[(size_t)c * 8 + i] - t0) * 1e-3;
        sum += v;
        ++cnt;
        if (v > mx) mx = v;
      }
      fprintf(stderr, " %s %.2f/%.2f", names[i], cnt ? sum / cnt : 0.0, mx);
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
  if (metric == EF_METRIC_L2 || k > 24 || nc_pad > kAccCols || n <= 0 || n >= (1ll << 31) - 512) return false;
  if (ceil_div(D, BLOCK_K) < kCluster) return false;
  StreamLayout L;
  return plan_layout(nc_pad, kq, fused_epilogue_kpad(k), (int)ceil_div(n, kGalTile), &L);
}

size_t stream_gallery_bytes(int k, int64_t n) {
  return (size_t)ceil_div(n, kGalTile) * kGalTile * (size_t)fused_epilogue_kpad(k) * sizeof(float);
}

// float32 image of a prepared gallery (gp [n][kr], rows normalised for COSINE_SK; ginv = 1/|g| for COSINE_G1)
int stream_gallery_image(const double* gp, int kr, const double* ginv, int64_t n, int k, int metric, void* img,
                         cudaStream_t stream) {
  if (n <= 0) return EF_OK;
  EF_CUDA(cudaMemsetAsync(img, 0, stream_gallery_bytes(k, n), stream));
  EF_LAUNCH(stream_gallery_kernel, (unsigned)ceil_div(n * kr, 256), 256, 0, stream, gp, kr, ginv, (int)n, k, metric,
            reinterpret_cast<float*>(img));
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
  const int PS = stream_plane_stride(S);
  const int nc_pad = (int)round_up((int64_t)kq * PS, 16);
  if (nc_pad > wfm_rows || (ldw & 15) || (reinterpret_cast<uintptr_t>(Wfm) & 15)) return EF_ERR_UNSUPPORTED;
  if (!encode_fn()) return EF_ERR_UNSUPPORTED;
  StreamArgs a{};
  int max_B = 0;
  for (int g = 0; g < nb; ++g) {
    const StreamBatchDesc& d = batches[g];
    if (d.B <= 0 || !d.x || (d.ldx & 15) || (reinterpret_cast<uintptr_t>(d.x) & 15)) return EF_ERR_UNSUPPORTED;
    StreamBatch& b = a.batch[g];
    if (!make_map(&b.map, d.x, (uint64_t)D, (uint64_t)d.B, (uint64_t)d.ldx, BLOCK_M)) return EF_ERR_UNSUPPORTED;
    b.B = d.B;
    b.sumsq_ext = d.sumsq_ext;
    b.out_proj = d.out_proj; b.out_resid = d.out_resid; b.out_score = d.out_score; b.out_index = d.out_index;
    b.out_label = d.out_label; b.threshold = d.threshold;
    max_B = std::max(max_B, d.B);
  }
  if (!make_map(&a.map_w, Wfm, (uint64_t)ldw, (uint64_t)wfm_rows, (uint64_t)ldw, (uint32_t)nc_pad)) return EF_ERR_UNSUPPORTED;
  StreamLayout L;
  const int g_tiles = (int)ceil_div(n, kGalTile);
  if (!plan_layout(nc_pad, kq, kpad, g_tiles, &L)) return EF_ERR_UNSUPPORTED;
  a.nb = nb; a.D = D; a.nc_pad = nc_pad; a.k = k; a.kq = kq; a.S = S; a.PS = PS;
  a.kb_total = (int)ceil_div(D, BLOCK_K);
  a.stages = L.stages; a.recv_bufs = L.recv_bufs; a.ring = L.ring; a.resident = L.resident;
  a.col_exp = col_exp; a.bias = bias; a.c0 = c0;
  a.gp = gp_padded; a.gnorm = gnorm; a.ginv = ginv; a.labels = labels; a.n = (int)n;
  a.g_tiles = g_tiles;
  a.gimg = reinterpret_cast<const float*>(gimg);
  a.status = status;
  a.off_recv = L.off_recv; a.off_ps = L.off_ps; a.off_pe = L.off_pe; a.off_gal = L.off_gal; a.off_sh = L.off_sh;
  const int m_tiles = (int)ceil_div(max_B, BLOCK_M);
  switch (metric) {
    case EF_METRIC_COSINE_SK: return dispatch_kr<EF_METRIC_COSINE_SK>(a, L, kpad, m_tiles, stream);
    case EF_METRIC_COSINE_G1: return dispatch_kr<EF_METRIC_COSINE_G1>(a, L, kpad, m_tiles, stream);
    default: return EF_ERR_UNSUPPORTED;
  }
}

}  // namespace ef
