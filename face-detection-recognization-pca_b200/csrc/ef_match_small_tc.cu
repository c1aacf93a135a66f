// K2t: residual + nearest-gallery search + threshold/label in ONE launch for the shipped model shapes (k = 50 ... 178,
// galleries of 178 ... 590 rows), with the all-pairs scan on TENSOR CORES and float64 only where it decides.
//
// match_small_kernel (ef_match_small.cu) forms every (query, gallery row) dot product on the FP64 pipe: 24-43 us per
// 4096 crops, latency bound.  Here the scan is the float16 hi/lo filter GEMM of ef_match_tc.cu (tcgen05.mma kind::f16,
// float32 accumulation in TMEM; the SAME gallery image, built once per model by ef_match_tc_prepare_device) and only
// the rows inside the proven error band of a query's approximate maximum are scored in float64 -- with EXACTLY the
// arithmetic of match_small_kernel / match_kernel (same |p|^2 order, same sequential fma chain, same quotient, same
// tie rule: lowest index), so score, index, label and residual are bit identical.
//
// Schedule: mst_query_kernel (a warp per query) forms |p| in the order of match_small_kernel, finishes the residual,
// and writes the query's float16 [hi | hi | lo] operand row into a pre-swizzled image (and p / |p| for the sklearn
// metric).  match_small_tc_kernel: grid = (128-query tiles) x (64-row gallery pieces); a CTA bulk-loads its query tile
// (<= 9 slabs of 16 KB) and its 64 rows of every gallery K slab (8 KB each: rows of a SWIZZLE_128B slab are
// contiguous), issues <= 36 UMMAs (M = 128, N = 64) into a 64-column TMEM accumulator, and each of 128 scanning
// threads (= one query) takes its 64 approximate scores with two tcgen05.ld, keeps the rows within 2 eps of their
// maximum (one row unless the gallery holds near-duplicates) and scores them in float64.  The per-piece winners meet
// in global memory; the last CTA of a query tile to finish (a self-resetting counter) picks the best of the pieces and
// writes score / index / label.  (A first version converted the queries inside every CTA: 185 us at k = 178 -- six
// lone warps per SM at 13 cycles per instruction; ncu source view in profiles/r2_summary.md.)
//
// Replaces (per batch) the cosine loop + max of useless/scan.py:121-130 and cosine_similarity + argmax + threshold of
// scan-template-v4.py:274-287 for the shapes the fused k <= 32 kernels do not cover.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <vector>
#include <cstdlib>
#include <cuda_fp16.h>
#include <math_constants.h>

#include "ef_common.cuh"
#include "ef_internal.cuh"
#include "ef_tc_common.cuh"

namespace {

using namespace ef_tc;

constexpr int kThreads = 192;               // warp 0 bulk loads, warp 1 MMA + TMEM, warps 2..5 scan and re-score
constexpr int BNP = 64;                     // smallest piece of gallery rows per CTA (64, 128 or 256: MstArgs::bnp)
constexpr int kSlab = 64;                   // halfs of K per slab = one 128-byte swizzle row
constexpr int kSlabBytesA = BLOCK_M * 128;  // 16 KB
constexpr int kTileRows = 256;              // rows per tile of the ef_match_tc image
constexpr int kTileSlabBytes = kTileRows * 128;
constexpr int kMaxSlabs = 9;                // resident operand tiles: 3 (k + 1) <= 576
constexpr int kMaxSlabsStream = 49;         // operand tiles streamed through a ring: k <= 1024
constexpr int kListCap = 512;                // streaming mode: (query, candidate) pairs listed per CTA
constexpr int kRing = 6;                    // most stages of that ring (MstArgs::ring of them in use)
constexpr int kMaxRows = 4096;              // 64 pieces

struct MstArgs {
  int B, k, ka, n_slabs, metric;
  const double* Pe;              // rows the exact chain multiplies: p / |p| (COSINE_SK) or p
  int64_t lde;
  const uint8_t* qimg;           // [query tile][slab][128 rows x 128 B, SWIZZLE_128B] from mst_query_kernel
  const double* pn;              // [b_pad] |p| (1 for a zero query under COSINE_SK)
  const uint8_t* img;
  const double* gscale;          // L2: largest gallery norm (image trailer)
  const double* G;
  int64_t ldg;
  const double* gnorm;
  int n;
  const int32_t* labels;
  double threshold;
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;            // nullable
  int sh_off;                    // byte offset of MstShared behind the operand tiles / the staging area
  int stream;                    // n_slabs > 9: the K slabs of both operands stream through a ring of kRing stages and the
                                 // exact chains read global memory (the float64 rows of k > 191 do not fit beside each other)
  int cluster;                   // the CTAs of a query tile (one per gallery piece) form a thread-block cluster
  int ex_off;                    // cluster mode: byte offset of the exchange area [pieces][128] (double, then int)
  int exm_off;                   // ... and of the approximate maxima [pieces][128] float
  int ring;                      // streaming mode: stages of the ring
  int slots;                     // streaming mode: (query, candidate) pairs staged per round
  int list_off;                  // streaming mode: byte offset of the pair list
  int bnp;                       // gallery rows per CTA = UMMA N = TMEM columns
  int pieces, b_pad, rows_round, bulk;   // staged float64 rows: query rows per round; bulk copies possible
  double* part_s;                // [pieces][b_pad]
  int* part_i;                   // [pieces][b_pad]
  unsigned int* counters;        // [query tiles], zero between launches
  int* status;
  float eps;
  long long* trace;              // debug (EF_MST_TRACE): [CTA][8] clock64 stamps
};

struct MstShared {
  unsigned long long full_bar;
  unsigned long long ring_full[kRing];       // streaming mode: stage filled / drained
  unsigned long long ring_empty[kRing];
  unsigned long long tmem_full_bar;
  uint32_t tmem_base;
  int failed;
  int is_last;
};

__device__ __forceinline__ bool better(int metric, double s, int i, double bs, int bi) {
  if (metric == EF_METRIC_L2) return s < bs || (s == bs && i < bi);
  return s > bs || (s == bs && i < bi);
}

// exact score of (query row p, gallery row j): the arithmetic of match_small_kernel -- ONE fma chain over ascending
// components.  The operands are independent of the chain: sixteen components are loaded ahead of it (128-bit loads when
// both rows allow), so the chain waits for memory once per sixteen steps, not once per step.
template <bool L2>
__device__ __forceinline__ double chain(const double* __restrict__ p, const double* __restrict__ g, int k) {
  double acc = 0.0;
  int c = 0;
  if ((((uintptr_t)p | (uintptr_t)g) & 15) == 0) {
    const double2* __restrict__ p2 = reinterpret_cast<const double2*>(p);
    const double2* __restrict__ g2 = reinterpret_cast<const double2*>(g);
    for (; c + 16 <= k; c += 16) {
      double2 pv[8], gv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { pv[i] = p2[(c >> 1) + i]; gv[i] = g2[(c >> 1) + i]; }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (L2) {
          const double d0 = pv[i].x - gv[i].x, d1 = pv[i].y - gv[i].y;
          acc = fma(d0, d0, acc);
          acc = fma(d1, d1, acc);
        } else {
          acc = fma(pv[i].x, gv[i].x, acc);
          acc = fma(pv[i].y, gv[i].y, acc);
        }
      }
    }
  }
  for (; c + 8 <= k; c += 8) {
    double pv[8], gv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { pv[i] = p[c + i]; gv[i] = g[c + i]; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (L2) {
        const double d = pv[i] - gv[i];
        acc = fma(d, d, acc);
      } else {
        acc = fma(pv[i], gv[i], acc);
      }
    }
  }
  for (; c < k; ++c) {
    if (L2) {
      const double d = p[c] - g[c];
      acc = fma(d, d, acc);
    } else {
      acc = fma(p[c], g[c], acc);
    }
  }
  return acc;
}

__device__ __forceinline__ double exact_score(const MstArgs& a, const double* __restrict__ pe, double pn, int j) {
  const double* __restrict__ g = a.G + (int64_t)j * a.ldg;
  if (a.metric == EF_METRIC_L2) return chain<true>(pe, g, a.k);
  const double acc = chain<false>(pe, g, a.k);
  if (a.metric == EF_METRIC_COSINE_SK) return acc;
  const double gn = a.gnorm[j];
  return (pn == 0.0 || gn == 0.0) ? 0.0 : acc / (pn * gn);
}

// ---- one warp per query: exact |p| (the order of match_small_kernel: lane-strided fma chains + xor-shuffle tree), the
// residual, p / |p| for the sklearn metric, and the float16 [hi | hi | lo] operand row, normalised in float32.  Rows
// B .. b_pad of the last tile are written as zeros (a shorter batch after a longer one); the K padding of the image is
// zero from the reservation and never written.
struct MstQueryArgs {
  // COMBINED split-K slabs of the projection (part != null; long long [split][crop][ld_part], component c at [2c] = hi,
  // [2c + 1] = lo): the kernel first forms the float64 features itself -- the work of finalize_slabs_hilo_kernel
  // (ef_epilogue.cu: exact sum of the pairs over the splits, one rounding), one warp per crop -- and writes them to P
  const long long* part;
  int splits, ld_part, kq;
  long long slab_stride;
  const int32_t* col_exp;
  const double* bias;
  double* P;
  int64_t ldp;
  int B, b_pad, k, ka, n_slabs, metric;
  const double* gscale;
  double* sumsq;                 // nullable; consumed and cleared
  double c0;
  double* resid2;                // nullable; holds x . u~ on entry, the reconstruction error on exit
  uint8_t* qimg;
  double* pn;
  double* pe;                    // [b_pad][lde] rows of the exact chain: p / |p| (COSINE_SK) or p, pitch k | 1
  int lde;
};

__global__ void __launch_bounds__(256)
mst_query_kernel(const MstQueryArgs a) {
  // programmatic dependent launch: the matcher's CTAs may be scheduled (barriers, TMEM, shared-memory carve-out) while
  // this grid runs; they wait for its completion (griddepcontrol.wait) before they read what it writes
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");              // (itself launched early behind the features' producer)
  const int q = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= a.b_pad) return;
  const bool live = q < a.B;
  double* __restrict__ p = a.P + (int64_t)q * a.ldp;
  constexpr int PER = 6;                             // k <= 192
  double pv[PER];
  double acc = 0.0;
  if (a.part) {
    double xu = 0.0;                                 // the residual column x . u~ (column k when kq > k)
#pragma unroll
    for (int i = 0; i < PER + 1; ++i) {
      const int c = lane + 32 * i;
      double v = 0.0;
      if (live && c < a.kq) {
        long long hi = 0, lo = 0;
        for (int sp = 0; sp < a.splits; ++sp) {
          const longlong2 t = __ldcg(reinterpret_cast<const longlong2*>(a.part + (size_t)sp * a.slab_stride +
                                                                          (size_t)q * a.ld_part) + c);
          hi += t.x;
          lo += t.y;
        }
        v = ldexp(ef::hilo_to_double(hi, lo), a.col_exp[c]);
        if (c < a.k) {
          v = v - a.bias[c];
          p[c] = v;
        } else {
          xu = v;
          v = 0.0;
        }
      }
      if (i < PER) pv[i] = c < a.k ? v : 0.0;
    }
    if (a.resid2 && a.kq > a.k) {
      xu = ef::warp_sum(xu);                         // one lane holds it, the others add zeros
      if (lane == 0 && live) a.resid2[q] = xu;
    }
  } else {
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const int c = lane + 32 * i;
      pv[i] = (live && c < a.k) ? p[c] : 0.0;
    }
  }
#pragma unroll
  for (int i = 0; i < PER; ++i)
    if (lane + 32 * i < a.k) acc = fma(pv[i], pv[i], acc);
  const double s2 = ef::warp_sum(acc);
  double pn = sqrt(s2);
  if (a.metric == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
  if (lane == 0) {
    a.pn[q] = pn;
    if (a.resid2 && live) {
      const double rr = a.sumsq[q] - 2.0 * a.resid2[q] + a.c0 - s2;
      a.resid2[q] = rr > 0.0 ? rr : 0.0;
      a.sumsq[q] = 0.0;
    }
  }
  const double dinv = s2 > 0.0 ? 1.0 / sqrt(s2) : 0.0;
  float rinv = (float)dinv;
  float last = 0.f;                                  // L2: the extra query component -t r (see ef_match_tc.cu)
  if (a.metric == EF_METRIC_L2) {
    const double G = *a.gscale;
    const double rq = s2 > 0.0 ? 0.5 * G * dinv : 1.0;
    const double t = rq > 1.0 ? 1.0 / rq : 1.0;
    rinv = (float)(dinv * t);
    last = live ? -(float)(rq * t) : 0.f;
  }
  const int tile = q / BLOCK_M, r = q - tile * BLOCK_M;
  uint8_t* __restrict__ row = a.qimg + (size_t)tile * a.n_slabs * kSlabBytesA + (size_t)r * 128;
  const int rx = r & 7;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;                     // component of the augmented row (c == k: the L2 extra; ka <= 192)
    if (c >= a.ka) continue;
    const double pc = pv[i];
    if (live && c < a.k) a.pe[(int64_t)q * a.lde + c] = a.metric == EF_METRIC_COSINE_SK ? pc / pn : pc;   // normalize()
    const float v = c < a.k ? (float)pc * rinv : last;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
#pragma unroll
    for (int seg = 0; seg < 3; ++seg) {
      const int kk = c + seg * a.ka;
      const int slab = kk >> 6, kin = kk & 63;
      *reinterpret_cast<__half*>(row + (size_t)slab * kSlabBytesA + ((((kin >> 3) ^ rx)) << 4) + (kin & 7) * 2) =
          seg < 2 ? hi : lo;
    }
  }
}

// The same for k > 192 (the row does not fit the registers of a warp: two passes over it, no slab finalize)
__global__ void __launch_bounds__(256)
mst_query_big_kernel(const MstQueryArgs a) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int q = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= a.b_pad) return;
  const bool live = q < a.B;
  const double* __restrict__ p = a.P + (int64_t)q * a.ldp;
  double acc = 0.0;
  if (live)
    for (int c = lane; c < a.k; c += 32) {
      const double v = p[c];
      acc = fma(v, v, acc);
    }
  const double s2 = ef::warp_sum(acc);
  double pn = sqrt(s2);
  if (a.metric == EF_METRIC_COSINE_SK && pn == 0.0) pn = 1.0;
  if (lane == 0) {
    a.pn[q] = pn;
    if (a.resid2 && live) {
      const double rr = a.sumsq[q] - 2.0 * a.resid2[q] + a.c0 - s2;
      a.resid2[q] = rr > 0.0 ? rr : 0.0;
      a.sumsq[q] = 0.0;
    }
  }
  const double dinv = s2 > 0.0 ? 1.0 / sqrt(s2) : 0.0;
  float rinv = (float)dinv;
  float last = 0.f;
  if (a.metric == EF_METRIC_L2) {
    const double G = *a.gscale;
    const double rq = s2 > 0.0 ? 0.5 * G * dinv : 1.0;
    const double t = rq > 1.0 ? 1.0 / rq : 1.0;
    rinv = (float)(dinv * t);
    last = live ? -(float)(rq * t) : 0.f;
  }
  const int tile = q / BLOCK_M, r = q - tile * BLOCK_M;
  uint8_t* __restrict__ row = a.qimg + (size_t)tile * a.n_slabs * kSlabBytesA + (size_t)r * 128;
  const int rx = r & 7;
  for (int c = lane; c < a.ka; c += 32) {
    const double pc = (live && c < a.k) ? p[c] : 0.0;
    if (live && c < a.k) a.pe[(int64_t)q * a.lde + c] = a.metric == EF_METRIC_COSINE_SK ? pc / pn : pc;
    const float v = c < a.k ? (float)pc * rinv : last;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
#pragma unroll
    for (int seg = 0; seg < 3; ++seg) {
      const int kk = c + seg * a.ka;
      const int slab = kk >> 6, kin = kk & 63;
      *reinterpret_cast<__half*>(row + (size_t)slab * kSlabBytesA + ((((kin >> 3) ^ rx)) << 4) + (kin & 7) * 2) =
          seg < 2 ? hi : lo;
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
match_small_tc_kernel(const MstArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0) atomicExch(a.status, 2);
    return;
  }
  uint8_t* sA = smem;                                            // [n_slabs][128 rows][128 B]
  uint8_t* sB = smem + (size_t)a.n_slabs * kSlabBytesA;          // [n_slabs][bnp rows][128 B]
  const int kPieceBytes = a.bnp * 128;
  MstShared* sh = reinterpret_cast<MstShared*>(smem + a.sh_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, piece = blockIdx.y;
  const int row_base = piece * a.bnp;
  // the next kernel of the stream (the projection of the following batch: it only reads until its own
  // griddepcontrol.wait) may be scheduled on SMs this grid leaves free
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  long long* tr = a.trace ? a.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 : nullptr;
#define MST_STAMP(i) if (tr && threadIdx.x == 64) tr[i] = clock64()
  MST_STAMP(0);

  if (tid == 0) {
    mbar_init(&sh->full_bar, 1);
    for (int st = 0; st < kRing; ++st) {
      mbar_init(&sh->ring_full[st], 1);
      mbar_init(&sh->ring_empty[st], 1);
    }
    mbar_init(&sh->tmem_full_bar, 1);
    sh->failed = 0;
    sh->is_last = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"((uint32_t)a.bnp)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // everything up to griddepcontrol.wait overlaps the tail of mst_query_kernel (programmatic dependent launch),
  // including the loads of this CTA's 64 rows of every gallery K slab (the gallery image is static); after it the
  // kernel's outputs (query image, |p|, rows of the exact chain) are complete and visible
  if (tid == 0 && !a.stream) {
    const int tile = row_base / kTileRows, sub_bytes = (row_base % kTileRows) * 128;
    mbar_arrive_expect_tx(&sh->full_bar, (uint32_t)(a.n_slabs * (kSlabBytesA + kPieceBytes)));
    for (int slab = 0; slab < a.n_slabs; ++slab)
      bulk_load(sB + (size_t)slab * kPieceBytes,
                a.img + ((size_t)tile * a.n_slabs + slab) * kTileSlabBytes + (size_t)sub_bytes,
                (uint32_t)kPieceBytes, &sh->full_bar);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (tid == 0 && !a.stream) {
    for (int slab = 0; slab < a.n_slabs; ++slab)                 // the query tile, pre-swizzled by mst_query_kernel
      bulk_load(sA + (size_t)slab * kSlabBytesA, a.qimg + ((size_t)qt * a.n_slabs + slab) * kSlabBytesA,
                (uint32_t)kSlabBytesA, &sh->full_bar);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;
  volatile int* failed = &sh->failed;

  if (a.stream) {
    // ---- k > 191: one K slab of the query tile (16 KB) + the piece's rows of the same slab of the gallery per stage
    const int stage_bytes = kSlabBytesA + kPieceBytes;
    if (warp == 0 && lane == 0) {
      const int tile = row_base / kTileRows, sub_bytes = (row_base % kTileRows) * 128;
      int stage = 0;
      uint32_t phase = 0;
      for (int slab = 0; slab < a.n_slabs; ++slab) {
        if (!mbar_wait(&sh->ring_empty[stage], phase ^ 1, failed)) break;
        mbar_arrive_expect_tx(&sh->ring_full[stage], (uint32_t)stage_bytes);
        uint8_t* dst = smem + (size_t)stage * stage_bytes;
        bulk_load(dst, a.qimg + ((size_t)qt * a.n_slabs + slab) * kSlabBytesA, (uint32_t)kSlabBytesA, &sh->ring_full[stage]);
        bulk_load(dst + kSlabBytesA, a.img + ((size_t)tile * a.n_slabs + slab) * kTileSlabBytes + (size_t)sub_bytes,
                  (uint32_t)kPieceBytes, &sh->ring_full[stage]);
        if (++stage == a.ring) { stage = 0; phase ^= 1; }
      }
    } else if (warp == 1 && lane == 0) {
      const uint32_t idesc = umma_idesc_f16(a.bnp);
      int stage = 0;
      uint32_t phase = 0;
      bool fed = true;
      for (int slab = 0; slab < a.n_slabs; ++slab) {
        if (!mbar_wait(&sh->ring_full[stage], phase, failed)) { fed = false; break; }
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint32_t b_addr = a_addr + kSlabBytesA;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_f16(tmem_base, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc,
                   (slab > 0 || ks > 0) ? 1u : 0u);
        umma_commit(&sh->ring_empty[stage]);
        if (++stage == a.ring) { stage = 0; phase ^= 1; }
      }
      if (fed) umma_commit(&sh->tmem_full_bar);
    }
  } else if (warp == 1 && lane == 0 && mbar_wait(&sh->full_bar, 0, failed)) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(a.bnp);
    for (int slab = 0; slab < a.n_slabs; ++slab) {
      const uint32_t a_addr = smem_u32(sA + (size_t)slab * kSlabBytesA);
      const uint32_t b_addr = smem_u32(sB + (size_t)slab * kPieceBytes);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_f16(tmem_base, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc,
                 (slab > 0 || ks > 0) ? 1u : 0u);
    }
    umma_commit(&sh->tmem_full_bar);
  }
  __syncwarp();
  // every warp waits for the accumulator: from here on the operand tiles are dead and their shared memory is the
  // staging area of the float64 rows
  const bool scanning = warp >= 2;
  const int lane_group = warp & 3;
  const int r = lane_group * 32 + lane;
  const int q = qt * BLOCK_M + r;                                 // a scanning thread's query
  const bool live = scanning && q < a.B;
  const int valid = min(a.bnp, a.n - row_base);
  MST_STAMP(1);
  const bool ok = __syncthreads_and(mbar_wait(&sh->tmem_full_bar, 0, failed)) != 0;
  MST_STAMP(2);
  double best = (a.metric == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
  int best_i = INT_MAX;
  unsigned mask[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};            // candidate rows of this thread's query, 32 per word
  const int groups = a.bnp >> 5;
  float amax = -CUDART_INF_F;                                     // approximate maximum of this thread's query
  if (ok && scanning) {
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_group * 32) << 16);
    // pass 1: the approximate maximum over the piece (four independent chains)
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < groups && g * 32 < valid) {                          // warp uniform
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(g * 32), v);
        const int vg = valid - g * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          if (i < vg) m0 = fmaxf(m0, __uint_as_float(v[i]));
          if (i + 1 < vg) m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
          if (i + 2 < vg) m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
          if (i + 3 < vg) m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
        }
      }
    }
    amax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
  }
  if (a.cluster) {
    // ---- the pieces of a query tile tell each other their approximate maxima (DSMEM, one cluster barrier): a row is
    // a candidate only inside the band of the maximum over the WHOLE gallery, so a query costs ~one exact chain in
    // the cluster instead of one per piece (the same argument as the shared running maxima of ef_match_tc.cu)
    float* ex_m = reinterpret_cast<float*>(smem + a.exm_off);     // [pieces][128]: row `rank` is written by CTA `rank`
    const uint32_t rank = cluster_ctarank();
    if (scanning) {
      const uint32_t mine = smem_u32(ex_m + (size_t)rank * BLOCK_M + r);
      for (int pc = 0; pc < a.pieces; ++pc) st_cluster_u32(map_to_cta(mine, (uint32_t)pc), __float_as_uint(amax));
    }
    cluster_arrive();
    cluster_wait();
    if (scanning)
      for (int pc = 0; pc < a.pieces; ++pc) amax = fmaxf(amax, ex_m[(size_t)pc * BLOCK_M + r]);
  }
  if (ok && scanning) {
    const uint32_t taddr = tmem_base + ((uint32_t)(lane_group * 32) << 16);
    const float thr = amax - 2.f * a.eps;
    // pass 2: the rows inside the band (the accumulator is read again: 16 cycles per 32 columns)
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (g < groups && g * 32 < valid) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(g * 32), v);
        unsigned mk = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) mk |= (__uint_as_float(v[i]) >= thr ? 1u : 0u) << i;
        const int vg = valid - g * 32;
        if (vg < 32) mk &= (1u << vg) - 1u;
        mask[g] = live ? mk : 0u;
      }
    }
  }
  MST_STAMP(3);
  if (ok && !a.stream) {
    // ---- exact score of every thread's FIRST candidate (for all but near-duplicate galleries: its only one) from
    // shared memory: the 128 query rows and the 64 gallery rows are staged with coalesced cp.async (row pitch odd:
    // conflict-free 64-bit reads down a row per thread), in one or two K segments.  Straight from global memory the
    // chain is bound by the L1 tag rate -- every lane walks its own two rows: 32 lines per load instruction, 20 us at
    // k = 178 (ncu source view).
    // The gallery piece (64 rows x k, contiguous in the prepared gallery) and the query rows (contiguous in the copy
    // mst_query_kernel wrote with an odd pitch: conflict-free 64-bit reads down a row per thread) arrive as ONE bulk
    // copy each; when 128 query rows do not fit next to the piece they come in rounds of 64 / 32 rows.  (One bulk copy
    // per row costs ~60 cycles of issue each: 192 copies = 10.8 k cycles at k = 50 -- EF_MST_TRACE.)
    double* Gs = reinterpret_cast<double*>(smem);
    double* Ps = Gs + (((size_t)a.bnp * a.k + 1) & ~(size_t)1);
    int first = -1;
#pragma unroll
    for (int g = 7; g >= 0; --g)
      if (mask[g]) first = g * 32 + __ffs((int)mask[g]) - 1;
    const int rows_q = min(BLOCK_M, a.B - qt * BLOCK_M);
    // the norms of the quotient: requested before the staging so that they are there when the chain ends
    const double pn = first >= 0 ? a.pn[q] : 1.0;
    const double gn_first = (first >= 0 && a.metric == EF_METRIC_COSINE_G1) ? a.gnorm[row_base + first] : 1.0;
    double acc = 0.0;
    uint32_t parity = 1;                                           // full_bar's first phase brought the operand tiles
    for (int r0 = 0; r0 < rows_q; r0 += a.rows_round) {
      const int nr = min(a.rows_round, rows_q - r0);
      if (r0 > 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                           // the previous round has been consumed
      }
      if (a.bulk) {
        if (tid == 0) {
          const uint32_t gbytes = r0 == 0 ? (uint32_t)(valid * a.k * 8) : 0u;
          const uint32_t pbytes = (uint32_t)(((nr + 1) & ~1) * a.lde * 8);   // (odd pitch: an even number of rows)
          mbar_arrive_expect_tx(&sh->full_bar, gbytes + pbytes);
          if (gbytes) bulk_load(Gs, a.G + (int64_t)row_base * a.ldg, gbytes, &sh->full_bar);
          bulk_load(Ps, a.Pe + (int64_t)(qt * BLOCK_M + r0) * a.lde, pbytes, &sh->full_bar);
        }
        mbar_wait(&sh->full_bar, parity, failed);
        parity ^= 1;
      } else {
        for (int rr = warp; rr < nr + (r0 == 0 ? valid : 0); rr += kThreads / 32) {
          const bool is_q = rr < nr;
          const int lr = is_q ? rr : rr - nr;
          const double* src = is_q ? a.Pe + (int64_t)(qt * BLOCK_M + r0 + lr) * a.lde
                                   : a.G + (int64_t)(row_base + lr) * a.ldg;
          const uint32_t dst = smem_u32(is_q ? Ps + (size_t)lr * a.lde : Gs + (size_t)lr * a.k);
          for (int cc = lane; cc < a.k; cc += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * cc), "l"(src + cc) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
      }
      MST_STAMP(4);
      if (first >= 0 && r >= r0 && r < r0 + nr) {
        const double* __restrict__ ps = Ps + (size_t)(r - r0) * a.lde;
        const double* __restrict__ gs = Gs + (size_t)first * a.k;
        if (a.metric == EF_METRIC_L2) {
#pragma unroll 8
          for (int c = 0; c < a.k; ++c) {
            const double d = ps[c] - gs[c];
            acc = fma(d, d, acc);
          }
        } else {
#pragma unroll 8
          for (int c = 0; c < a.k; ++c) acc = fma(ps[c], gs[c], acc);
        }
      }
    }
    MST_STAMP(5);
    if (first >= 0) {
      const int j = row_base + first;
      double s = acc;
      if (a.metric == EF_METRIC_COSINE_G1) s = (pn == 0.0 || gn_first == 0.0) ? 0.0 : acc / (pn * gn_first);
      best = s;
      best_i = j;
      // further rows inside the band (duplicates, near-duplicates, an all-zero query): straight from global memory
      const double* __restrict__ p = a.Pe + (int64_t)q * a.lde;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        unsigned mk = mask[g];
        if (g == (first >> 5)) mk &= mk - 1u;                      // (the first candidate is the lowest bit set)
        while (mk) {                                               // ascending rows: the first best wins
          const int i = g * 32 + __ffs((int)mk) - 1;
          mk &= mk - 1u;
          const double s2 = exact_score(a, p, pn, row_base + i);
          if (better(a.metric, s2, row_base + i, best, best_i)) {
            best = s2;
            best_i = row_base + i;
          }
        }
      }
    }
  }
  if (a.stream) {
    // ---- k > 191: the rows of a (query, candidate) pair do not fit shared memory for all 128 queries, and a chain that
    // reads them from global memory waits for L2 once per sixteen steps (30 k cycles per chain, EF_MST_TRACE).  After
    // the exchange of the maxima a CTA holds ~128 / pieces candidates: they are listed in shared memory and scored in
    // rounds of `slots` pairs -- ALL threads copy the two rows of every pair of the round into the dead ring
    // (coalesced 8-byte cp.async), then one thread per pair runs the chain from shared memory.
    __shared__ int list_n;
    double* list_s = reinterpret_cast<double*>(smem + a.list_off);  // [kListCap] scores, then the pairs
    short* list_q = reinterpret_cast<short*>(list_s + kListCap);
    short* list_j = list_q + kListCap;
    if (tid == 0) list_n = 0;
    __syncthreads();
    bool overflow = false;
    if (ok && live) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        unsigned mk = mask[g];
        while (mk) {
          const int i = g * 32 + __ffs((int)mk) - 1;
          mk &= mk - 1u;
          const int e = atomicAdd(&list_n, 1);
          if (e < kListCap) { list_q[e] = (short)r; list_j[e] = (short)i; }
          else overflow = true;
        }
      }
    }
    __syncthreads();
    const int n_list = min(list_n, kListCap);
    const int pitch = a.k | 1;                                     // doubles per staged row (odd: no bank conflicts)
    double* rows = reinterpret_cast<double*>(smem);                // [slots][2][pitch]
    for (int e0 = 0; e0 < n_list; e0 += a.slots) {
      const int ne = min(a.slots, n_list - e0);
      for (int w = warp; w < 2 * ne; w += kThreads / 32) {         // a warp per row: coalesced
        const int e = e0 + (w >> 1);
        const double* src = (w & 1) ? a.G + (int64_t)(row_base + list_j[e]) * a.ldg
                                    : a.Pe + (int64_t)(qt * BLOCK_M + list_q[e]) * a.lde;
        const uint32_t dst = smem_u32(rows + (size_t)w * pitch);
        for (int c = lane; c < a.k; c += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 8u * c), "l"(src + c) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncthreads();
      if (tid < ne) {
        const int e = e0 + tid;
        const double* __restrict__ ps = rows + (size_t)(2 * tid) * pitch;
        const double* __restrict__ gs = ps + pitch;
        double acc = 0.0;
        if (a.metric == EF_METRIC_L2) {
#pragma unroll 8
          for (int c = 0; c < a.k; ++c) {
            const double d = ps[c] - gs[c];
            acc = fma(d, d, acc);
          }
        } else {
#pragma unroll 8
          for (int c = 0; c < a.k; ++c) acc = fma(ps[c], gs[c], acc);
        }
        if (a.metric == EF_METRIC_COSINE_G1) {
          const double pn = a.pn[qt * BLOCK_M + list_q[e]], gn = a.gnorm[row_base + list_j[e]];
          acc = (pn == 0.0 || gn == 0.0) ? 0.0 : acc / (pn * gn);
        }
        list_s[e] = acc;
      }
      __syncthreads();
    }
    if (ok && live) {
      for (int e = 0; e < n_list; ++e)
        if (list_q[e] == r) {
          const int j = row_base + list_j[e];
          if (better(a.metric, list_s[e], j, best, best_i)) {
            best = list_s[e];
            best_i = j;
          }
        }
      if (overflow) {                                              // (degenerate: hundreds of rows inside the band)
        const double* __restrict__ p = a.Pe + (int64_t)q * a.lde;
        const double pn = a.pn[q];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          unsigned mk = mask[g];
          while (mk) {
            const int i = g * 32 + __ffs((int)mk) - 1;
            mk &= mk - 1u;
            const double s2 = exact_score(a, p, pn, row_base + i);
            if (better(a.metric, s2, row_base + i, best, best_i)) {
              best = s2;
              best_i = row_base + i;
            }
          }
        }
      }
    }
  }
  MST_STAMP(6);
  if (a.cluster) {
    // ---- the per-piece winners meet in the shared memory of the cluster's first CTA (DSMEM stores, one cluster
    // barrier) instead of global memory + a counter (two fences, an atomic round trip, L2 reads: 4-5 k cycles)
    const uint32_t rank = cluster_ctarank();
    double* ex_s = reinterpret_cast<double*>(smem + a.ex_off);
    int* ex_i = reinterpret_cast<int*>(ex_s + (size_t)a.pieces * BLOCK_M);
    if (scanning && rank != 0) {
      st_cluster_u64(map_to_cta(smem_u32(ex_s + (size_t)rank * BLOCK_M + r), 0),
                     (unsigned long long)__double_as_longlong(best));
      st_cluster_u32(map_to_cta(smem_u32(ex_i + (size_t)rank * BLOCK_M + r), 0), (uint32_t)best_i);
    }
    cluster_arrive();
    cluster_wait();
    if (rank != 0) best_i = -2;                                    // the first CTA writes the results
    else if (live) {
      for (int pc = 1; pc < a.pieces; ++pc) {
        const double s = ex_s[(size_t)pc * BLOCK_M + r];
        const int i = ex_i[(size_t)pc * BLOCK_M + r];
        if (i != INT_MAX && better(a.metric, s, i, best, best_i)) {
          best = s;
          best_i = i;
        }
      }
    }
  }
  if (scanning) {
    if (a.pieces > 1 && !a.cluster) {
      if (live) {
        a.part_s[(size_t)piece * a.b_pad + q] = best;
        a.part_i[(size_t)piece * a.b_pad + q] = best_i;
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        const unsigned int done = atomicAdd(a.counters + qt, 1u);
        const int last = done == (unsigned)(a.pieces - 1);
        if (last) a.counters[qt] = 0u;                             // ready for the next launch
        sh->is_last = last;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!sh->is_last) best_i = -2;                               // not this CTA's job
      else {
        __threadfence();
        if (live) {
          best = (a.metric == EF_METRIC_L2) ? CUDART_INF : -CUDART_INF;
          best_i = INT_MAX;
          for (int pc = 0; pc < a.pieces; ++pc) {
            const double s = __ldcg(a.part_s + (size_t)pc * a.b_pad + q);
            const int i = __ldcg(a.part_i + (size_t)pc * a.b_pad + q);
            if (i != INT_MAX && better(a.metric, s, i, best, best_i)) {
              best = s;
              best_i = i;
            }
          }
        }
      }
    }
    if (live && best_i != -2) {
      const bool found = best_i != INT_MAX;
      a.out_score[q] = best;
      a.out_index[q] = found ? best_i : -1;
      if (a.out_label) {
        const bool pass = found && (a.metric == EF_METRIC_L2 ? best <= a.threshold : best >= a.threshold);
        a.out_label[q] = pass ? (a.labels ? a.labels[best_i] : best_i) : -1;
      }
    }
  }

  MST_STAMP(7);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)a.bnp) : "memory");
  }
  if (tid == 0 && sh->failed) atomicExch(a.status, 1);
}

int ka_for(int k, int metric) { return metric == EF_METRIC_L2 ? k + 1 : k; }
int n_slabs_for(int k, int metric) { return (int)ef::ceil_div(3 * (int64_t)ka_for(k, metric), kSlab); }

struct MstLayout {
  size_t counters, part_s, part_i, pn, phat, qimg, total;
};

MstLayout mst_layout(int B, int64_t n, int k, int metric) {
  MstLayout L{};
  const size_t b_pad = (size_t)ef::round_up(std::max(B, 1), BLOCK_M);
  const size_t q_tiles = b_pad / BLOCK_M;
  const size_t pieces = (size_t)ef::ceil_div(n, BNP);
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  L.counters = take(sizeof(unsigned int) * q_tiles);
  L.qimg = take(q_tiles * (size_t)n_slabs_for(k, metric) * kSlabBytesA);
  L.part_s = take(sizeof(double) * pieces * b_pad);
  L.part_i = take(sizeof(int) * pieces * b_pad);
  L.pn = take(sizeof(double) * b_pad);
  L.phat = take(sizeof(double) * b_pad * (size_t)(k | 1) + 16);
  L.total = off;
  return L;
}

}  // namespace

namespace ef {

bool match_small_tc_supported(int k, int64_t n, int metric) {
  if (k <= 0 || n <= 0 || n > kMaxRows) return false;
  if (metric < EF_METRIC_COSINE_SK || metric > EF_METRIC_L2) return false;
  return n_slabs_for(k, metric) <= kMaxSlabsStream;
}

size_t match_small_tc_image_bytes(int k, int64_t n, int metric) { return ef_match_tc_image_bytes_metric(n, k, metric); }

int match_small_tc_image(const double* gp, int64_t ldgp, const double* gnorm, int64_t n, int k, int metric, void* image,
                         cudaStream_t stream) {
  return ef_match_tc_prepare_device(gp, ldgp, gnorm, n, k, metric, image, reinterpret_cast<ef_stream_t>(stream));
}

// scratch (256-byte aligned) for batches of up to cap_B crops: query-tile counters, the float16 query image, per-piece
// winners, |p| and the rows of the exact chain.  ALL of it must be zero before the first launch (counters; K padding of
// the image) and every launch must be given the SAME cap_B (the layout depends on it, not on the batch): the kernels
// then keep the counters and the padding zero.
size_t match_small_tc_work_bytes(int cap_B, int64_t n, int k, int metric) { return mst_layout(cap_B, n, k, metric).total; }

int match_small_tc(double* proj, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm,
                   const void* image, int64_t n, const int32_t* labels, int metric, double threshold, double* sumsq,
                   double c0, double* resid2, double* out_score, int32_t* out_index, int32_t* out_label, void* work,
                   int cap_B, int* status, cudaStream_t stream, const MatchSmallTcSlabs* slabs) {
  if (B <= 0) return EF_OK;
  if (B > cap_B) return EF_ERR_INVALID;
  if (slabs && (slabs->S != 8 || !slabs->combined || (slabs->ld_part & 3) || slabs->kq > 224 || slabs->kq < k ||
                n_slabs_for(k, metric) > kMaxSlabs))
    return EF_ERR_INVALID;
  if (!match_small_tc_supported(k, n, metric) || !image || !work || !status) return EF_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(work) & 255) return EF_ERR_INVALID;
  const MstLayout L = mst_layout(cap_B, n, k, metric);
  char* w = reinterpret_cast<char*>(work);
  MstArgs a{};
  a.B = B; a.k = k; a.ka = ka_for(k, metric); a.n_slabs = n_slabs_for(k, metric);
  a.metric = metric;
  a.img = reinterpret_cast<const uint8_t*>(image);
  a.gscale = reinterpret_cast<const double*>(a.img + (size_t)ceil_div(n, kTileRows) * a.n_slabs * kTileSlabBytes);
  a.G = gp; a.ldg = ldgp; a.gnorm = gnorm; a.n = (int)n;
  a.labels = labels; a.threshold = threshold;
  a.out_score = out_score; a.out_index = out_index; a.out_label = out_label;
  a.b_pad = (int)round_up(B, BLOCK_M);
  a.counters = reinterpret_cast<unsigned int*>(w + L.counters);
  a.part_s = reinterpret_cast<double*>(w + L.part_s);
  a.part_i = reinterpret_cast<int*>(w + L.part_i);
  a.pn = reinterpret_cast<const double*>(w + L.pn);
  a.qimg = reinterpret_cast<const uint8_t*>(w + L.qimg);
  a.Pe = reinterpret_cast<const double*>(w + L.phat);
  a.lde = k | 1;
  a.status = status;
  // |approximate key - exact key|: 3 ka float16 products accumulated in float32 (see ef_match_tc.cu)
  const bool deep = 3 * a.ka > 384;
  a.eps = metric == EF_METRIC_L2 ? (deep ? 6e-4f : 4e-4f) : (deep ? 3e-4f : 2e-4f);
  a.stream = a.n_slabs > kMaxSlabs ? 1 : 0;
  if (a.stream) {                                     // 3 ka products of <= 2^-22 each + the split error, 25 % margin
    const float e = 3.f * (float)a.ka * 3e-7f + 2e-5f;
    a.eps = metric == EF_METRIC_L2 ? 2.f * e : e;
  }

  MstQueryArgs qa{};
  if (slabs) {
    qa.part = reinterpret_cast<const long long*>(slabs->part); qa.splits = slabs->splits;
    qa.ld_part = slabs->ld_part / 4; qa.kq = slabs->kq;
    qa.slab_stride = (long long)B * (slabs->ld_part / 4);
    qa.col_exp = slabs->col_exp; qa.bias = slabs->bias;
  }
  qa.P = proj; qa.ldp = ldp; qa.B = B; qa.b_pad = a.b_pad; qa.k = k; qa.ka = a.ka; qa.n_slabs = a.n_slabs;
  qa.metric = metric; qa.gscale = a.gscale;
  qa.sumsq = resid2 ? sumsq : nullptr; qa.c0 = c0; qa.resid2 = resid2;
  qa.qimg = reinterpret_cast<uint8_t*>(w + L.qimg);
  qa.pn = reinterpret_cast<double*>(w + L.pn);
  qa.pe = reinterpret_cast<double*>(w + L.phat);
  qa.lde = k | 1;
  if (a.stream)
    EF_LAUNCH_PDL(mst_query_big_kernel, (unsigned)ceil_div((int64_t)a.b_pad * 32, 256), 256, 0, stream, qa);
  else
    EF_LAUNCH_PDL(mst_query_kernel, (unsigned)ceil_div((int64_t)a.b_pad * 32, 256), 256, 0, stream, qa);

  // float64 staging area of the exact chain (reuses the operand tiles): all of k when (128 + 64) rows fit 200 KB
  // Bulk copies of the float64 rows need 16-byte aligned blocks of a multiple of 16 bytes: k even, a packed gallery
  a.bulk = (k % 2 == 0) && ldgp == k && ((reinterpret_cast<uintptr_t>(gp) & 15) == 0) && !getenv("EF_MST_NO_BULK");
  // Gallery rows per CTA: the smallest piece whose grid is one wave of resident CTAs (more, smaller pieces = more SMs at
  // work), else the largest that fits shared memory (fewer pieces = fewer redundant copies of the query rows, fewer
  // exact chains).  Shared memory holds the operand tiles, then (reusing them) the float64 rows of the exact chain:
  // the gallery piece + as many query rows as fit.
  const int q_tiles = a.b_pad / BLOCK_M;
  size_t operands = 0, staging = 0;
  a.bnp = 0;
  int forced = 0;
  if (const char* e = getenv("EF_MST_BNP")) forced = atoi(e);
  for (int bnp : {64, 128, 256}) {
    size_t ops, stg;
    int rows = BLOCK_M;
    int ring = 0;
    if (a.stream) {
      // two CTAs per SM for the smaller pieces (ring <= ~100 KB), four stages alone on the SM for 256-row pieces
      const size_t stage_bytes = kSlabBytesA + (size_t)bnp * 128;
      ring = bnp == 256 ? 4 : (int)std::min<size_t>(kRing, (100 * 1024) / stage_bytes);
      ops = (size_t)ring * stage_bytes;
      stg = 0;
    } else {
      ops = (size_t)a.n_slabs * (kSlabBytesA + (size_t)bnp * 128);
      const size_t g_bytes = (((size_t)bnp * k + 1) & ~(size_t)1) * sizeof(double);
      while (rows > 8 && g_bytes + (size_t)rows * a.lde * sizeof(double) > 200 * 1024) rows /= 2;
      stg = g_bytes + (size_t)rows * a.lde * sizeof(double);
    }
    if (std::max(ops, stg) > 216 * 1024) break;
    a.bnp = bnp; a.rows_round = rows; a.ring = ring; operands = ops; staging = stg;
    if (forced == bnp) break;
    const int64_t occ = std::max<int64_t>(1, std::min<int64_t>(512 / bnp, (220 * 1024) / (int64_t)(std::max(ops, stg) + 2048)));
    if (!forced && (int64_t)q_tiles * ceil_div(n, bnp) <= (int64_t)sm_count() * occ) break;
  }
  if (!a.bnp) return EF_ERR_UNSUPPORTED;
  if (a.stream) {
    a.slots = (int)std::min<size_t>(kThreads, operands / (2 * (size_t)(k | 1) * sizeof(double)));
    if (a.slots < 1) return EF_ERR_UNSUPPORTED;
  }
  a.pieces = (int)ceil_div(n, a.bnp);
  size_t smem = std::max(operands, (staging + 15) & ~(size_t)15) + sizeof(MstShared) + 64;
  a.sh_off = (int)std::max(operands, (staging + 15) & ~(size_t)15);
  // the CTAs of a query tile as one cluster (portable size, exchange area fits): winners meet through DSMEM
  const size_t ex_bytes = (size_t)a.pieces * BLOCK_M * (sizeof(double) + sizeof(int) + sizeof(float));
  a.cluster = a.pieces > 1 && a.pieces <= 8 && smem + ex_bytes <= (size_t)kSmemLimit && !getenv("EF_MST_NO_CLUSTER");
  if (a.cluster) {
    a.ex_off = (int)((smem + 15) & ~(size_t)15);
    a.exm_off = a.ex_off + (int)((size_t)a.pieces * BLOCK_M * (sizeof(double) + sizeof(int)));
    smem = (size_t)a.ex_off + ex_bytes;
  }
  if (a.stream) {
    a.list_off = (int)((smem + 15) & ~(size_t)15);
    smem = (size_t)a.list_off + (size_t)kListCap * (sizeof(double) + 2 * sizeof(short));
    if (smem > (size_t)kSmemLimit) return EF_ERR_UNSUPPORTED;
  }
  EF_ENSURE_SMEM(match_small_tc_kernel, smem);
  const dim3 grid((unsigned)q_tiles, (unsigned)a.pieces);
  static long long* trace_buf = nullptr;
  const bool trace = getenv("EF_MST_TRACE") != nullptr;
  const size_t n_cta = (size_t)grid.x * grid.y;
  if (trace) {
    if (!trace_buf) EF_CUDA(cudaMalloc(&trace_buf, sizeof(long long) * 8 * 65536));
    if (n_cta <= 65536) a.trace = trace_buf;
  }
  if (a.cluster) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = 1;
    attrs[0].val.clusterDim.y = (unsigned)a.pieces;
    attrs[0].val.clusterDim.z = 1;
    attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = getenv("EF_NO_PDL") ? 1 : 2;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, match_small_tc_kernel, a);
    ef::g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) { ef::set_error_detail("match_small_tc_kernel (cluster)", e); return EF_ERR_CUDA; }
  } else {
    EF_LAUNCH_PDL(match_small_tc_kernel, grid, kThreads, smem, stream, a);
  }
  if (a.trace) {
    EF_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> h(n_cta * 8);
    EF_CUDA(cudaMemcpy(h.data(), trace_buf, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    double d[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t c = 0; c < n_cta; ++c)
      for (int i = 0; i < 7; ++i) d[i] += (double)(h[c * 8 + i + 1] - h[c * 8 + i]) / n_cta;
    fprintf(stderr, "mst trace (cycles, mean over %zu CTAs): init+issue %.0f | loads+mma %.0f | scan %.0f | stage %.0f | "
            "chain %.0f | extra candidates %.0f | combine %.0f\n", n_cta, d[0], d[1], d[2], d[3], d[4], d[5], d[6]);
  }
  return EF_OK;
}

}  // namespace ef
