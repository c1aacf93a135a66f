// K1: crop preprocess, bit-exact with OpenCV's BGR2GRAY + INTER_LINEAR resize on uint8.
//
// Replaces cv2.cvtColor + cv2.resize + .flatten() at scan-template-v4.py:257-263, train-v5.py:329-332 and
// useless/scan.py:248-252 for a whole batch of detection boxes.  HBM-bound byte work: one CTA walks crops;
// per crop the coefficient tables live in shared memory, the source rows a band of output rows needs are STAGED in
// shared memory as gray bytes (16-byte coalesced loads for gray frames, fused BGR->gray for colour frames) so that
// every source byte is read from HBM/L2 once, the four taps of an output pixel come from shared memory, the x-axis
// taps / weights of a thread's column stay in registers, and the destination row is written coalesced.
//
// Fixed-point spec (the same one oracle/preprocess.py states and tests pin against cv2):
//   gray = (3735 B + 19235 G + 9798 R + 2^14) >> 15
//   scale = 1 / (dst / src) (double); f = float((d + .5) scale - .5); s = floor f; f -= s
//   x: clamp s into [0, w-1] zeroing f when clamped; y: weights from the unclamped f, rows clipped
//   a = rint(f * 2048) (float32, half-even); H = s0 * a0 + s1 * a1
//   out = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
//   src == 2 dst on both axes -> 2x2 box (a + b + c + d + 2) >> 2 ; src == dst -> copy
#include <algorithm>
#include <cstdlib>

#include "ef_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxSide = 1024;  // largest dw / dh the shared-memory tables are sized for

__device__ __forceinline__ int gray_at(const uint8_t* __restrict__ row, int x, int channels) {
  if (channels == 1) return row[x];
  const uint8_t* p = row + 3 * x;
  return (3735 * p[0] + 19235 * p[1] + 9798 * p[2] + (1 << 14)) >> 15;
}

// One axis of OpenCV's coefficient table, computed with explicitly rounded (never fused) operations.
__device__ __forceinline__ double axis_scale(int src, int dst) {
  const double inv_scale = __ddiv_rn((double)dst, (double)src);
  return __ddiv_rn(1.0, inv_scale);
}
__device__ __forceinline__ void axis_coeff(int d, int src, double scale, bool clamp_frac, int& s0, int& s1, int& w0,
                                           int& w1) {
  const double fd = __dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
  float f = __double2float_rn(fd);
  const float fl = floorf(f);
  int s = (int)fl;
  f = __fsub_rn(f, fl);
  if (clamp_frac) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
  s0 = min(max(s, 0), src - 1);
  s1 = min(max(s + 1, 0), src - 1);
}

__global__ void __launch_bounds__(kThreads, 4)
preprocess_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride, int pitch, int width, int height,
                  int channels, int n_frames, const ef_box_t* __restrict__ boxes, int n_boxes, int dw, int dh,
                  uint8_t* __restrict__ out, int64_t out_stride, int* __restrict__ bad_boxes, int stage_bytes,
                  unsigned int* __restrict__ next_box, int debug_flags) {
  extern __shared__ __align__(16) int tab[];
  __shared__ double scales[2];
  int* xs0 = tab;
  int* xs1 = xs0 + dw;
  int* xa0 = xs1 + dw;
  int* xa1 = xa0 + dw;
  int* ys0 = xa1 + dw;
  int* ys1 = ys0 + dh;
  int* yb0 = ys1 + dh;
  int* yb1 = yb0 + dh;
  uint8_t* stage = reinterpret_cast<uint8_t*>(yb1 + dh);   // byte offset 16 (dw + dh): 16-byte aligned
  const int tid = threadIdx.x;
  const int npix = dw * dh;

  // dynamic schedule: crops differ in size by up to 9x, so every CTA takes the next unprocessed box when it is free
  __shared__ int next_s;
  for (;;) {
    __syncthreads();                                 // previous crop finished with the tables / staging / next_s
    if (tid == 0) next_s = (int)atomicAdd(next_box, 1u);
    __syncthreads();
    const int b = next_s;
    if (b >= n_boxes) break;
    const ef_box_t box = boxes[b];
    uint8_t* __restrict__ o = out + (int64_t)b * out_stride;
    const bool ok = box.frame >= 0 && box.frame < n_frames && box.w > 0 && box.h > 0 && box.x >= 0 && box.y >= 0 &&
                    box.x + box.w <= width && box.y + box.h <= height;
    if (!ok) {
      for (int i = tid; i < npix; i += kThreads) o[i] = 0;
      if (tid == 0 && bad_boxes) atomicAdd(bad_boxes, 1);
      continue;
    }
    const uint8_t* __restrict__ src =
        frames + (int64_t)box.frame * frame_stride + (int64_t)box.y * pitch + (int64_t)box.x * channels;
    const int w = box.w, h = box.h;

    if (w == dw && h == dh) {
      // cv2.resize returns a copy when the size already matches.
      const bool vec = channels == 1 && (dw % 16 == 0) && (pitch % 16 == 0) &&
                       ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
      if (vec) {
        const int per_row = dw / 16;
        for (int i = tid; i < per_row * dh; i += kThreads) {
          const int y = i / per_row, xv = i - y * per_row;
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (int64_t)y * pitch) + xv);
          reinterpret_cast<uint4*>(o + (int64_t)y * dw)[xv] = v;
        }
      } else {
        for (int i = tid; i < npix; i += kThreads) {
          const int y = i / dw, x = i - y * dw;
          o[i] = (uint8_t)gray_at(src + (int64_t)y * pitch, x, channels);
        }
      }
    } else if (w == 2 * dw && h == 2 * dh) {
      // OpenCV switches INTER_LINEAR to INTER_AREA for an exact 2x decimation.
      for (int i = tid; i < npix; i += kThreads) {
        const int y = i / dw, x = i - y * dw;
        const uint8_t* r0 = src + (int64_t)(2 * y) * pitch;
        const uint8_t* r1 = r0 + pitch;
        const int s = gray_at(r0, 2 * x, channels) + gray_at(r0, 2 * x + 1, channels) +
                      gray_at(r1, 2 * x, channels) + gray_at(r1, 2 * x + 1, channels);
        o[i] = (uint8_t)((s + 2) >> 2);
      }
    } else {
      // the two double-precision divisions of an axis scale are done once per crop (threads 0 and 32), not per entry
      if (tid == 0) scales[0] = axis_scale(w, dw);
      if (tid == 32) scales[1] = axis_scale(h, dh);
      __syncthreads();
      {
        const double sx = scales[0], sy = scales[1];
        for (int d = tid; d < dw + dh; d += kThreads) {
          if (d < dw) {
            axis_coeff(d, w, sx, true, xs0[d], xs1[d], xa0[d], xa1[d]);
          } else {
            const int e = d - dw;
            axis_coeff(e, h, sy, false, ys0[e], ys1[e], yb0[e], yb1[e]);
          }
        }
      }
      __syncthreads();
      // shared-memory row pitch of the staged gray rows; gray frames keep the 16-byte phase of the source address so
      // that the staging loads are aligned uint4 (pixel x of a row sits at column `phase + x`)
      const bool vec = channels == 1 && (pitch % 16 == 0) && ((reinterpret_cast<uintptr_t>(frames) & 15) == 0) &&
                       (frame_stride % 16 == 0);
      const int phase = vec ? (int)((reinterpret_cast<uintptr_t>(src)) & 15) : 0;
      // colour frames: every row of the ROI starts at the same offset inside a 32-bit word when the pitches are multiples
      // of four -- the staging then works on aligned words
      const bool bgr_vec = !(debug_flags & 2) && channels == 3 && (pitch % 4 == 0) && (frame_stride % 4 == 0) &&
                           ((reinterpret_cast<uintptr_t>(frames) & 3) == 0);
      const int bgr_align = bgr_vec ? (int)(reinterpret_cast<uintptr_t>(src) & 3) : 0;
      // end of the last frame's rows (frame_stride may be 0 for a single frame: a size-1 axis has an arbitrary stride)
      const uint8_t* frames_end = frames + (int64_t)(n_frames - 1) * frame_stride + (int64_t)height * pitch;
      const int wp = ((phase + w + 15) & ~15);
      const int rows_fit = stage_bytes / wp;
      if (rows_fit < 2 || dw > kThreads) {
        // ROI row too wide for the staging buffer (or very wide output): taps straight from global memory
        for (int i = tid; i < npix; i += kThreads) {
          const int y = i / dw, x = i - y * dw;
          const uint8_t* r0 = src + (int64_t)ys0[y] * pitch;
          const uint8_t* r1 = src + (int64_t)ys1[y] * pitch;
          const int x0 = xs0[x], x1 = xs1[x], a0 = xa0[x], a1 = xa1[x];
          const int h0 = gray_at(r0, x0, channels) * a0 + gray_at(r0, x1, channels) * a1;
          const int h1 = gray_at(r1, x0, channels) * a0 + gray_at(r1, x1, channels) * a1;
          const int v = (((yb0[y] * (h0 >> 4)) >> 16) + ((yb1[y] * (h1 >> 4)) >> 16) + 2) >> 2;
          o[i] = (uint8_t)min(max(v, 0), 255);
        }
      } else {
        // this thread's FOUR adjacent output columns and its row phase; the x-axis taps and weights stay in registers
        // (dw % 4 == 0 and 4-byte aligned output rows: one 32-bit store per thread and row; otherwise one column each)
        const bool quad = (dw % 4 == 0) && (out_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0);
        const int cols = quad ? dw / 4 : dw;               // work items per output row
        const int R = kThreads / cols;                     // output rows in flight per pass
        const int ph = tid / cols, cx = tid - ph * cols;
        const bool worker = ph < R;
        int tx0[4], tx1[4], ta0[4], ta1[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int dx = quad ? 4 * cx + j : cx;
          tx0[j] = worker ? phase + xs0[dx] : 0;
          tx1[j] = worker ? phase + xs1[dx] : 0;
          ta0[j] = worker ? xa0[dx] : 0;
          ta1[j] = worker ? xa1[dx] : 0;
        }
        int dy0 = 0;
        while (dy0 < dh) {
          // band of output rows [dy0, dy1) whose source rows [r_lo, r_hi] fit the staging buffer
          const int r_lo = ys0[dy0];
          int dy1 = dh;                                     // common case: everything that is left fits
          if (ys1[dh - 1] - r_lo + 1 > rows_fit) {
            int lo = dy0 + 1, hi = dh;                      // largest dy1 with ys1[dy1 - 1] - r_lo + 1 <= rows_fit
            while (lo < hi) {
              const int mid = (lo + hi + 1) >> 1;
              if (ys1[mid - 1] - r_lo + 1 <= rows_fit) lo = mid; else hi = mid - 1;
            }
            dy1 = lo;
          }
          const int r_hi = ys1[dy1 - 1];
          const int n_rows = r_hi - r_lo + 1;
          if (vec) {
            const int vec_per_row = wp >> 4;
            const uint8_t* base = src - phase + (int64_t)r_lo * pitch;          // 16-byte aligned
            // 8 / 16 / 32 lanes per source row (power of two: shifts, no divisions)
            const int lpr_log2 = vec_per_row <= 8 ? 3 : (vec_per_row <= 16 ? 4 : 5);
            const int lane_v = tid & ((1 << lpr_log2) - 1), row_slot = tid >> lpr_log2, slots = kThreads >> lpr_log2;
            // cp.async: every thread's copies are in flight together (one memory latency per band, not one per row).
            // Pointers advance by a row-slot stride: the loop body is the copy, two adds and the branch (the
            // per-row 64-bit multiply-adds of the first version were 29 % of the kernel's instructions)
            if (vec_per_row <= (1 << lpr_log2)) {
              if (lane_v < vec_per_row) {
                const uint8_t* g = base + (int64_t)row_slot * pitch + 16 * lane_v;
                unsigned sm = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)(row_slot * wp + 16 * lane_v);
                const int64_t g_step = (int64_t)slots * pitch;
                const unsigned sm_step = (unsigned)(slots * wp);
#pragma unroll 4
                for (int r = row_slot; r < n_rows; r += slots, g += g_step, sm += sm_step)
                  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sm), "l"(g));
              }
            } else {
              for (int r = row_slot; r < n_rows; r += slots) {
                const uint8_t* g = base + (int64_t)r * pitch;
                const unsigned sm = (unsigned)__cvta_generic_to_shared(stage + r * wp);
                for (int v = lane_v; v < vec_per_row; v += 1 << lpr_log2)
                  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sm + 16u * v), "l"(g + 16 * v));
              }
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::);
          } else if (bgr_vec) {
            // BGR -> gray fused into the staging, four pixels (12 bytes) per thread and step: four aligned 32-bit loads,
            // three funnel shifts to the pixel boundary, one byte permute per pixel to [B, G, R, 0], then the luma as two
            // dp2a (16-bit coefficient x 8-bit channel) -- 27 instructions per four pixels instead of ~38
            const int quads = (w + 3) >> 2;
            const int sh8 = bgr_align * 8;
            for (int r = tid >> 5; r < n_rows; r += kThreads >> 5) {
              const uint8_t* g = src + (int64_t)(r_lo + r) * pitch - bgr_align;      // 4-byte aligned
              for (int qd = tid & 31; qd < quads; qd += 32) {
                const unsigned* wp4 = reinterpret_cast<const unsigned*>(g + 12 * qd);
                unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                // (the last quad of the last row of the last frame may look past the allocation: bounded by frames_end)
                if (reinterpret_cast<const uint8_t*>(wp4) + 4 <= frames_end) w0 = __ldg(wp4);
                if (reinterpret_cast<const uint8_t*>(wp4) + 8 <= frames_end) w1 = __ldg(wp4 + 1);
                if (reinterpret_cast<const uint8_t*>(wp4) + 12 <= frames_end) w2 = __ldg(wp4 + 2);
                if (reinterpret_cast<const uint8_t*>(wp4) + 16 <= frames_end) w3 = __ldg(wp4 + 3);
                const unsigned b0 = __funnelshift_r(w0, w1, sh8), b1 = __funnelshift_r(w1, w2, sh8),
                               b2 = __funnelshift_r(w2, w3, sh8);
                const unsigned p0 = __byte_perm(b0, 0u, 0x4210), p1 = __byte_perm(b0, b1, 0x7543),
                               p2 = __byte_perm(b1, b2, 0x7432), p3 = __byte_perm(b2, 0u, 0x4321);
                const unsigned cbg = 3735u | (19235u << 16), cr = 9798u;
                const unsigned g0 = __dp2a_hi(cr, p0, __dp2a_lo(cbg, p0, 1u << 14)) >> 15;
                const unsigned g1 = __dp2a_hi(cr, p1, __dp2a_lo(cbg, p1, 1u << 14)) >> 15;
                const unsigned g2 = __dp2a_hi(cr, p2, __dp2a_lo(cbg, p2, 1u << 14)) >> 15;
                const unsigned g3 = __dp2a_hi(cr, p3, __dp2a_lo(cbg, p3, 1u << 14)) >> 15;
                *reinterpret_cast<unsigned*>(stage + r * wp + 4 * qd) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
              }
            }
          } else {
            // one warp per source row: coalesced byte reads, BGR -> gray fused into the staging
            for (int r = tid >> 5; r < n_rows; r += kThreads >> 5) {
              const uint8_t* g = src + (int64_t)(r_lo + r) * pitch;
#pragma unroll 4
              for (int x = tid & 31; x < w; x += 32) stage[r * wp + x] = (uint8_t)gray_at(g, x, channels);
            }
          }
          __syncthreads();
          if (worker && quad && w >= dw && h >= dh && !(debug_flags & 1)) {
            // Downscale (or equal on one axis): no tap is ever clamped to a DIFFERENT pixel -- x1 = x0 + 1 and
            // y1 = y0 + 1 except where float rounding pushed s onto the last index, and there OpenCV zeroes the
            // fraction, so the weight of the second tap is exactly 0 (the byte read in its place -- one past the row /
            // band, inside the staging slack -- does not matter).  The four taps of a pixel are then [P], [P + 1],
            // [P + wp], [P + wp + 1]: two address adds per pixel instead of four, and no clamps on the result
            // (<= 255 by construction: the weights of an axis sum to at most 2049).
            for (int y = dy0 + ph; y < dy1; y += R) {
              const uint8_t* r0 = stage + (ys0[y] - r_lo) * wp;
              const int b0 = yb0[y], b1 = yb1[y];
              unsigned packed = 0;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint8_t* p0 = r0 + tx0[j];
                const uint8_t* p1 = p0 + wp;
                const int h0 = p0[0] * ta0[j] + p0[1] * ta1[j];
                const int h1 = p1[0] * ta0[j] + p1[1] * ta1[j];
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                packed |= (unsigned)v << (8 * j);
              }
              *reinterpret_cast<unsigned*>(o + y * dw + 4 * cx) = packed;
            }
          } else if (worker) {
            for (int y = dy0 + ph; y < dy1; y += R) {
              const uint8_t* r0 = stage + (ys0[y] - r_lo) * wp;
              const uint8_t* r1 = stage + (ys1[y] - r_lo) * wp;
              const int b0 = yb0[y], b1 = yb1[y];
              if (quad) {
                unsigned packed = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int h0 = r0[tx0[j]] * ta0[j] + r0[tx1[j]] * ta1[j];
                  const int h1 = r1[tx0[j]] * ta0[j] + r1[tx1[j]] * ta1[j];
                  const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                  packed |= (unsigned)min(max(v, 0), 255) << (8 * j);
                }
                *reinterpret_cast<unsigned*>(o + y * dw + 4 * cx) = packed;
              } else {
                const int h0 = r0[tx0[0]] * ta0[0] + r0[tx1[0]] * ta1[0];
                const int h1 = r1[tx0[0]] * ta0[0] + r1[tx1[0]] * ta1[0];
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                o[y * dw + cx] = (uint8_t)min(max(v, 0), 255);
              }
            }
          }
          __syncthreads();                                  // the staging buffer is refilled by the next band
          dy0 = dy1;
        }
      }
      __syncthreads();  // tables are rewritten by the next crop
    }
  }
}

}  // namespace

extern "C" int ef_preprocess(const uint8_t* frames, int64_t frame_stride, int32_t pitch, int32_t width,
                             int32_t height, int32_t channels, int32_t n_frames, const ef_box_t* boxes,
                             int32_t n_boxes, int32_t dw, int32_t dh, uint8_t* out, int64_t out_stride,
                             int32_t* bad_boxes, ef_stream_t stream) {
  if (n_boxes == 0) return EF_OK;
  if (!frames || !boxes || !out) return EF_ERR_INVALID;
  if (n_boxes < 0 || width <= 0 || height <= 0 || n_frames <= 0 || dw <= 0 || dh <= 0) return EF_ERR_INVALID;
  if (channels != 1 && channels != 3) return EF_ERR_INVALID;
  if (pitch < width * channels || out_stride < (int64_t)dw * dh) return EF_ERR_INVALID;
  if (dw > kMaxSide || dh > kMaxSide) return EF_ERR_UNSUPPORTED;
  if (n_boxes == 0) return EF_OK;
  // coefficient tables + a 48 KB staging buffer for the source rows of a band: four CTAs per SM (64 registers)
  int kStageBytes = 48 * 1024;                      // four CTAs per SM; a 220 x 220 ROI fits one band
  if (const char* e = getenv("EF_PRE_STAGE_KB")) { const int v = atoi(e); if (v >= 8 && v <= 96) kStageBytes = v * 1024; }
  int debug_flags = 0;                              // EF_PRE_DEBUG: 1 = no downscale fast path, 2 = scalar BGR staging
  if (const char* e = getenv("EF_PRE_DEBUG")) debug_flags = atoi(e);
  const size_t tables = sizeof(int) * (4 * (size_t)(dw + dh) + 4);
  // + slack: the downscale path reads (with weight 0) one byte past a staged row and one row past a staged band
  const size_t smem = tables + kStageBytes + 2048;
  if (EF_FIRST_ON_DEVICE()) {
    EF_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    // four CTAs of 52 KB per SM need the largest shared-memory carve-out (the default heuristic settles for less)
    EF_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared));
  }
  const int64_t per_sm = std::max<int64_t>(1, std::min<int64_t>(8, (220 * 1024) / (int64_t)(smem + 1024)));
  const int64_t resident = (int64_t)ef::sm_count() * per_sm;
  const int grid = (int)(n_boxes < resident ? n_boxes : resident);
  // work counter of this launch: one of 256 slots of a device array owned by the library (concurrent launches on other
  // streams take other slots), zeroed on the launching stream
  static unsigned int* counters_of[64] = {nullptr};
  static std::atomic<unsigned int> next_slot{0};
  int dev = 0;
  EF_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return EF_ERR_UNSUPPORTED;
  if (!counters_of[dev]) EF_CUDA(cudaMalloc(&counters_of[dev], 256 * sizeof(unsigned int)));
  unsigned int* counter = counters_of[dev] + (next_slot.fetch_add(1) & 255u);
  EF_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned int), ef::as_stream(stream)));
  EF_LAUNCH(preprocess_kernel, grid, kThreads, smem, ef::as_stream(stream), frames, frame_stride, pitch, width,
            height, channels, n_frames, boxes, n_boxes, dw, dh, out, out_stride, bad_boxes, kStageBytes, counter,
            debug_flags);
  return EF_OK;
}
