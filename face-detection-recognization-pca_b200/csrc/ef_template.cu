// K6: template-matching detector -- cv2.matchTemplate(frame, template, TM_CCOEFF_NORMED) + cv2.minMaxLoc for a batch of
// (template, scale) jobs against one gray frame.
//
// Replaces the inner loop of MultiModelFaceScanner.template_match_all_models (scan-template-v4.py:129-197): per person,
// per template image, per scale in {0.8, 1.0, 1.2}: matchTemplate (:170) + minMaxLoc (:171).
//
// Arithmetic.  OpenCV forms the cross term with a float32 DFT and the window statistics with integral images
// (modules/imgproc/src/templmatch.cpp, crossCorr + common_matchTemplate).  Here every sum is an exact integer:
//   sum(T I)   dp4a over the window (u8 x u8 -> u32 per 16 template rows, int64 across row blocks)
//   sum(I), sum(I^2)   integral images (int64 / uint64)
//   sum(T), sum(T^2)   per template
// and the score follows common_matchTemplate's own formula in float64 from exact numerators,
//   num = sum(TI) - sum(T) sum(I) / n ;  t = sqrt(max(sum(I^2) - sum(I)^2 / n, 0)) * sqrt(sum(T^2) - sum(T)^2 / n)
//   |num| < t -> num / t ;  |num| < 1.125 t -> +-1 ;  else 0 ;  flat template -> 1 everywhere
// rounded to float32 like the cv2 result map.  The arg-max is the first maximum in row-major order (minMaxLoc).
// The map agrees with cv2 to the accuracy of cv2's float32 DFT (about 1e-5); the tests pin that.
//
// Schedule: one CTA per 64 x 32 tile of result positions of one job; the frame patch and 16 template rows at a time are
// staged in shared memory as 32-bit words; a thread owns 4 consecutive x positions of 2 rows and slides over the template
// row four pixels per step.  The byte shift between neighbouring x positions is put into the TEMPLATE (four copies of
// each row, moved right by 0..3 bytes, one 16-byte broadcast load), so the patch words stay aligned and are shared by
// the four positions: 1 LDS.128 + 2 LDS + 8 dp4a = 32 MAC per 11 instructions, no byte permutes.  Integer/byte work on the CUDA cores: every job has its own K = w h and a single output column, there is
// no GEMM to hand to the tensor cores.
#include <cfloat>
#include <climits>

#include "ef_common.cuh"
#include "ef_internal.cuh"

namespace {

constexpr int kMaxJobs = 64;
constexpr int TX = 64, TY = 32;      // result positions per CTA
constexpr int VC = 16;               // template rows per shared-memory block
constexpr int kThreads = 256;
constexpr int kMaxTemplateW = 1024;  // u32 dp4a accumulators: VC * (w + 3) * 255^2 < 2^32

struct TmJob {
  long long t_off, r_off;            // template bytes offset; result floats offset
  int w, h, rw, rh;                  // template size; result size (W - w + 1, H - h + 1)
  int tiles_x, tile0;                // tiles per result row; first CTA of this job
};

struct TmArgs {
  const uint8_t* frame;
  long long ldf;
  int W, H;
  const uint8_t* tmpl;
  const long long* S1;               // integral images, (H + 1) x (W + 1)
  const unsigned long long* S2;
  const long long* tsum;             // per job sum(T), sum(T^2)
  const long long* tsum2;
  float* result;                     // nullable
  float* cta_val;
  int* cta_pos;
  int n_jobs, max_pitch_words;
  TmJob jobs[kMaxJobs];
};

// ---- integral images: row prefix sums, then column prefix sums in place
__global__ void tm_row_prefix_kernel(const uint8_t* __restrict__ frame, long long ldf, int W, int H,
                                     long long* __restrict__ S1, unsigned long long* __restrict__ S2) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= H) return;
  const uint8_t* src = frame + (long long)row * ldf;
  long long* d1 = S1 + (long long)(row + 1) * (W + 1);
  unsigned long long* d2 = S2 + (long long)(row + 1) * (W + 1);
  long long c1 = 0;
  unsigned long long c2 = 0;
  if (lane == 0) { d1[0] = 0; d2[0] = 0; }
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    const unsigned v = x < W ? src[x] : 0u;
    long long a = v;
    unsigned long long b = v * v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long ta = __shfl_up_sync(0xffffffffu, a, o);
      const unsigned long long tb = __shfl_up_sync(0xffffffffu, b, o);
      if (lane >= o) { a += ta; b += tb; }
    }
    if (x < W) { d1[x + 1] = c1 + a; d2[x + 1] = c2 + b; }
    c1 += __shfl_sync(0xffffffffu, a, 31);
    c2 += __shfl_sync(0xffffffffu, b, 31);
  }
}

// 32 columns per CTA, 32 row segments per column: segment sums, a scan over the segments in shared memory, then every
// thread walks its segment again with the offset (H / 32 dependent steps instead of H)
__global__ void __launch_bounds__(1024)
tm_col_prefix_kernel(int W, int H, long long* __restrict__ S1, unsigned long long* __restrict__ S2) {
  __shared__ long long p1[32][33];
  __shared__ unsigned long long p2[32][33];
  const int cx = threadIdx.x, seg = threadIdx.y;
  const int x = blockIdx.x * 32 + cx;
  const int rows = (H + 31) / 32;
  const int y_begin = 1 + seg * rows, y_end = min(H + 1, y_begin + rows);
  long long c1 = 0;
  unsigned long long c2 = 0;
  if (x <= W)
    for (int y = y_begin; y < y_end; ++y) {
      const long long i = (long long)y * (W + 1) + x;
      c1 += S1[i];
      c2 += S2[i];
    }
  p1[seg][cx] = c1;
  p2[seg][cx] = c2;
  __syncthreads();
  if (seg == 0) {
    long long a1 = 0;
    unsigned long long a2 = 0;
    for (int s = 0; s < 32; ++s) {
      const long long t1 = p1[s][cx];
      const unsigned long long t2 = p2[s][cx];
      p1[s][cx] = a1;
      p2[s][cx] = a2;
      a1 += t1;
      a2 += t2;
    }
    if (x <= W) { S1[x] = 0; S2[x] = 0; }
  }
  __syncthreads();
  if (x > W) return;
  c1 = p1[seg][cx];
  c2 = p2[seg][cx];
  for (int y = y_begin; y < y_end; ++y) {
    const long long i = (long long)y * (W + 1) + x;
    c1 += S1[i];
    c2 += S2[i];
    S1[i] = c1;
    S2[i] = c2;
  }
}

__global__ void tm_template_stats_kernel(const TmArgs a, long long* __restrict__ tsum, long long* __restrict__ tsum2) {
  const TmJob& j = a.jobs[blockIdx.x];
  const uint8_t* t = a.tmpl + j.t_off;
  long long s1 = 0, s2 = 0;
  for (int i = threadIdx.x; i < j.w * j.h; i += blockDim.x) {
    const long long v = t[i];
    s1 += v;
    s2 += v * v;
  }
  __shared__ long long r1[32], r2[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t1 = 0, t2 = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { t1 += r1[w]; t2 += r2[w]; }
    tsum[blockIdx.x] = t1;
    tsum2[blockIdx.x] = t2;
  }
}

__device__ __forceinline__ bool tm_better(float v, int p, float bv, int bp) { return v > bv || (v == bv && p < bp); }

__global__ void __launch_bounds__(kThreads)
tm_corr_kernel(const TmArgs a) {
  extern __shared__ __align__(16) uint32_t tm_smem[];
  // which job / tile
  int ji = 0;
  while (ji + 1 < a.n_jobs && (int)blockIdx.x >= a.jobs[ji + 1].tile0) ++ji;
  const TmJob& j = a.jobs[ji];
  const int tile = blockIdx.x - j.tile0;
  const int x0 = (tile % j.tiles_x) * TX, y0 = (tile / j.tiles_x) * TY;
  // The four x positions of a thread read the SAME aligned patch words; the shift lives in the template instead:
  // copy i of a template row is the row moved right by i bytes (zero filled), so that
  // out[x + i] = sum_u' T_i[u'] I[x + u'] -- no byte permutes in the inner loop.
  const int tw_words = (j.w + 3 + 3) >> 2;                 // words of a shifted template row (w + 3 bytes)
  // patch pitch: TX + 4 * tw_words bytes are read; = 8 (mod 16) words keeps the two rows of a warp on different banks
  const int pitch = ((TX / 4 + tw_words + 1 + 15) & ~15) + 8;
  uint32_t* patch = tm_smem;                               // [TY + VC - 1][pitch]
  uint4* tmw = reinterpret_cast<uint4*>(patch + (((TY + VC - 1) * pitch + 3) & ~3));   // [VC][tw_words] x 4 shifts
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const uint8_t* T = a.tmpl + j.t_off;

  long long tot[2][4];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[r][i] = 0;

  for (int v0 = 0; v0 < j.h; v0 += VC) {
    const int vc = min(VC, j.h - v0);
    __syncthreads();
    // frame patch rows [y0 + v0, y0 + v0 + TY + vc - 1), columns [x0, x0 + 4 pitch): zero outside the frame
    const int prow = TY + vc - 1;
    for (int e = tid; e < prow * pitch; e += kThreads) {
      const int r = e / pitch, c = e - r * pitch;
      const int y = y0 + v0 + r, x = x0 + 4 * c;
      uint32_t word = 0;
      if (y < a.H) {
        const uint8_t* src = a.frame + (long long)y * a.ldf + x;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (x + b < a.W) word |= (uint32_t)src[b] << (8 * b);
      }
      patch[e] = word;
    }
    for (int e = tid; e < vc * tw_words; e += kThreads) {
      const int r = e / tw_words, c = e - r * tw_words;
      const uint8_t* src = T + (long long)(v0 + r) * j.w;
      uint32_t word[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int u = 4 * c + b - i;                     // template column under byte b of shifted copy i
          if (u >= 0 && u < j.w) word[i] |= (uint32_t)src[u] << (8 * b);
        }
      tmw[e] = make_uint4(word[0], word[1], word[2], word[3]);
    }
    __syncthreads();
    unsigned acc[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = 0u;
    for (int v = 0; v < vc; ++v) {
      const uint4* trow = tmw + v * tw_words;
      const uint32_t* r0 = patch + (2 * ty + v) * pitch + tx;
      const uint32_t* r1 = r0 + pitch;
#pragma unroll 4
      for (int u = 0; u < tw_words; ++u) {
        const uint4 t = trow[u];
        const uint32_t a0 = r0[u], b0 = r1[u];
        acc[0][0] = __dp4a(a0, t.x, acc[0][0]);
        acc[0][1] = __dp4a(a0, t.y, acc[0][1]);
        acc[0][2] = __dp4a(a0, t.z, acc[0][2]);
        acc[0][3] = __dp4a(a0, t.w, acc[0][3]);
        acc[1][0] = __dp4a(b0, t.x, acc[1][0]);
        acc[1][1] = __dp4a(b0, t.y, acc[1][1]);
        acc[1][2] = __dp4a(b0, t.z, acc[1][2]);
        acc[1][3] = __dp4a(b0, t.w, acc[1][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) tot[r][i] += acc[r][i];
  }

  // ---- scores (common_matchTemplate, TM_CCOEFF_NORMED, one channel) and the CTA's first maximum
  const double n = (double)j.w * j.h;
  const long long ts = a.tsum[ji], ts2 = a.tsum2[ji];
  // products and differences rounded separately (no fma contraction): the oracle's numpy arithmetic, bit for bit
  const double tnum = __dsub_rn(__dmul_rn(n, (double)ts2), __dmul_rn((double)ts, (double)ts));      // n^2 * variance
  const bool flat = tnum / (n * n) < DBL_EPSILON;
  const double tnorm = sqrt(tnum / n);                                // sqrt(sum (T - mean)^2)
  const double tmean = (double)ts / n;
  float best = -FLT_MAX;
  int best_p = INT_MAX;
  const int W1 = a.W + 1;
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = x0 + 4 * tx + i, y = y0 + 2 * ty + r;
      if (x >= j.rw || y >= j.rh) continue;
      float out = 1.0f;
      if (!flat) {
        const long long i00 = (long long)y * W1 + x, i10 = (long long)(y + j.h) * W1 + x;
        const long long sI = a.S1[i10 + j.w] - a.S1[i00 + j.w] - a.S1[i10] + a.S1[i00];
        const unsigned long long sI2 = a.S2[i10 + j.w] - a.S2[i00 + j.w] - a.S2[i10] + a.S2[i00];
        double num = __dsub_rn((double)tot[r][i], __dmul_rn((double)sI, tmean));
        const double wnd = __dsub_rn((double)sI2, __dmul_rn((double)sI, (double)sI) / n);
        const double t = sqrt(wnd > 0.0 ? wnd : 0.0) * tnorm;
        if (fabs(num) < t) num /= t;
        else if (fabs(num) < t * 1.125) num = num > 0.0 ? 1.0 : -1.0;
        else num = 0.0;
        out = (float)num;
      }
      const int p = y * j.rw + x;
      if (a.result) a.result[j.r_off + p] = out;
      if (tm_better(out, p, best, best_p)) { best = out; best_p = p; }
    }
  __shared__ float red_v[kThreads / 32];
  __shared__ int red_p[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int p2 = __shfl_xor_sync(0xffffffffu, best_p, o);
    if (tm_better(v2, p2, best, best_p)) { best = v2; best_p = p2; }
  }
  if ((tid & 31) == 0) { red_v[tid >> 5] = best; red_p[tid >> 5] = best_p; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kThreads / 32; ++w)
      if (tm_better(red_v[w], red_p[w], best, best_p)) { best = red_v[w]; best_p = red_p[w]; }
    a.cta_val[blockIdx.x] = best;
    a.cta_pos[blockIdx.x] = best_p;
  }
}

__global__ void tm_pick_kernel(const TmArgs a, int total_tiles, double* __restrict__ best_val,
                               int32_t* __restrict__ best_xy) {
  const int ji = blockIdx.x, lane = threadIdx.x;             // one warp per job
  const TmJob& j = a.jobs[ji];
  const int end = ji + 1 < a.n_jobs ? a.jobs[ji + 1].tile0 : total_tiles;
  float best = -FLT_MAX;
  int best_p = INT_MAX;
  for (int t = j.tile0 + lane; t < end; t += 32)
    if (tm_better(a.cta_val[t], a.cta_pos[t], best, best_p)) { best = a.cta_val[t]; best_p = a.cta_pos[t]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int p2 = __shfl_xor_sync(0xffffffffu, best_p, o);
    if (tm_better(v2, p2, best, best_p)) { best = v2; best_p = p2; }
  }
  if (lane == 0) {
    best_val[ji] = (double)best;
    best_xy[2 * ji] = best_p % j.rw;
    best_xy[2 * ji + 1] = best_p / j.rw;
  }
}

struct Layout {
  size_t s1, s2, tsum, tsum2, cta_val, cta_pos, total;
};

Layout layout(int W, int H, int n_jobs, long long total_tiles) {
  Layout l{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  l.s1 = take(sizeof(long long) * (size_t)(W + 1) * (H + 1));
  l.s2 = take(sizeof(long long) * (size_t)(W + 1) * (H + 1));
  l.tsum = take(sizeof(long long) * n_jobs);
  l.tsum2 = take(sizeof(long long) * n_jobs);
  l.cta_val = take(sizeof(float) * (size_t)total_tiles);
  l.cta_pos = take(sizeof(int) * (size_t)total_tiles);
  l.total = off;
  return l;
}

long long count_tiles(int W, int H, int n_jobs, const int32_t* tw, const int32_t* th) {
  long long tiles = 0;
  for (int i = 0; i < n_jobs; ++i) {
    if (tw[i] <= 0 || th[i] <= 0 || tw[i] > W || th[i] > H || tw[i] > kMaxTemplateW) return -1;
    tiles += ef::ceil_div(W - tw[i] + 1, TX) * ef::ceil_div(H - th[i] + 1, TY);
  }
  return tiles;
}

}  // namespace

extern "C" {

size_t ef_template_match_work_bytes(int32_t W, int32_t H, int32_t n_jobs, const int32_t* tw, const int32_t* th) {
  if (W <= 0 || H <= 0 || n_jobs <= 0 || n_jobs > kMaxJobs || !tw || !th) return 0;
  const long long tiles = count_tiles(W, H, n_jobs, tw, th);
  if (tiles <= 0) return 0;
  return layout(W, H, n_jobs, tiles).total;
}

int ef_template_match_device(const uint8_t* frame, int64_t ldf, int32_t W, int32_t H, const uint8_t* templates,
                             const int64_t* t_off, const int32_t* tw, const int32_t* th, int32_t n_jobs, float* result,
                             const int64_t* r_off, double* best_val, int32_t* best_xy, void* work, size_t work_bytes,
                             ef_stream_t stream) {
  if (!frame || !templates || !t_off || !tw || !th || !best_val || !best_xy || !work) return EF_ERR_INVALID;
  if (W <= 0 || H <= 0 || ldf < W || n_jobs <= 0 || n_jobs > kMaxJobs || (result && !r_off)) return EF_ERR_INVALID;
  const long long tiles = count_tiles(W, H, n_jobs, tw, th);
  if (tiles <= 0 || tiles > INT_MAX) return EF_ERR_INVALID;
  const Layout l = layout(W, H, n_jobs, tiles);
  if (work_bytes < l.total) return EF_ERR_INVALID;
  cudaStream_t st = ef::as_stream(stream);
  char* wk = reinterpret_cast<char*>(work);
  TmArgs a{};
  a.frame = frame;
  a.ldf = ldf;
  a.W = W;
  a.H = H;
  a.tmpl = templates;
  a.S1 = reinterpret_cast<long long*>(wk + l.s1);
  a.S2 = reinterpret_cast<unsigned long long*>(wk + l.s2);
  a.tsum = reinterpret_cast<long long*>(wk + l.tsum);
  a.tsum2 = reinterpret_cast<long long*>(wk + l.tsum2);
  a.result = result;
  a.cta_val = reinterpret_cast<float*>(wk + l.cta_val);
  a.cta_pos = reinterpret_cast<int*>(wk + l.cta_pos);
  a.n_jobs = n_jobs;
  int tile0 = 0, max_words = 0;
  for (int i = 0; i < n_jobs; ++i) {
    TmJob& j = a.jobs[i];
    j.t_off = t_off[i];
    j.r_off = result ? r_off[i] : 0;
    j.w = tw[i];
    j.h = th[i];
    j.rw = W - tw[i] + 1;
    j.rh = H - th[i] + 1;
    j.tiles_x = (int)ef::ceil_div(j.rw, TX);
    j.tile0 = tile0;
    tile0 += j.tiles_x * (int)ef::ceil_div(j.rh, TY);
    const int tw_words = (j.w + 3 + 3) >> 2;
    const int pitch = ((TX / 4 + tw_words + 1 + 15) & ~15) + 8;
    const int words = (((TY + VC - 1) * pitch + 3) & ~3) + 4 * VC * tw_words;
    if (words > max_words) max_words = words;
  }
  const size_t smem = sizeof(uint32_t) * (size_t)max_words;
  if (smem > 48 * 1024) EF_ENSURE_SMEM(tm_corr_kernel, smem);
  EF_LAUNCH(tm_row_prefix_kernel, (unsigned)ef::ceil_div(H, 8), 256, 0, st, frame, (long long)ldf, W, H,
            const_cast<long long*>(a.S1), const_cast<unsigned long long*>(a.S2));
  EF_LAUNCH(tm_col_prefix_kernel, (unsigned)ef::ceil_div(W + 1, 32), dim3(32, 32), 0, st, W, H,
            const_cast<long long*>(a.S1), const_cast<unsigned long long*>(a.S2));
  EF_LAUNCH(tm_template_stats_kernel, (unsigned)n_jobs, 256, 0, st, a, const_cast<long long*>(a.tsum),
            const_cast<long long*>(a.tsum2));
  EF_LAUNCH(tm_corr_kernel, (unsigned)tiles, kThreads, smem, st, a);
  EF_LAUNCH(tm_pick_kernel, (unsigned)n_jobs, 32, 0, st, a, (int)tiles, best_val, best_xy);
  return EF_OK;
}

}  // extern "C"
