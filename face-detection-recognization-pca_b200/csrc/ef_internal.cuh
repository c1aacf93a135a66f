// Internal (non-ABI) entry points shared between the translation units of libeigenfaces_b200.
#pragma once
#include <algorithm>

#include "ef_common.cuh"

namespace ef {

#ifdef __CUDACC__
// Digit planes -> float64, the ONE place every kernel path goes through: v = sum_s plane[s] 2^-(7s+6) with exact integer
// plane sums (|plane[s]| < 2^31).  Planes 0..3 and 4..7 are first combined as exact 64-bit integers,
//   hi = sum_{s<4} plane[s] 2^(7(3-s)),  lo = sum_{s>=4} plane[s] 2^(7(7-s))   (|hi|, |lo| < 2^53),
// so that partial (hi, lo) pairs of K ranges / CTAs / GPUs add exactly; each converts to float64 exactly and
// v = hi 2^-27 + lo 2^-55 is rounded ONCE: the correctly rounded value of the exact sum, whatever the split.
__device__ __forceinline__ void planes_to_hilo(const int32_t (&plane)[8], long long& hi, long long& lo) {
  hi = ((long long)plane[0] << 21) + ((long long)plane[1] << 14) + ((long long)plane[2] << 7) + (long long)plane[3];
  lo = ((long long)plane[4] << 21) + ((long long)plane[5] << 14) + ((long long)plane[6] << 7) + (long long)plane[7];
}
__device__ __forceinline__ double hilo_to_double(long long hi, long long lo) {
  // 2^-27 = 0x3E40..., 2^-55 = 0x3C80...: power-of-two scalings are exact, the fma rounds once
  return fma((double)lo, __longlong_as_double(0x3C80000000000000ll), (double)hi * __longlong_as_double(0x3E40000000000000ll));
}
__device__ __forceinline__ double planes_to_double(const int32_t (&plane)[8]) {
  long long hi, lo;
  planes_to_hilo(plane, hi, lo);
  return hilo_to_double(hi, lo);
}
#endif

// Accumulator convention shared by all projection / epilogue kernels: int32 acc_t[NC][ld_acc] (plane-major, crop b in
// column b), all zero on entry to a projection kernel, consumed AND cleared by the epilogue kernels.

// ef_project.cu -- CUDA-core exact-integer projection (reference implementation of the integer semantics)
int project_dp4a(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int32_t* acc_t,
                 int ld_acc, cudaStream_t stream);
// overlap: the kernel launched just before on this stream (project_tc: it lets its dependents be scheduled at once) only
// READS what this one reads and does not touch `out` -- the weighted kernel then runs concurrently with it and waits for
// it at its end (programmatic dependent launch); never set it behind a kernel that still reads `out`.
int row_sumsq(const uint8_t* X, int64_t ldx, int B, int D, const double* qq, double* out, cudaStream_t stream,
              bool overlap = false);

// ef_project_tc.cu -- tcgen05 kind::i8 projection (same integers as project_dp4a); sumsq (may be null) receives
// += sum_d x^2 per crop.  Returns EF_ERR_UNSUPPORTED when the buffers do not meet the TMA alignment rules.
// part == null: stream-K schedule, partial tiles merged into acc_t with int32 RED atomics.
// part != null: split-K schedule, partial tiles STORED row-major as part[split][crop][ld_part] (shape from
// project_tc_split_shape, bytes from project_tc_part_bytes); acc_t is not touched; project_finalize_slabs consumes it.
// combine (S = 8, Wq FEATURE-major: plane s of component c in row c * 8 + s): the slabs hold the exact (hi, lo) int64 pair
// of every component instead of its eight int32 planes (half the bytes): long long [split][crop][ld_part / 4].
int project_tc(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows,
               int32_t* acc_t, int ld_acc, double* sumsq, int* status, cudaStream_t stream, int32_t* part = nullptr,
               bool combine = false);
void project_tc_split_shape(int B, int D, int NC, int* splits, int* ld_part);
// Tail split of the combined slabs (more tiles than SMs): tiles [0, first) are unsplit in slab 0, tile t >= first has
// `splits` partial tiles at part + region + ((t - first) * splits + s) * 128 * block_n (int32 units, rows of block_n);
// tile id = column tile * m_tiles + crop tile.  first < 0: no tail.
struct TcTail {
  int first, splits, block_n, m_tiles;
  long long region;
};
bool project_tc_tail_shape(int B, int D, int NC, TcTail* t);
size_t project_tc_part_elems(int B, int D, int NC);   // int32 elements of a slab buffer for batches of up to B crops
size_t project_tc_part_bytes(int B, int D, int NC);

// ef_recognize_cluster.cu -- single-kernel form (cluster of 4, DSMEM reduction, fused match); EF_ERR_UNSUPPORTED when
// the shape is outside its coverage (k <= 32, S*(k+1) <= 256, aligned buffers)
int recognize_cluster(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows,
                      int k, int kq, int S, const int32_t* col_exp, const double* bias, const double* sumsq_ext,
                      bool want_resid, double c0, const double* gp_padded, int kpad, const double* gnorm,
                      const double* ginv, const void* gimg, int64_t n, const int32_t* labels, int metric,
                      double threshold, double* out_proj, double* out_score, int32_t* out_index, int32_t* out_label,
                      double* out_resid, int* status, cudaStream_t stream);
// ef_recognize_pipe.cu -- software-pipelined form: streams batch i while matching batch i-1 (rows carried in global
// memory between launches); B = 0 flushes, Bp = 0 is the first submit
bool pipe_supported(int k, int NC, int metric, int64_t n);
int filter_kf(int k);
int recognize_pipe(const uint8_t* X, int64_t ldx, int B, int D, const int8_t* Wq, int64_t ldw, int NC, int wq_rows, int k,
                   int kq, int S, const int32_t* col_exp, const double* bias, const double* sumsq_ext, bool want_resid,
                   double c0, double* out_proj, double* out_resid, double* carry_pe, double* carry_pn, void* carry_img,
                   int Bp, const double* prev_pe, const double* prev_pn, const void* prev_img, double* out_score,
                   int32_t* out_index, int32_t* out_label, double threshold, const double* gp_padded, int kpad,
                   const double* gnorm, const double* ginv, const void* gimg, int64_t n, const int32_t* labels,
                   int metric, int* status, cudaStream_t stream);
// ef_recognize_stream.cu -- persistent serving kernel over a queue of batches (ef_model_submit_device / _flush_device):
// loads, MMAs, cluster exchange, features and matching of consecutive batches overlap inside ONE launch.  Wfm is the
// FEATURE-MAJOR copy of the digit planes: row c * stream_plane_stride(S) + s holds plane s of column c.
constexpr int kStreamMaxBatches = 32;
struct StreamBatchDesc {
  alignas(64) unsigned char tmap[128];   // CUtensorMap of the crops, encoded at submit time (stream_encode_batch)
  const uint8_t* x;
  int64_t ldx;
  int B;
  const double* sumsq_ext;     // weighted sum of squares (standardised models with residual) or null
  double* out_proj;
  double* out_resid;
  double* out_score;
  int32_t* out_index;
  int32_t* out_label;
  double threshold;
};
int stream_plane_stride(int S);
// fills d->tmap from d->x / d->ldx / d->B (host work of ~1 us per batch, spread over the submits instead of the launch)
int stream_encode_batch(StreamBatchDesc* d, int D);

bool stream_supported(int D, int k, int kq, int S, int metric, int64_t n);
int recognize_stream(const StreamBatchDesc* batches, int nb, int D, const int8_t* Wfm, int64_t ldw, int wfm_rows, int k,
                     int kq, int S, const int32_t* col_exp, const double* bias, double c0, const double* gp_padded,
                     int kpad, const double* gnorm, const double* ginv, const void* gimg, int64_t n,
                     const int32_t* labels, int metric, int* status, cudaStream_t stream);
// float16 [g_hi | g_lo | g_hi] image of a prepared gallery for the tensor-core filter of the cluster kernel
size_t gallery_image_bytes(int k, int64_t n);
int gallery_image(const double* gp, int kr, const double* ginv, int64_t n, int k, int metric, void* img,
                  cudaStream_t stream);

// ef_epilogue.cu
bool fused_epilogue_supported(int k, int64_t n);
int fused_epilogue_kpad(int k);   // column count the prepared gallery must be zero-padded to for the fused kernel
int fused_epilogue(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp, const double* bias,
                   double* sumsq, double c0, const double* gp_padded, const double* gnorm, const double* ginv, int64_t n,
                   const int32_t* labels, int metric, double threshold, double* out_proj, double* out_score,
                   int32_t* out_index, int32_t* out_label, double* out_resid, cudaStream_t stream);
// resid_pass = false leaves x . u~ in resid2 and sumsq untouched: the caller's next kernel (match_small) finishes it
int project_finalize(int32_t* acc_t, int ld_acc, int B, int k, int kq, int S, const int32_t* col_exp,
                     const double* bias, double* proj, int64_t ldp, double* sumsq, double c0, double* resid2,
                     bool resid_pass, cudaStream_t stream);

// split-K slabs of project_tc -> float64 features (and x . u~ into resid2 when kq > k); nothing to clear
int project_finalize_slabs(const int32_t* part, int splits, int ld_part, int B, int k, int kq, int S,
                           const int32_t* col_exp, const double* bias, double* proj, int64_t ldp, double* resid2,
                           cudaStream_t stream, bool combined = false, const TcTail* tail = nullptr);

int project_resid(const double* proj, int64_t ldp, int B, int k, double* sumsq, double c0, double* resid2,
                  cudaStream_t stream);

// ef_match.cu
// Gallery preparation: gn[j][:] = g[j][:] / |g_j| (COSINE_SK), copy + norms (COSINE_G1), copy (L2).
// Columns [k, ldgp) of gp are left untouched (callers that need zero padding clear gp first); ginv (nullable)
// receives 1/|g| (0 for a zero row).
int gallery_prepare(const double* g, int64_t ldg, int64_t n, int k, int metric, double* gp, int64_t ldgp, double* gnorm,
                    double* ginv, cudaStream_t stream);
size_t match_work_bytes(int B, int64_t n);
int match(const double* p, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm, int64_t n,
          int64_t index_base, int metric, double* out_score, int64_t* out_index, void* work, cudaStream_t stream);
int label_lookup(const double* score, const int64_t* index, int B, const int32_t* labels, int metric, double threshold,
                 int32_t* out_index32, int32_t* out_label, cudaStream_t stream);

// ef_match_small.cu -- residual + match + threshold/label in one launch when every CTA can sweep the whole gallery
// (same float64 arithmetic as finalize_resid_kernel + match + label_lookup); resid2 (nullable) holds x . u~ on entry
bool match_small_supported(int B, int k, int64_t n);
int match_small(const double* proj, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm,
                int64_t n, const int32_t* labels, int metric, double threshold, double* sumsq, double c0, double* resid2,
                double* out_score, int32_t* out_index, int32_t* out_label, cudaStream_t stream);

// ef_match_small_tc.cu -- the same one-launch residual + match + threshold/label with the all-pairs scan on tensor cores
// (float16 hi/lo filter, exact float64 re-score of the rows inside the error band): 3 (k + 1) <= 576, n <= 4096.
// image = ef_match_tc_prepare_device's image of the prepared gallery; work = match_small_tc_work_bytes(cap_B) (256-byte
// aligned), zero before the first launch, the same cap_B >= B at every launch; status[0] is set when the tcgen05 pipeline timed out.
bool match_small_tc_supported(int k, int64_t n, int metric);
size_t match_small_tc_image_bytes(int k, int64_t n, int metric);
int match_small_tc_image(const double* gp, int64_t ldgp, const double* gnorm, int64_t n, int k, int metric, void* image,
                         cudaStream_t stream);
size_t match_small_tc_work_bytes(int cap_B, int64_t n, int k, int metric);
// slabs != null: proj is an OUTPUT -- the query kernel forms the features from the COMBINED split-K slabs of project_tc
// itself (what project_finalize_slabs does, one launch less) and leaves x . u~ in resid2 before the residual pass.
struct MatchSmallTcSlabs {
  const int32_t* part;
  int splits, ld_part, kq, S;
  bool combined;
  const int32_t* col_exp;
  const double* bias;
};
int match_small_tc(double* proj, int64_t ldp, int B, int k, const double* gp, int64_t ldgp, const double* gnorm,
                   const void* image, int64_t n, const int32_t* labels, int metric, double threshold, double* sumsq,
                   double c0, double* resid2, double* out_score, int32_t* out_index, int32_t* out_label, void* work,
                   int cap_B, int* status, cudaStream_t stream, const MatchSmallTcSlabs* slabs = nullptr);

// ef_gram_tc.cu -- exact integer Gram A A^T of uint8 rows on tensor cores (upper triangle + mirror), G int64 += .
// overwrite: G = A A^T (previous content ignored; no read of G, both triangles written by the tile epilogues)
// mn_major: A is X [K][lda] with the n outputs' axis contiguous (no transposed copy; n >= 256)
int gram_tc(const uint8_t* A, int64_t lda, int64_t n, int64_t K, int64_t* G, int64_t ldg, int* status,
            cudaStream_t stream, bool overwrite = false, bool mn_major = false);
int transpose_u8(const uint8_t* in, int64_t ldi, int64_t rows, int cols, uint8_t* out, int64_t ldo, cudaStream_t stream);

}  // namespace ef
