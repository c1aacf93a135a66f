/*
 * eigenfaces_b200.h -- C ABI of the B200-native Eigenfaces engine (libeigenfaces_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of saladbkp/face-detection-recognization-PCA:
 * crop preprocess -> projection on k eigenfaces -> nearest gallery identity (+ threshold, + reconstruction
 * error), and the PCA fit (Gram + eigendecomposition + back-projection).  The reference has no FFI of its
 * own (it is pure Python over numpy / OpenCV / scikit-learn), so every entry point names the reference
 * Python interface (file:line under the reference tree) it replaces.  INTEGRATION.md shows the ctypes stub
 * a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns EF_OK (0) or a negative EF_ERR_* code and
 *     never throws; ef_error_string() describes the code, ef_last_error_detail() the last CUDA message.
 *   - "host" entry points take caller-owned host memory (numpy arrays), are synchronous, and include the
 *     host<->device copies; "device" entry points take device pointers valid on the CURRENT CUDA device,
 *     enqueue work on `stream` (a cudaStream_t passed as void*) and return without synchronising.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns EF_ERR_CUDA.
 *   - double = IEEE binary64 everywhere; crops are uint8, row-major, one crop per row.
 */
#ifndef EIGENFACES_B200_H
#define EIGENFACES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EF_ABI_VERSION 1

typedef void* ef_stream_t; /* cudaStream_t */

enum ef_status {
  EF_OK = 0,
  EF_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, misaligned pitch) */
  EF_ERR_UNSUPPORTED = -2, /* shape outside what the kernels cover (limits documented per function) */
  EF_ERR_CUDA = -3,        /* CUDA runtime error, see ef_last_error_detail() */
  EF_ERR_NOMEM = -4,       /* device or host allocation failed */
  EF_ERR_NOCONVERGE = -5   /* Jacobi sweeps exhausted before reaching the tolerance */
};

/* Matching rule.  The reference only has cosine; L2 is the north-star extra. */
enum ef_metric {
  EF_METRIC_COSINE_SK = 0, /* sklearn cosine_similarity: normalise rows (zero norm -> 1) then dot; argmax.
                              scan-template-v4.py:274-275 */
  EF_METRIC_COSINE_G1 = 1, /* dot / (|a| |b|), zero norm -> 0.0; max.  useless/scan.py:58-78,121-127 */
  EF_METRIC_L2 = 2         /* squared Euclidean distance; argmin.  (north star; reference has none) */
};

/* ------------------------------------------------------------------------------------------------ misc */
int ef_version(void);                     /* EF_ABI_VERSION of the loaded library */
const char* ef_error_string(int status);  /* static string */
const char* ef_last_error_detail(void);   /* thread-local, last CUDA error text */
int64_t ef_launch_count(void);            /* kernels launched by this library since load (bench "gpu_launches") */
int ef_device_sm_count(int* sm_count);    /* SMs of the current device */

/* -------------------------------------------------------------------------------------- K1: preprocess */
/* One detection box: crop frames[frame][y:y+h, x:x+w].  Replaces the per-detection slicing of
 * scan-template-v4.py:360,390 and useless/scan.py:245. */
typedef struct ef_box {
  int32_t frame;
  int32_t x, y, w, h;
} ef_box_t;

/* Gray (BGR2GRAY, 15-bit fixed point) + bilinear INTER_LINEAR resize (11-bit fixed point, 2x2 box when the
 * crop is exactly 2x the target, copy when equal) + flatten, bit-exact with OpenCV.
 * Replaces cv2.cvtColor + cv2.resize + .flatten() at scan-template-v4.py:257-263, train-v5.py:329-332,
 * useless/scan.py:248-252.
 *   frames   device uint8 [n_frames][height][pitch] ; channels = 1 (gray) or 3 (BGR interleaved)
 *   boxes    device ef_box_t [n_boxes] ; every box must lie inside its frame (checked on device: a bad box
 *            yields an all-zero crop and sets *bad_boxes (device int32, may be NULL) to a non-zero count)
 *   out      device uint8 [n_boxes][out_stride], first dw*dh bytes of each row written
 */
int ef_preprocess(const uint8_t* frames, int64_t frame_stride, int32_t pitch, int32_t width, int32_t height,
                  int32_t channels, int32_t n_frames, const ef_box_t* boxes, int32_t n_boxes, int32_t dw,
                  int32_t dh, uint8_t* out, int64_t out_stride, int32_t* bad_boxes, ef_stream_t stream);

/* ------------------------------------------------------------------------------ recognition model handle */
typedef struct ef_model ef_model_t;

/* Everything a pickled reference model holds that recognition needs (all HOST pointers, copied at create).
 * Gen-1 dict (useless/train.py:147-158): basis = eigenfaces [D,k] (Fortran order: stride_d = 1, stride_k = D),
 *   mean = mean_face, scale = pca_mean = NULL, gallery = projected_data, labels = NULL, metric = COSINE_G1.
 * Gen-2 dict (train-v5.py:449-461): basis = pca.components_ [k,D] (C order: stride_d = 1, stride_k = D),
 *   mean = scaler.mean_, scale = scaler.scale_, pca_mean = pca.mean_, gallery = face_features,
 *   labels = face_labels, metric = COSINE_SK. */
typedef struct ef_model_desc {
  int32_t D;                  /* pixels per crop (face_dimensions / 64*64) */
  int32_t k;                  /* components */
  const double* basis;        /* element (d, c) at basis[d * basis_stride_d + c * basis_stride_k] */
  int64_t basis_stride_d;
  int64_t basis_stride_k;
  const double* mean;         /* [D] */
  const double* scale;        /* [D] or NULL: x -> (x - mean) / scale      (StandardScaler.transform) */
  const double* pca_mean;     /* [D] or NULL: subtracted after scaling      (PCA.transform's mean_) */
  const double* gallery;      /* [n_gallery][gallery_ld] stored projections */
  int64_t gallery_ld;
  int32_t n_gallery;
  const int32_t* labels;      /* [n_gallery] or NULL (label = gallery row index) */
  int32_t metric;             /* enum ef_metric */
  int32_t n_slices;           /* 7-bit digit planes of the basis (1..8); 0 = default 8 (float64-equivalent) */
  int32_t with_residual;      /* also produce the squared reconstruction error */
} ef_model_desc_t;

/* Output arrays of one recognition call; all optional except score/index.  Host or device pointers
 * depending on the entry point. */
typedef struct ef_result {
  double* proj;    /* [B][k]  projection (features)                       scan.py:96 / scan-template-v4.py:266 */
  double* score;   /* [B]     best cosine similarity (or min squared L2)   scan.py:127 / scan-template-v4.py:276 */
  int32_t* index;  /* [B]     gallery row of the best match, ties -> lowest index (np.argmax rule) */
  int32_t* label;  /* [B]     labels[index] if the score passes the threshold else -1   scan-template-v4.py:278-287 */
  double* resid2;  /* [B]     squared distance from face space (needs with_residual)  (north star) */
} ef_result_t;

int ef_model_create(ef_model_t** out, const ef_model_desc_t* desc);
void ef_model_destroy(ef_model_t* model);
/* Pre-size the internal workspaces (and, for the host path, device staging) for batches up to max_batch so
 * that later calls do not allocate. */
int ef_model_reserve(ef_model_t* model, int32_t max_batch);
int ef_model_dims(const ef_model_t* model, int32_t* D, int32_t* k, int32_t* n_gallery, int32_t* n_slices);
/* Select the recognition kernels: 2 (default) = single cluster kernel (TMA + tcgen05 kind::i8 + DSMEM reduction +
 * fused match) when the shape is covered, 1 = tcgen05 stream-K projection kernel + separate epilogue kernel(s),
 * 0 = CUDA-core dp4a projection.  All three produce the same integers and the same labels. */
int ef_model_set_tensor_cores(ef_model_t* model, int32_t enable);
/* Measurement hook (bench.py roofline): when enabled, every recognise call brackets its projection kernel (the
 * dominant kernel) with CUDA events on the launching stream; _read returns the mean duration over the calls since
 * the hook was (re)enabled.  Off by default. */
/* Health flag of the tensor-core pipeline: non-zero when an mbarrier wait inside the tcgen05 kernel timed out (the
 * kernel drains instead of hanging).  Synchronous 4-byte read; the host entry points check it on every call. */
int ef_model_status(ef_model_t* model, int32_t* tc_pipeline_timeouts);
int ef_model_kernel_timing(ef_model_t* model, int32_t enable);
int ef_model_kernel_timing_read(ef_model_t* model, int32_t* n_calls, double* project_ms_mean,
                                int32_t* used_tensor_cores);

/* Replaces, for a whole batch at once, project_face_to_eigenspace + recognize_face (useless/scan.py:80-132)
 * or extract_face_features (after the resize) + recognize_face_with_model (scan-template-v4.py:263-287).
 *   x    uint8 [B][ldx] preprocessed crops (ldx >= D; device variant: ldx % 16 == 0 and x 16-byte aligned)
 *   threshold  similarity_threshold (cosine: recognised when score >= threshold; L2: when score <= threshold)
 */
int ef_model_recognize_device(ef_model_t* model, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                              const ef_result_t* out, ef_stream_t stream);
int ef_model_recognize_host(ef_model_t* model, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                            const ef_result_t* out);
/* Queued submission for a STREAM of batches on one CUDA stream (serving loop; replaces the same reference lines as
 * ef_model_recognize_device).  A submitted batch joins a queue; ONE persistent kernel recognises all queued batches back
 * to back: its loads, tensor-core projections, cluster exchange, float64 features and nearest-gallery search run on
 * different warps and overlap ACROSS consecutive batches, so HBM streams without a pause between them.  The queue is
 * launched when queue_depth batches (default and maximum 16) are waiting, at ef_model_flush_device, and -- adaptive
 * depth -- as soon as the previous launch of the queue has finished (an idle GPU never waits for the queue to fill).
 *   - x and every array of out must stay valid and unmodified until ef_model_flush_device (or the next
 *     ef_model_recognize_* call on the model) has been enqueued; all results are complete, in stream order, after it;
 *   - every value is bit identical to ef_model_recognize_device.
 * Shapes outside the serving kernels (L2 metric, k > 21, more than 128 digit-plane columns, unaligned crops) are
 * recognised immediately, exactly like ef_model_recognize_device.  Any other recognise call on the model flushes the
 * queue first.  All calls of one queue must use the same stream (a submit on another stream flushes first).
 * ef_model_set_serving: kernel 0 (default) = the persistent queue kernel, 1 = the pipelined kernel of round 1 (one launch
 * per submit: streams batch i and matches batch i-1; out->proj / out->resid2 written by the batch's own launch, the rest by
 * the next submit or the flush); queue_depth 1..32 (0 keeps the current value; a negative value -d sets depth d and
 * switches the adaptive early launch off: the queue then goes out only when full or flushed).  Only with nothing
 * queued. */
int ef_model_submit_device(ef_model_t* model, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                           const ef_result_t* out, ef_stream_t stream);
int ef_model_flush_device(ef_model_t* model, ef_stream_t stream);
int ef_model_set_serving(ef_model_t* model, int32_t kernel, int32_t queue_depth);
/* Asynchronous form of ef_model_recognize_host for a serving loop: at most two batches in flight.  submit enqueues the
 * chunked host->device copy, the kernels and the device->host copy of the results and returns a ticket (0 or 1); wait
 * blocks until that batch is done and fills the caller's arrays.  The copy of batch i+1 overlaps the kernels and the
 * result copy of batch i, so a steady stream of batches runs at the PCIe rate.  x must stay valid (and should be
 * page-locked) until the matching wait.  want: bit 0 = features, bit 1 = resid2, bit 2 = labels (score and index are
 * always produced); wait may only ask for what submit requested.  Same results as ef_model_recognize_host. */
int ef_model_submit_host(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold, int32_t want,
                         int32_t* ticket);
int ef_model_wait_host(ef_model_t* m, int32_t ticket, const ef_result_t* out);

/* Same, starting from frames + boxes (K1 then K2).  Host variant copies the frames and boxes in.
 * Every box must lie inside its frame (the reference slices numpy arrays, scan-template-v4.py:360,390: an out-of-frame
 * box there is clipped or makes cv2.resize raise).  Here a bad box is never recognised silently: the host variant
 * returns EF_ERR_INVALID (ef_last_error_detail() says how many), the asynchronous device variant writes an all-zero
 * crop's results and counts the box; ef_model_bad_boxes synchronises `stream`, returns the count accumulated by the
 * device-variant calls enqueued on it so far and clears it. */
int ef_model_recognize_boxes_device(ef_model_t* model, const uint8_t* frames, int64_t frame_stride, int32_t pitch,
                                    int32_t width, int32_t height, int32_t channels, int32_t n_frames,
                                    const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh,
                                    double threshold, const ef_result_t* out, ef_stream_t stream);
int ef_model_recognize_boxes_host(ef_model_t* model, const uint8_t* frames, int64_t frame_stride, int32_t pitch,
                                  int32_t width, int32_t height, int32_t channels, int32_t n_frames,
                                  const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh,
                                  double threshold, const ef_result_t* out);
/* recognize_face_all_models (scan-template-v4.py:289-319) in ONE call: the same boxes against every model of `models`
 * (all with D = dw * dh).  Frames and boxes are uploaded once, K1 runs once, one device->host copy and one
 * synchronisation return score / index / label as [n_models][n_boxes] row-major host arrays (label = -1 below the
 * threshold).  The keep-the-best / name fallback rules of the reference stay with the caller (they need the pickles'
 * person_id_map).  Uses models[0]'s stream and staging; no model may have batches queued on its serving queue. */
int ef_models_recognize_boxes_host(ef_model_t* const* models, int32_t n_models, const uint8_t* frames,
                                   int64_t frame_stride, int32_t pitch, int32_t width, int32_t height, int32_t channels,
                                   int32_t n_frames, const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh,
                                   double threshold, double* score, int32_t* index, int32_t* label);

int ef_model_bad_boxes(ef_model_t* model, ef_stream_t stream, int32_t* count);

/* Nearest-gallery search on already projected features (device pointers).  Used for the sharded gallery:
 * every rank prepares and matches against its shard, then the (score, index) pairs are all-gathered (NCCL) and
 * reduced with ef_match_reduce_device.
 *   ef_gallery_prepare_device: gallery [n][ldg] -> prepared [n][ldp] (rows L2-normalised for COSINE_SK, copied
 *     otherwise) and norms [n] (required for COSINE_G1, may be NULL otherwise).
 *   ef_match_device: p [B][ldp] features; index_base is added to the returned row index (global row id);
 *     out_score [B], out_index [B] (int64 because of index_base); work = device scratch of
 *     ef_match_work_bytes(B, n) bytes (may be NULL: the gallery is then not split across CTAs). */
size_t ef_match_work_bytes(int32_t B, int64_t n);
int ef_gallery_prepare_device(const double* gallery, int64_t ldg, int64_t n, int32_t k, int32_t metric,
                              double* prepared, int64_t ldp, double* norms, ef_stream_t stream);
int ef_match_device(const double* p, int64_t ldp, int32_t B, int32_t k, const double* prepared, int64_t ldg,
                    const double* norms, int64_t n, int64_t index_base, int32_t metric, double* out_score,
                    int64_t* out_index, void* work, ef_stream_t stream);
/* The same search on TENSOR CORES for large galleries (k <= 128, all three metrics): a float16 hi/lo filter GEMM
 * (tcgen05.mma kind::f16, float32 accumulation in TMEM) finds, per query, every gallery row within a proven error band
 * of the approximate maximum; only those rows are scored in float64, with exactly the arithmetic of ef_match_device, so
 * out_score / out_index are bit identical to it.
 *   image: device buffer of ef_match_tc_image_bytes_metric(n, k, metric) bytes (ef_match_tc_image_bytes(n, k) = the
 *          cosine size) filled once per gallery (shard) by ef_match_tc_prepare_device from the prepared rows (+ norms
 *          for COSINE_G1 and L2).  EF_METRIC_L2 carries one extra component (|g|^2 against the largest gallery norm)
 *          so that the same GEMM orders rows by distance; scores are sum (p - g)^2 as in ef_match_device;
 *   work:  device scratch of ef_match_tc_work_bytes(B, n, k) bytes, 256-byte aligned.  After the call has completed,
 *          ef_match_tc_flags(work, flags) reads {pipeline timeout, candidates re-scored, candidate-list overflow};
 *          on overflow (thousands of rows of every query inside the band: degenerate gallery) the outputs are
 *          incomplete and the caller must use ef_match_device.
 * EF_ERR_UNSUPPORTED for k > 128. */
size_t ef_match_tc_image_bytes(int64_t n, int32_t k);
size_t ef_match_tc_image_bytes_metric(int64_t n, int32_t k, int32_t metric);
int ef_match_tc_prepare_device(const double* prepared, int64_t ldg, const double* norms, int64_t n, int32_t k,
                               int32_t metric, void* image, ef_stream_t stream);
size_t ef_match_tc_work_bytes(int32_t B, int64_t n, int32_t k);
int ef_match_tc_device(const double* p, int64_t ldp, int32_t B, int32_t k, const double* prepared, int64_t ldg,
                       const double* norms, const void* image, int64_t n, int64_t index_base, int32_t metric,
                       double* out_score, int64_t* out_index, void* work, size_t work_bytes, ef_stream_t stream);
int ef_match_tc_flags(const void* work, int32_t* flags3);
/* Reduce R candidate lists [R][B] (as produced by an all-gather of ef_match_device outputs) to the best per
 * query: higher cosine / lower L2 wins, ties -> smallest global index. */
int ef_match_reduce_device(const double* scores, const int64_t* indices, int32_t R, int32_t B, int32_t metric,
                           double* out_score, int64_t* out_index, ef_stream_t stream);

/* ------------------------------------------------------------------------------- template-matching detector */
/* cv2.matchTemplate(frame, template, cv2.TM_CCOEFF_NORMED) + cv2.minMaxLoc for n_jobs <= 64 templates against one gray
 * frame -- the inner loop of MultiModelFaceScanner.template_match_all_models, scan-template-v4.py:147-174 (every
 * template image of every person at the scales 0.8 / 1.0 / 1.2 is one job).
 *   frame      device uint8 [H][ldf]
 *   templates  device uint8, job i = th[i] rows of tw[i] bytes, contiguous, starting at templates + t_off[i]
 *   t_off, tw, th, r_off   HOST arrays [n_jobs]; every template must fit the frame and be at most 1024 wide
 *   result     device float32 or NULL; job i's map [H - th + 1][W - tw + 1] is written at result + r_off[i]
 *   best_val   device double [n_jobs]   the map's maximum (float32 value, like minMaxLoc on the cv2 result)
 *   best_xy    device int32 [n_jobs][2] its first position in row-major order (x, y) = minMaxLoc's max_loc
 *   work       device scratch of ef_template_match_work_bytes(...) bytes
 * All sums are exact integers (dp4a cross term, integral images); the score is common_matchTemplate's formula in
 * float64, rounded to float32.  Agrees with cv2 to the accuracy of cv2's float32 DFT (about 1e-5). */
size_t ef_template_match_work_bytes(int32_t W, int32_t H, int32_t n_jobs, const int32_t* tw, const int32_t* th);
int ef_template_match_device(const uint8_t* frame, int64_t ldf, int32_t W, int32_t H, const uint8_t* templates,
                             const int64_t* t_off, const int32_t* tw, const int32_t* th, int32_t n_jobs, float* result,
                             const int64_t* r_off, double* best_val, int32_t* best_xy, void* work, size_t work_bytes,
                             ef_stream_t stream);

/* --------------------------------------------------------------------------------------------- PCA fit */
/* manual_pca(data_matrix, n_components) -- useless/train.py:56-128 -- for uint8 crops.
 *   X host uint8 [N][ldx]; k = n_components (clamped to the number of eigenvalues like :114)
 *   eigenfaces  host double [D][k] in FORTRAN order (column c contiguous), like the reference's slice
 *   mean [D], projected [N][k] row-major, eigenvalues [k] (descending, of the covariance /(N-1))
 * Snapshot branch (N < D): Gram N x N on the device, Jacobi eigensolver, back-projection of the top k,
 * column normalisation, projection.  Covariance branch (N >= D) forms the D x D covariance instead.
 * Limits: min(N, D) <= 4096 (Jacobi in L2-resident global memory). */
typedef struct ef_fit_info {
  int32_t sweeps;       /* Jacobi sweeps used */
  int32_t branch;       /* 0 = snapshot N x N, 1 = covariance D x D */
  double off_norm;      /* final max |cos| between rotated column pairs */
  double gpu_ms;        /* device time of the whole fit (CUDA events), excluding H2D/D2H */
} ef_fit_info_t;

int ef_fit_gen1_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, double* eigenfaces,
                     double* mean, double* projected, double* eigenvalues, ef_fit_info_t* info);

/* MultiFaceTrainer.train_pca_model -- train-v5.py:349-385 -- StandardScaler.fit_transform + PCA(k, solver full)
 * .fit_transform on uint8 crops.  All outputs host double:
 *   mean_face [D] (pixel mean, :366); scaler_mean/var/scale [D]; pca_mean [D]; components [k][D] row-major with
 *   sklearn's svd_flip sign rule; explained_variance [k]; explained_variance_ratio [k]; singular_values [k];
 *   *noise_variance; features [N][k] (= U S, the stored face_features). */
typedef struct ef_gen2_fit {
  double* mean_face;
  double* scaler_mean;
  double* scaler_var;
  double* scaler_scale;
  double* pca_mean;
  double* components;
  double* explained_variance;
  double* explained_variance_ratio;
  double* singular_values;
  double* noise_variance;
  double* features;
} ef_gen2_fit_t;

int ef_fit_gen2_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                     ef_fit_info_t* info);

/* The scripts/manual generation of the same trainer -- ManualStandardScaler + ManualPCA.fit_transform,
 * scripts/manual/train-v2.py:9-72, called at :189-193 -- on uint8 crops.  Differences to ef_fit_gen2_host: scale_ is
 * np.std (population) with EXACT zeros mapped to 1 (:61-62; sklearn also maps near-constant columns), and the components
 * are the leading eigenvectors of np.cov (:25-36), i.e. the same directions as the SVD's with the sign left to LAPACK by
 * the reference (here: the svd_flip rule, so results are deterministic).  explained_variance_ratio = lambda_i / sum of ALL
 * eigenvalues (:39-40).  k <= min(N, D): beyond the rank the reference returns arbitrary null-space vectors.  scaler_var
 * receives scale_^2 before the zero rule. */
int ef_fit_manual_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                       ef_fit_info_t* info);
/* PCA(k, solver full).fit_transform / ManualPCA.fit_transform of an arbitrary float64 host matrix Z [N][ldz] (what the
 * estimator objects receive when a caller scales the data itself: train-v5.py:373, scripts/manual/train-v2.py:193).
 * Only pca_mean, components, explained_variance(_ratio), singular_values, noise_variance, features of `out` are
 * written (the other pointers may be NULL). */
int ef_pca_fit_f64_host(const double* Z, int64_t ldz, int32_t N, int32_t D, int32_t k, const ef_gen2_fit_t* out,
                        ef_fit_info_t* info);
/* StandardScaler.fit (flavour 0) / ManualStandardScaler.fit (flavour 1) alone, on uint8 crops: mean, var, scale [D]. */
int ef_scaler_fit_u8_host(const uint8_t* X, int64_t ldx, int32_t N, int32_t D, int32_t flavour, double* mean,
                          double* var, double* scale);

/* ----------------------------------------------------------------- fit building blocks (device pointers) */
/* Used by the row-sharded multi-GPU fit: each rank calls these on its rows, the partial sums are
 * all-reduced (NCCL) between the calls. */

/* Column sums of uint8 rows: out[d] = sum_n X[n][d]  (exact, int64). */
int ef_colsum_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int64_t* out, ef_stream_t stream);
/* Exact integer Gram on the small side.  side = 0: G = X X^T  (N x N, over pixels [d0, d1));
 * side = 1: G = X^T X (D x D, over the rows given).  G int64 [n][n] row-major, ACCUMULATED into (+=). */
int ef_gram_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                      int64_t* G, ef_stream_t stream);
/* The same exact integer Gram on TENSOR CORES: tcgen05 kind::i8 (u8 x u8 -> s32 in TMEM), TMA-staged 128 x 256
 * tiles, s32 segments of <= 32768 bytes of K flushed into the int64 result, only upper-triangle tiles computed and the
 * lower triangle mirrored.  Same contract as ef_gram_u8_device (G += ...; G must be symmetric on entry, e.g. zero);
 * side 1 transposes X into `work` first so that both operands are K-major.
 *   work: device scratch of ef_gram_u8_tc_work_bytes(N, D, side) bytes, 256-byte aligned.  Its first int32 is a
 *   health flag: non-zero after the call completes means an mbarrier wait inside the kernel timed out.
 * EF_ERR_UNSUPPORTED (use ef_gram_u8_device) when X or ldx is not 16-byte aligned, d0 % 16 != 0, or side 1 with a
 * pixel sub-range.  Replaces np.dot(Xc, Xc.T) / np.cov(Xc.T) at useless/train.py:84,99 together with
 * ef_gram_center_device. */
size_t ef_gram_u8_tc_work_bytes(int64_t N, int32_t D, int32_t side);
/* ef_gram_u8_tc_store_device: the same call with G = (instead of +=): whatever G held is ignored, so a fresh Gram matrix
 * needs neither a zero fill nor the int64 read-add-write of the result (800 MB each for D = 10 000), and the tile
 * epilogues write the lower triangle themselves instead of a mirror pass reading the upper one back. */
int ef_gram_u8_tc_store_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                               int64_t* G, void* work, size_t work_bytes, ef_stream_t stream);
int ef_gram_u8_tc_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, int32_t d0, int32_t d1, int32_t side,
                         int64_t* G, void* work, size_t work_bytes, ef_stream_t stream);
/* Centre an integer Gram into the float64 covariance-like matrix with an exact integer numerator:
 * side 0: C[i][j] = alpha * (n^2 G[i][j] - n (r_i + r_j) + g) / n^2, r = row sums of G, g = grand sum (double centring)
 * side 1: C[a][b] = alpha * (N G[a][b] - s_a s_b) / N, s = column sums of X, N = total rows.
 * work: device scratch of ef_gram_center_work_bytes(n) (side 0 only). */
size_t ef_gram_center_work_bytes(int32_t n);
int ef_gram_center_device(const int64_t* G, int32_t n, int32_t side, const int64_t* colsum, int64_t N,
                          double alpha, double* C, void* work, ef_stream_t stream);
/* Symmetric eigendecomposition by one-sided Jacobi: A [n][n] float64 (destroyed), evals [n] descending,
 * evecs [n][n] row i = i-th eigenvector.  work: device scratch of ef_eigh_work_bytes(n). */
size_t ef_eigh_work_bytes(int32_t n);
int ef_eigh_jacobi_device(double* A, int32_t n, double* evals, double* evecs, void* work, int32_t max_sweeps,
                          double tol, int32_t* sweeps_used, double* off_norm, ef_stream_t stream);
/* Cholesky factor and its inverse of a symmetric positive definite matrix (the CholeskyQR orthonormalisation of the
 * subspace solver that stands in for np.linalg.eigh of a 10 000 x 10 000 covariance, useless/train.py:103):
 * G [m][m] float64 row-major, lower triangle read, overwritten by L (G = L L^T); Linv [m][m] = L^-1 (lower triangular).
 * info (device int): 0, or 1 + the column at which a pivot was not positive (G is then partly overwritten, Linv
 * undefined).  m <= 640; one CTA. */
int ef_chol_inverse_device(double* G, int32_t m, double* Linv, int32_t* info, ef_stream_t stream);
/* General strided float64 GEMM: C[m][n] = alpha * sum_k A(m,k) B(k,n) + beta * C[m][n],
 * A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn], C row-major with ldc. */
int ef_dgemm_device(int32_t M, int32_t N, int32_t K, double alpha, const double* A, int64_t sam, int64_t sak,
                    const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc,
                    ef_stream_t stream);
/* The same product on the FP64 tensor-core path (mma.sync m8n8k4 f64, 128 x 128 x 16 tiles): one of the two strides of
 * each operand must be 1.  splits > 1 cuts K into that many ranges whose partial products (work: device scratch of
 * ef_dgemm_tc_work_bytes(M, N, splits)) are added in ascending order -- for small outputs with a long K.  The
 * summation order of an output element does not depend on M, N or the tile it falls in. */
size_t ef_dgemm_tc_work_bytes(int32_t M, int32_t N, int32_t splits);
int ef_dgemm_tc_device(int32_t M, int32_t N, int32_t K, double alpha, const double* A, int64_t sam, int64_t sak,
                       const double* B, int64_t sbk, int64_t sbn, double beta, double* C, int64_t ldc, int32_t splits,
                       void* work, ef_stream_t stream);
/* Z[n][d] = (X[n][d] - mean[d]) / scale[d] - shift[d]   (scale, shift may be NULL); Z float64 [N][ldz]. */
int ef_standardize_u8_device(const uint8_t* X, int64_t ldx, int64_t N, int32_t D, const double* mean,
                             const double* scale, const double* shift, double* Z, int64_t ldz,
                             ef_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EIGENFACES_B200_H */
