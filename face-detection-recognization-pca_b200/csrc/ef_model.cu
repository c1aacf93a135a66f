// Recognition model handle: device-resident sliced basis + prepared gallery, and the batch recognition entry
// points (host buffers and device buffers).
//
// ef_model_create does the one-time model preparation that the reference redoes per crop
// (scan-template-v4.py:265-266 re-validates and re-centres on every call; sklearn cosine_similarity re-normalises
// the whole gallery on every call, :274):
//   W_eff[d][c] = basis[d][c] / scale[d]            effective projection matrix
//   t[d]        = mean[d] + pca_mean[d] * scale[d]  so that features = (x - t) . W_eff
//   digit planes of W_eff (see ef_project.cu), bias[c] = t . W~[:,c] in float64
//   residual column u[d] = t[d] / scale[d]^2, c0 = sum t^2 / scale^2, qq[d] = 1 / scale[d]^2
#include <cmath>
#include <vector>

#include "ef_common.cuh"
#include "ef_internal.cuh"

struct ef_model {
  int D = 0, k = 0, kq = 0, S = 0, NC = 0, metric = 0;
  int64_t n_gallery = 0;
  int64_t ldw = 0;
  bool with_residual = false, has_scale = false;
  double c0 = 0.0;
  ef::DevBuf wq, col_exp, bias, qq, gp, gnorm, ginv, gimg, labels;
  int kpad = 0;                    // column pitch of the prepared gallery (zero padded for the fused epilogue)
  // workspaces (sized by reserve)
  int reserved = 0;
  ef::DevBuf acc, proj, sumsq, score, index64, match_work, status;
  ef::DevBuf mst_img, mst_work;     // tensor-core matcher of the shipped shapes (ef_match_small_tc.cu)
  int mst_cap = 0;                  // batch capacity mst_work is laid out for
  ef::DevBuf part;                 // split-K slabs of the tensor-core projection (k > 32)
  int nc_pad = 0;
  bool dirty = false;
  // host-path staging
  int host_reserved = 0;
  ef::DevBuf x_dev, resid_dev, index32_dev, label_dev, frames_dev, boxes_dev, bad_dev;
  int64_t x_ld = 0;
  cudaStream_t stream = nullptr;   // owned, used by the host entry points
  // pipelined submission (ef_model_submit_device): rows carried from one launch to the next, double buffered
  ef::DevBuf carry_pe[2], carry_pn[2], carry_img[2];
  int carry_reserved = 0;
  struct Pending {
    int B = 0;
    double* score = nullptr;
    int32_t* index = nullptr;
    int32_t* label = nullptr;
    double threshold = 0.0;
    int slot = 0;
    cudaStream_t stream = nullptr;                    // the stream the pending batch was submitted on
  } pending;
  cudaEvent_t flush_ev = nullptr;
  cudaEvent_t fan_ev = nullptr;     // multi-model fan-out: crops ready (first model) / this model's chain done
  void* pinned = nullptr;                             // owned page-locked staging for the results of the host path
  size_t pinned_bytes = 0;
  ef::DevBuf multi_dev;                               // ef_models_recognize_boxes_host: [n_models][n_boxes] results
  cudaStream_t copy_stream = nullptr;                 // owned: chunked H2D of the host path runs ahead of the kernels
  std::vector<cudaEvent_t> chunk_ev;                  // one per in-flight H2D chunk
  // asynchronous host path (ef_model_submit_host / ef_model_wait_host): two batches in flight, each with its own crop
  // buffer, page-locked result block and completion event
  struct HostSlot {
    ef::DevBuf x;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    cudaEvent_t done = nullptr;
    std::vector<cudaEvent_t> chunk_ev;
    int B = 0;
    bool busy = false, want_proj = false, want_resid = false, want_label = false;
  } hslot[2];
  int next_slot = 0;
  // queued submission (ef_model_submit_device with the persistent serving kernel): descriptors of the batches waiting
  // for their launch, all submitted on queue_stream; launched when queue_depth of them are waiting or at flush
  ef::DevBuf wq_fm;                // feature-major copy of the digit planes (row c * PS + s), read by the serving kernel
  int nc_fm = 0;
  std::vector<ef::StreamBatchDesc> queue;
  cudaStream_t queue_stream = nullptr;
  int queue_depth = ef::kStreamMaxBatches;
  cudaEvent_t queue_ev = nullptr;  // recorded behind every queue launch: "is the previous launch still running?"
  bool queue_ev_pending = false;
  bool queue_adaptive = true;
  int serving_kernel = 0;          // 0 persistent stream kernel, 1 pipelined kernel (one launch per batch)
  bool stream_ok = false;          // the persistent kernel covers this model's shape (decided once at create)
  ef::DevBuf sumsq_q[ef::kStreamMaxBatches];
  int ld_acc = 0;
  int tc_mode = 2;                 // 0 dp4a, 1 tcgen05 stream-K + epilogue kernels, 2 single cluster kernel
  int last_path = 0;
  ef::DevBuf sumsq_w;              // weighted sum of squares for the cluster kernel (standardised models)
  // optional per-kernel timing (bench roofline): event pairs around the projection kernel of every call
  bool timing = false;
  std::vector<cudaEvent_t> ev_a, ev_b;
  bool last_used_tc = false;
};

namespace {

int check_desc(const ef_model_desc_t* d) {
  if (!d || !d->basis || !d->mean || !d->gallery) return EF_ERR_INVALID;
  if (d->D <= 0 || d->k <= 0 || d->n_gallery <= 0 || d->gallery_ld < d->k) return EF_ERR_INVALID;
  if (d->metric < EF_METRIC_COSINE_SK || d->metric > EF_METRIC_L2) return EF_ERR_INVALID;
  if (d->n_slices < 0 || d->n_slices > 8) return EF_ERR_INVALID;
  // int32 accumulators: D * 255 * 64 must stay below 2^31
  if ((int64_t)d->D * 255 * 64 >= (1ll << 31)) return EF_ERR_UNSUPPORTED;
  return EF_OK;
}

// Signed 7-bit digit planes of one column, exact in float64: r_0 = w / 2^e in [-1, 1];
// q_s = rint(r_s * 2^(7s+6)), r_{s+1} = r_s - q_s * 2^-(7s+6), |r_{s+1}| <= 2^-(7s+7).
void slice_column(const std::vector<double>& w, int D, int S, int e, int8_t* planes, int64_t plane_stride,
                  std::vector<double>& wq_value) {
  for (int d = 0; d < D; ++d) {
    double r = std::ldexp(w[d], -e);
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
      const int sh = 7 * s + 6;
      double q = std::nearbyint(std::ldexp(r, sh));
      if (q > 64.0) q = 64.0;
      if (q < -64.0) q = -64.0;
      planes[(int64_t)s * plane_stride + d] = (int8_t)q;
      const double part = std::ldexp(q, -sh);
      r -= part;
      acc += part;
    }
    wq_value[d] = std::ldexp(acc, e);   // W~[d]: exact (the partial sums nest inside 53 bits for S <= 7, and
                                        // differ from the true sum by < 2^-53 relative for S = 8)
  }
}

// Uploads run on the stream that consumes the buffer (the model's own non-blocking stream does not order against the
// legacy default stream).  The host source is pageable: cudaMemcpyAsync stages it before returning, and create()
// synchronises the stream before any of the sources go out of scope.
int upload(ef::DevBuf& buf, const void* src, size_t bytes, cudaStream_t st) {
  EF_TRY(buf.ensure(bytes ? bytes : 16));
  if (bytes) EF_CUDA(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
  return EF_OK;
}

}  // namespace

extern "C" {

int ef_model_create(ef_model_t** out, const ef_model_desc_t* desc) {
  if (!out) return EF_ERR_INVALID;
  *out = nullptr;
  EF_TRY(check_desc(desc));
  int dev_count = 0;
  EF_CUDA(cudaGetDeviceCount(&dev_count));
  if (dev_count <= 0) return EF_ERR_CUDA;

  ef_model* m = new (std::nothrow) ef_model();
  if (!m) return EF_ERR_NOMEM;
  const int D = desc->D, k = desc->k;
  m->D = D;
  m->k = k;
  m->S = desc->n_slices == 0 ? 8 : desc->n_slices;
  m->with_residual = desc->with_residual != 0;
  m->has_scale = desc->scale != nullptr;
  m->kq = k + (m->with_residual ? 1 : 0);
  m->NC = m->S * m->kq;
  m->metric = desc->metric;
  m->n_gallery = desc->n_gallery;
  m->ldw = ef::round_up(D, 128);
  const int nc_pad = (int)ef::round_up(m->NC, 16);
  m->nc_pad = nc_pad;

  {
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { ef::set_error_detail("cudaStreamCreate", e); delete m; return EF_ERR_CUDA; }
  }
  // ---- host-side preparation (float64, exact digit extraction)
  std::vector<double> t(D), inv_s(D, 1.0);
  std::vector<double> qq;
  for (int d = 0; d < D; ++d) {
    const double s = desc->scale ? desc->scale[d] : 1.0;
    const double pm = desc->pca_mean ? desc->pca_mean[d] : 0.0;
    t[d] = desc->mean[d] + pm * s;
  }
  std::vector<int8_t> wq((size_t)nc_pad * m->ldw, 0);
  std::vector<int32_t> col_exp(m->kq, 0);
  std::vector<double> bias(m->kq, 0.0), col(D), colq(D);
  for (int c = 0; c < m->kq; ++c) {
    if (c < k) {
      for (int d = 0; d < D; ++d) {
        const double b = desc->basis[(int64_t)d * desc->basis_stride_d + (int64_t)c * desc->basis_stride_k];
        col[d] = desc->scale ? b / desc->scale[d] : b;
      }
    } else {
      for (int d = 0; d < D; ++d) {
        const double s = desc->scale ? desc->scale[d] : 1.0;
        col[d] = desc->scale ? t[d] / (s * s) : t[d];
      }
    }
    double mx = 0.0;
    for (int d = 0; d < D; ++d) {
      if (!std::isfinite(col[d])) { ef_model_destroy(m); return EF_ERR_INVALID; }
      mx = std::fmax(mx, std::fabs(col[d]));
    }
    int e = 0;
    if (mx > 0.0) {
      std::frexp(mx, &e);            // mx = f * 2^e, f in [0.5, 1)  ->  |col| / 2^e < 1
    }
    col_exp[c] = e;
    // plane s of column c is row (s * kq + c) of wq
    slice_column(col, D, m->S, e, wq.data() + (int64_t)c * m->ldw, (int64_t)m->kq * m->ldw, colq);
    if (c < k) {
      long double b = 0.0L;
      for (int d = 0; d < D; ++d) b += (long double)t[d] * (long double)colq[d];
      bias[c] = (double)b;
    }
  }
  if (m->with_residual) {
    long double c0 = 0.0L;
    for (int d = 0; d < D; ++d) {
      const double s = desc->scale ? desc->scale[d] : 1.0;
      c0 += (long double)t[d] * (long double)t[d] / ((long double)s * (long double)s);
    }
    m->c0 = (double)c0;
    if (desc->scale) {
      qq.resize(D);
      for (int d = 0; d < D; ++d) qq[d] = 1.0 / (desc->scale[d] * desc->scale[d]);
      int st = upload(m->qq, qq.data(), sizeof(double) * D, m->stream);
      if (st != EF_OK) { ef_model_destroy(m); return st; }
    }
  }

  // feature-major copy for the serving kernel: all planes of a column in adjacent rows
  const int PS = ef::stream_plane_stride(m->S);
  m->nc_fm = (int)ef::round_up((int64_t)m->kq * PS, 16);
  std::vector<int8_t> wfm((size_t)m->nc_fm * m->ldw, 0);
  for (int c = 0; c < m->kq; ++c)
    for (int sl = 0; sl < m->S; ++sl)
      memcpy(&wfm[(size_t)(c * PS + sl) * m->ldw], &wq[(size_t)(sl * m->kq + c) * m->ldw], (size_t)m->ldw);
  int st = upload(m->wq, wq.data(), wq.size(), m->stream);
  if (st == EF_OK) st = upload(m->wq_fm, wfm.data(), wfm.size(), m->stream);
  if (st == EF_OK) st = upload(m->col_exp, col_exp.data(), sizeof(int32_t) * col_exp.size(), m->stream);
  if (st == EF_OK) st = upload(m->bias, bias.data(), sizeof(double) * bias.size(), m->stream);
  if (st == EF_OK && desc->labels)
    st = upload(m->labels, desc->labels, sizeof(int32_t) * (size_t)desc->n_gallery, m->stream);
  // gallery: upload raw (compacted to ld = k), prepare on the device
  ef::DevBuf raw;
  std::vector<double> g;
  if (st == EF_OK) {
    g.resize((size_t)desc->n_gallery * k);
    for (int64_t j = 0; j < desc->n_gallery; ++j)
      memcpy(&g[(size_t)j * k], desc->gallery + j * desc->gallery_ld, sizeof(double) * k);
    st = upload(raw, g.data(), sizeof(double) * g.size(), m->stream);
  }
  m->kpad = ef::fused_epilogue_supported(k, desc->n_gallery) ? ef::fused_epilogue_kpad(k) : k;
  const size_t gp_bytes = sizeof(double) * (size_t)desc->n_gallery * m->kpad;
  const size_t gn_bytes = sizeof(double) * ((size_t)desc->n_gallery + 8);
  if (st == EF_OK) st = m->gp.ensure(gp_bytes);
  if (st == EF_OK) st = m->gnorm.ensure(gn_bytes);
  if (st == EF_OK) st = m->ginv.ensure(gn_bytes);
  if (st == EF_OK) {
    // zero padding of the prepared gallery: on the stream gallery_prepare runs on, so that it cannot race the kernel
    cudaError_t e = cudaMemsetAsync(m->gp.p, 0, gp_bytes, m->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->gnorm.p, 0, gn_bytes, m->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->ginv.p, 0, gn_bytes, m->stream);
    if (e != cudaSuccess) { ef::set_error_detail("cudaMemsetAsync", e); st = EF_ERR_CUDA; }
  }
  if (st == EF_OK)
    st = ef::gallery_prepare(raw.as<double>(), k, desc->n_gallery, k, m->metric, m->gp.as<double>(), m->kpad,
                             m->gnorm.as<double>(), m->ginv.as<double>(), m->stream);
  // float16 image of the gallery for the tensor-core filter of the cluster kernel (cosine metrics, k <= 32)
  if (st == EF_OK && m->metric != EF_METRIC_L2 && ef::fused_epilogue_supported(k, desc->n_gallery)) {
    st = m->gimg.ensure(ef::gallery_image_bytes(k, desc->n_gallery));
    if (st == EF_OK)
      st = ef::gallery_image(m->gp.as<double>(), m->kpad, m->ginv.as<double>(), desc->n_gallery, k, m->metric,
                             m->gimg.p, m->stream);
  }
  // ... and for the one-launch matcher of the shapes those kernels do not cover (k = 50 ... 178, <= 4096 rows)
  if (st == EF_OK && !ef::fused_epilogue_supported(k, desc->n_gallery) &&
      ef::match_small_tc_supported(k, desc->n_gallery, m->metric)) {
    st = m->mst_img.ensure(ef::match_small_tc_image_bytes(k, desc->n_gallery, m->metric));
    if (st == EF_OK)
      st = ef::match_small_tc_image(m->gp.as<double>(), m->kpad, m->gnorm.as<double>(), desc->n_gallery, k, m->metric,
                                    m->mst_img.p, m->stream);
  }
  m->stream_ok = st == EF_OK && m->gimg.p && ef::stream_supported(D, k, m->kq, m->S, m->metric, desc->n_gallery);
  if (st == EF_OK) {
    cudaError_t e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) { ef::set_error_detail("gallery_prepare", e); st = EF_ERR_CUDA; }
  }
  if (st != EF_OK) { ef_model_destroy(m); return st; }
  *out = m;
  return EF_OK;
}

void ef_model_destroy(ef_model_t* m) {
  if (!m) return;
  // batches submitted with ef_model_submit_host / _device may still be in flight: drain before the buffers go away
  if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->pending.B > 0) cudaStreamSynchronize(m->pending.stream);
  if (m->stream) cudaStreamDestroy(m->stream);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->pinned) cudaFreeHost(m->pinned);
  if (m->flush_ev) cudaEventDestroy(m->flush_ev);
  if (m->fan_ev) cudaEventDestroy(m->fan_ev);
  if (m->queue_ev) cudaEventDestroy(m->queue_ev);
  for (cudaEvent_t e : m->chunk_ev) cudaEventDestroy(e);
  for (auto& hs : m->hslot) {
    if (hs.pinned) cudaFreeHost(hs.pinned);
    if (hs.done) cudaEventDestroy(hs.done);
    for (cudaEvent_t e : hs.chunk_ev) cudaEventDestroy(e);
  }
  for (cudaEvent_t e : m->ev_a) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_b) cudaEventDestroy(e);
  delete m;
}

int ef_model_dims(const ef_model_t* m, int32_t* D, int32_t* k, int32_t* n_gallery, int32_t* n_slices) {
  if (!m) return EF_ERR_INVALID;
  if (D) *D = m->D;
  if (k) *k = m->k;
  if (n_gallery) *n_gallery = (int32_t)m->n_gallery;
  if (n_slices) *n_slices = m->S;
  return EF_OK;
}

int ef_model_set_tensor_cores(ef_model_t* m, int32_t enable) {
  if (!m) return EF_ERR_INVALID;
  m->tc_mode = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return EF_OK;
}

int ef_model_kernel_timing(ef_model_t* m, int32_t enable) {
  if (!m) return EF_ERR_INVALID;
  for (cudaEvent_t e : m->ev_a) cudaEventDestroy(e);
  for (cudaEvent_t e : m->ev_b) cudaEventDestroy(e);
  m->ev_a.clear();
  m->ev_b.clear();
  m->timing = enable != 0;
  return EF_OK;
}

int ef_model_kernel_timing_read(ef_model_t* m, int32_t* n_calls, double* project_ms_mean, int32_t* used_tensor_cores) {
  if (!m || !n_calls || !project_ms_mean) return EF_ERR_INVALID;
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i < m->ev_a.size(); ++i) {
    float ms = 0.f;
    EF_CUDA(cudaEventSynchronize(m->ev_b[i]));
    EF_CUDA(cudaEventElapsedTime(&ms, m->ev_a[i], m->ev_b[i]));
    tot += ms;
    ++n;
  }
  *n_calls = n;
  *project_ms_mean = n ? tot / n : 0.0;
  if (used_tensor_cores) *used_tensor_cores = m->last_path;
  return EF_OK;
}

int ef_model_reserve(ef_model_t* m, int32_t max_batch) {
  if (!m || max_batch <= 0) return EF_ERR_INVALID;
  if (max_batch <= m->reserved) return EF_OK;
  const size_t B = (size_t)ef::round_up(max_batch, 128);
  m->ld_acc = (int)B + 32;          // plane pitch off the power of two: consecutive planes land in different L2 slices
  // plane-major accumulators and the sum-of-squares buffer start out zero; the epilogue kernels keep them zero
  EF_TRY(m->acc.ensure(sizeof(int32_t) * (B + 32) * m->nc_pad));
  EF_CUDA(cudaMemset(m->acc.p, 0, sizeof(int32_t) * (B + 32) * m->nc_pad));
  EF_TRY(m->proj.ensure(sizeof(double) * B * m->k));
  if (!ef::fused_epilogue_supported(m->k, m->n_gallery)) {
    // split-K slabs of the projection (+ the tail region of its last partial wave)
    EF_TRY(m->part.ensure(sizeof(int32_t) * ef::project_tc_part_elems((int)B, m->D, m->NC)));
  }
  EF_TRY(m->sumsq.ensure(sizeof(double) * (B + 32)));
  EF_TRY(m->sumsq_w.ensure(sizeof(double) * (B + 32)));
  EF_CUDA(cudaMemset(m->sumsq.p, 0, sizeof(double) * (B + 32)));
  EF_TRY(m->status.ensure(16));
  EF_CUDA(cudaMemset(m->status.p, 0, 16));
  m->dirty = false;
  EF_TRY(m->score.ensure(sizeof(double) * B));
  EF_TRY(m->index64.ensure(sizeof(int64_t) * B));
  EF_TRY(m->match_work.ensure(ef::match_work_bytes(max_batch, m->n_gallery) + 16));
  if (m->mst_img.p) {
    const size_t wb = ef::match_small_tc_work_bytes((int)B, m->n_gallery, m->k, m->metric);
    EF_TRY(m->mst_work.ensure(wb));
    EF_CUDA(cudaMemset(m->mst_work.p, 0, wb));                // counters and image padding start (and stay) zero
    m->mst_cap = (int)B;
  }
  // the zero fills above ran on the legacy default stream, the kernels that rely on them run on non-blocking streams
  // (the model's own or the caller's), which do not order against it: finish them here (reserve is rare)
  EF_CUDA(cudaDeviceSynchronize());
  m->reserved = max_batch;
  return EF_OK;
}

int ef_model_recognize_device(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                              const ef_result_t* out, ef_stream_t stream) {
  if (m && B == 0) return EF_OK;
  if (!m || !x || !out || B < 0 || ldx < m->D) return EF_ERR_INVALID;
  if ((ldx & 15) || (reinterpret_cast<uintptr_t>(x) & 15)) return EF_ERR_INVALID;
  if (out->resid2 && !m->with_residual) return EF_ERR_INVALID;
  if (B == 0) return EF_OK;
  EF_TRY(ef_model_reserve(m, B));
  if (m->pending.B > 0 || !m->queue.empty())
    EF_TRY(ef_model_flush_device(m, stream));                       // results of queued / pipelined batches come out first
  cudaStream_t st = ef::as_stream(stream);
  int32_t* acc = m->acc.as<int32_t>();
  // 1. exact integer digit-plane dot products
  cudaEvent_t ea = nullptr, eb = nullptr;
  if (m->timing && m->ev_a.size() < 4096) {
    EF_CUDA(cudaEventCreate(&ea));
    EF_CUDA(cudaEventCreate(&eb));
    m->ev_a.push_back(ea);
    m->ev_b.push_back(eb);
    EF_CUDA(cudaEventRecord(ea, st));
  }
  if (!out->score || !out->index) return EF_ERR_INVALID;
  if (m->dirty) {   // a previous call failed between projection and epilogue: restore the all-zero invariant
    EF_CUDA(cudaMemsetAsync(acc, 0, sizeof(int32_t) * (size_t)m->ld_acc * m->nc_pad, st));
    EF_CUDA(cudaMemsetAsync(m->sumsq.p, 0, sizeof(double) * (size_t)m->ld_acc, st));
  }
  const bool want_resid = out->resid2 != nullptr;
  double* sumsq = m->sumsq.as<double>();
  const int32_t* labels = m->labels.p ? m->labels.as<int32_t>() : nullptr;
  if (m->tc_mode >= 2) {
    // single-kernel cluster form (TMA + tcgen05 + DSMEM reduction + fused match); falls through when unsupported
    const double* sumsq_ext = nullptr;
    const bool cluster_shape = m->k <= 32 && m->nc_pad <= 256;      // what recognize_cluster covers (else it declines)
    if (want_resid && m->has_scale && cluster_shape) {
      EF_TRY(ef::row_sumsq(x, ldx, B, m->D, m->qq.as<double>(), m->sumsq_w.as<double>(), st));
      sumsq_ext = m->sumsq_w.as<double>();
    }
    const int stc = !cluster_shape ? EF_ERR_UNSUPPORTED : ef::recognize_cluster(
        x, ldx, B, m->D, m->wq.as<int8_t>(), m->ldw, m->NC, m->nc_pad, m->k, m->kq, m->S, m->col_exp.as<int32_t>(),
        m->bias.as<double>(), sumsq_ext, want_resid, m->c0, m->gp.as<double>(), m->kpad, m->gnorm.as<double>(),
        m->ginv.as<double>(), m->gimg.p, m->n_gallery, labels, m->metric, threshold, out->proj, out->score, out->index,
        out->label, want_resid ? out->resid2 : nullptr, m->status.as<int>(), st);
    if (stc == EF_OK) {
      m->last_used_tc = true;
      m->last_path = 2;
      if (eb) EF_CUDA(cudaEventRecord(eb, st));
      return EF_OK;
    }
    if (stc != EF_ERR_UNSUPPORTED) return stc;
  }
  m->dirty = true;
  m->last_path = 0;
  // the tensor-core kernel also produces the integer sum of squares (Gen-1 residual) from the staged crop tiles
  const bool tc_sumsq = want_resid && !m->has_scale;
  int st_tc = EF_ERR_UNSUPPORTED;
  // k > 32: split-K slabs (plain stores, no accumulator invariants) instead of stream-K + int32 RED atomics
  int32_t* part = (m->part.p && !getenv("EF_NO_SLABS")) ? m->part.as<int32_t>() : nullptr;
  // ... holding, for S = 8, the (hi, lo) int64 pair of every component instead of its eight int32 planes: the
  // feature-major basis puts the planes of a component into adjacent accumulator columns (half the slab bytes)
  const bool combine = part && m->S == 8 && m->NC == 8 * m->kq && m->nc_fm >= m->NC && !getenv("EF_NO_SLAB_COMBINE");
  if (m->tc_mode >= 1)
    st_tc = combine ? ef::project_tc(x, ldx, B, m->D, m->wq_fm.as<int8_t>(), m->ldw, m->NC, m->nc_fm, acc, m->ld_acc,
                                     tc_sumsq ? sumsq : nullptr, m->status.as<int>(), st, part, true)
                    : ef::project_tc(x, ldx, B, m->D, m->wq.as<int8_t>(), m->ldw, m->NC, m->nc_pad, acc, m->ld_acc,
                                     tc_sumsq ? sumsq : nullptr, m->status.as<int>(), st, part);
  m->last_used_tc = st_tc == EF_OK;
  m->last_path = m->last_used_tc ? 1 : 0;
  if (st_tc != EF_OK) {
    if (st_tc != EF_ERR_UNSUPPORTED) return st_tc;
    EF_TRY(ef::project_dp4a(x, ldx, B, m->D, m->wq.as<int8_t>(), m->ldw, m->NC, acc, m->ld_acc, st));
  }
  if (eb) EF_CUDA(cudaEventRecord(eb, st));
  // 2. residual ingredient not produced by the projection kernel (weighted for the standardised models)
  if (want_resid && !(m->last_used_tc && tc_sumsq))
    EF_TRY(ef::row_sumsq(x, ldx, B, m->D, m->has_scale ? m->qq.as<double>() : nullptr, sumsq, st,
                         m->last_used_tc && !getenv("EF_NO_SUMSQ_OVERLAP")));
  if (ef::fused_epilogue_supported(m->k, m->n_gallery)) {
    // 3. one launch: planes -> features (+ residual) -> nearest gallery row -> threshold / label
    EF_TRY(ef::fused_epilogue(acc, m->ld_acc, B, m->k, m->kq, m->S, m->col_exp.as<int32_t>(), m->bias.as<double>(),
                              sumsq, m->c0, m->gp.as<double>(), m->gnorm.as<double>(), m->ginv.as<double>(),
                              m->n_gallery, labels, m->metric, threshold, out->proj, out->score, out->index, out->label,
                              want_resid ? out->resid2 : nullptr, st));
  } else {
    // 3. planes -> float64 features (+ residual); 4. nearest gallery row; 5. threshold + label
    double* proj = out->proj ? out->proj : m->proj.as<double>();
    bool small = ef::match_small_supported(B, m->k, m->n_gallery) && !getenv("EF_NO_MATCH_SMALL");
    // A handful of crops against a long gallery (the reference's own call pattern: ONE face per call, k = N = 590): the
    // one or two CTAs of match_small_kernel would sweep the whole gallery alone (218 us at B = 1, 590 x 590); the split
    // chain spreads the gallery rows over the SMs instead.
    if (small && ef::ceil_div(B, 32) * 8 < ef::sm_count() && (int64_t)m->n_gallery * m->k >= 16384) small = false;
    // tensor-core filter + exact re-score in one launch (same results, bit for bit): takes over from both
    const bool small_tc = m->mst_img.p && m->mst_work.p && !getenv("EF_NO_MATCH_SMALL_TC") && !getenv("EF_NO_MATCH_SMALL");
    if (small_tc) small = true;
    ef::MatchSmallTcSlabs slabs{};
    bool fused_finalize = false;
    if (m->last_used_tc && part) {
      int splits = 1, ld_part = 0;
      ef::project_tc_split_shape(B, m->D, m->NC, &splits, &ld_part);
      ef::TcTail tail{};
      const bool tailed = combine && ef::project_tc_tail_shape(B, m->D, m->NC, &tail);
      if (small_tc && combine && !tailed && m->k <= 191 && !getenv("EF_MST_NO_FUSED_FINALIZE")) {
        // the matcher's query kernel forms the features from the (hi, lo) slabs itself: one launch less, the features
        // stay in its registers.  (With int32 plane slabs -- eight loads per component and split -- a warp per crop was
        // slower than the thread-per-component finalize: 15.6 us against 6.2 + 3.8.)
        slabs = ef::MatchSmallTcSlabs{part, splits, ld_part, m->kq, m->S, true, m->col_exp.as<int32_t>(),
                                      m->bias.as<double>()};
        fused_finalize = true;
      } else {
        EF_TRY(ef::project_finalize_slabs(part, splits, ld_part, B, m->k, m->kq, m->S, m->col_exp.as<int32_t>(),
                                          m->bias.as<double>(), proj, m->k, want_resid ? out->resid2 : nullptr, st,
                                          combine, tailed ? &tail : nullptr));
        if (!small && want_resid)
          EF_TRY(ef::project_resid(proj, m->k, B, m->k, sumsq, m->c0, out->resid2, st));
      }
    } else {
      EF_TRY(ef::project_finalize(acc, m->ld_acc, B, m->k, m->kq, m->S, m->col_exp.as<int32_t>(),
                                  m->bias.as<double>(), proj, m->k, sumsq, m->c0, want_resid ? out->resid2 : nullptr,
                                  !small, st));
    }
    if (small_tc) {
      EF_TRY(ef::match_small_tc(proj, m->k, B, m->k, m->gp.as<double>(), m->kpad, m->gnorm.as<double>(), m->mst_img.p,
                                m->n_gallery, labels, m->metric, threshold, sumsq, m->c0,
                                want_resid ? out->resid2 : nullptr, out->score, out->index, out->label, m->mst_work.p,
                                m->mst_cap, m->status.as<int>(), st, fused_finalize ? &slabs : nullptr));
    } else if (small) {
      EF_TRY(ef::match_small(proj, m->k, B, m->k, m->gp.as<double>(), m->kpad, m->gnorm.as<double>(), m->n_gallery,
                             labels, m->metric, threshold, sumsq, m->c0, want_resid ? out->resid2 : nullptr, out->score,
                             out->index, out->label, st));
    } else {
      EF_TRY(ef::match(proj, m->k, B, m->k, m->gp.as<double>(), m->kpad, m->gnorm.as<double>(), m->n_gallery, 0,
                       m->metric, out->score, m->index64.as<int64_t>(), m->match_work.p, st));
      EF_TRY(ef::label_lookup(out->score, m->index64.as<int64_t>(), B, labels, m->metric, threshold, out->index,
                              out->label, st));
    }
  }
  m->dirty = false;
  return EF_OK;
}

static int pipe_launch(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                       const ef_result_t* out, cudaStream_t st) {
  const int kr = m->kpad;
  const int slot = m->pending.slot ^ 1;              // rows of the batch submitted now go to the other buffer
  const int cap = std::max(B, m->pending.B);
  if (cap > m->carry_reserved) {
    // growing the carry buffers would lose the pending rows: finish the pending batch with an empty launch first
    if (m->pending.B > 0) EF_TRY(pipe_launch(m, nullptr, 0, 0, 0.0, nullptr, st));
    EF_CUDA(cudaStreamSynchronize(st));
    const size_t rows = (size_t)ef::round_up(cap, 128);
    for (int i = 0; i < 2; ++i) {
      EF_TRY(m->carry_pe[i].ensure(sizeof(double) * rows * kr));
      EF_TRY(m->carry_pn[i].ensure(sizeof(double) * rows));
      EF_TRY(m->carry_img[i].ensure(2 * rows * (size_t)ef::filter_kf(m->k)));
    }
    m->carry_reserved = (int)rows;
  }
  const bool want_resid = B > 0 && out->resid2 != nullptr;
  const double* sumsq_ext = nullptr;
  if (want_resid && m->has_scale) {
    EF_TRY(ef::row_sumsq(x, ldx, B, m->D, m->qq.as<double>(), m->sumsq_w.as<double>(), st));
    sumsq_ext = m->sumsq_w.as<double>();
  }
  const int32_t* labels = m->labels.p ? m->labels.as<int32_t>() : nullptr;
  const ef_model::Pending& pv = m->pending;
  EF_TRY(ef::recognize_pipe(
      x, ldx, B, m->D, m->wq.as<int8_t>(), m->ldw, m->NC, m->nc_pad, m->k, m->kq, m->S, m->col_exp.as<int32_t>(),
      m->bias.as<double>(), sumsq_ext, want_resid, m->c0, B > 0 ? out->proj : nullptr, want_resid ? out->resid2 : nullptr,
      m->carry_pe[slot].as<double>(), m->carry_pn[slot].as<double>(), m->carry_img[slot].p, pv.B,
      m->carry_pe[pv.slot].as<double>(), m->carry_pn[pv.slot].as<double>(), m->carry_img[pv.slot].p, pv.score, pv.index,
      pv.label, pv.threshold, m->gp.as<double>(), m->kpad, m->gnorm.as<double>(), m->ginv.as<double>(), m->gimg.p,
      m->n_gallery, labels, m->metric, m->status.as<int>(), st));
  m->pending.B = B;
  m->pending.slot = slot;
  m->pending.stream = st;
  if (B > 0) {
    m->pending.score = out->score;
    m->pending.index = out->index;
    m->pending.label = out->label;
    m->pending.threshold = threshold;
  }
  m->last_used_tc = true;
  m->last_path = 3;
  return EF_OK;
}

// Launch the queued batches: ONE persistent kernel over all of them, on the stream they were submitted on.
static int queue_launch(ef_model_t* m) {
  if (m->queue.empty()) return EF_OK;
  const int32_t* labels = m->labels.p ? m->labels.as<int32_t>() : nullptr;
  const int nb = (int)m->queue.size();
  const int stq = ef::recognize_stream(m->queue.data(), nb, m->D, m->wq_fm.as<int8_t>(), m->ldw, m->nc_fm, m->k, m->kq,
                                       m->S, m->col_exp.as<int32_t>(), m->bias.as<double>(), m->c0, m->gp.as<double>(),
                                       m->kpad, m->gnorm.as<double>(), m->ginv.as<double>(), m->gimg.p, m->n_gallery,
                                       labels, m->metric, m->status.as<int>(), m->queue_stream);
  m->queue.clear();
  if (stq == EF_OK) {
    m->last_used_tc = true;
    m->last_path = 4;
    if (!m->queue_ev) EF_CUDA(cudaEventCreateWithFlags(&m->queue_ev, cudaEventDisableTiming));
    EF_CUDA(cudaEventRecord(m->queue_ev, m->queue_stream));
    m->queue_ev_pending = true;
  }
  return stq;
}

int ef_model_set_serving(ef_model_t* m, int32_t kernel, int32_t queue_depth) {
  if (!m || kernel < 0 || kernel > 1) return EF_ERR_INVALID;
  if (m->pending.B > 0 || !m->queue.empty()) return EF_ERR_INVALID;      // flush first
  m->serving_kernel = kernel;
  if (queue_depth > 0) m->queue_depth = queue_depth > ef::kStreamMaxBatches ? ef::kStreamMaxBatches : queue_depth;
  m->queue_adaptive = queue_depth >= 0;               // negative: fixed depth (launch only when full or flushed)
  if (queue_depth < 0) m->queue_depth = -queue_depth > ef::kStreamMaxBatches ? ef::kStreamMaxBatches : -queue_depth;
  return EF_OK;
}

int ef_model_submit_device(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                           const ef_result_t* out, ef_stream_t stream) {
  if (m && B == 0) return EF_OK;
  if (!m || !x || !out || B < 0 || ldx < m->D || !out->score || !out->index) return EF_ERR_INVALID;
  if (out->resid2 && !m->with_residual) return EF_ERR_INVALID;
  const bool aligned = !(ldx & 15) && !(reinterpret_cast<uintptr_t>(x) & 15);
  cudaStream_t st = ef::as_stream(stream);
  if (m->serving_kernel == 0 && m->tc_mode >= 2 && aligned && m->gimg.p && m->stream_ok) {
    // persistent serving kernel: the batch joins the queue; the launch happens when the queue is full or at flush
    if (m->pending.B > 0) EF_TRY(ef_model_flush_device(m, stream));
    if (!m->queue.empty() && m->queue_stream != st) EF_TRY(ef_model_flush_device(m, stream));
    if (!m->status.p) EF_TRY(ef_model_reserve(m, 128));
    ef::StreamBatchDesc d{};
    d.x = x; d.ldx = ldx; d.B = B;
    d.out_proj = out->proj; d.out_resid = out->resid2; d.out_score = out->score; d.out_index = out->index;
    d.out_label = out->label; d.threshold = threshold;
    if (out->resid2 && m->has_scale) {
      ef::DevBuf& sq = m->sumsq_q[m->queue.size()];
      EF_TRY(sq.ensure(sizeof(double) * ((size_t)B + 32)));
      EF_TRY(ef::row_sumsq(x, ldx, B, m->D, m->qq.as<double>(), sq.as<double>(), st));
      d.sumsq_ext = sq.as<double>();
    }
    EF_TRY(ef::stream_encode_batch(&d, m->D));
    m->queue.push_back(d);
    m->queue_stream = st;
    if ((int)m->queue.size() >= m->queue_depth) return queue_launch(m);
    // adaptive depth: while the previous launch is still running the queue keeps growing (nothing would start earlier
    // anyway); once it has finished the queued batches go out at once, so a stream that starts from an idle GPU is
    // not held back until queue_depth batches have been submitted
    if (m->queue_ev_pending && cudaEventQuery(m->queue_ev) == cudaSuccess) m->queue_ev_pending = false;
    // (a launch costs ~14 us on top of ~8.4 us per batch: a lone batch waits for a second one, or for the flush)
    if (m->queue_adaptive && !m->queue_ev_pending && m->queue.size() >= 2) return queue_launch(m);
    return EF_OK;
  }
  if (m->tc_mode < 2 || !aligned || !m->gimg.p || !ef::pipe_supported(m->k, m->NC, m->metric, m->n_gallery))
    return ef_model_recognize_device(m, x, ldx, B, threshold, out, stream);    // not pipelined: results right away
  EF_TRY(ef_model_reserve(m, B));
  if (m->pending.B > 0 && m->pending.stream != st) EF_TRY(ef_model_flush_device(m, stream));
  return pipe_launch(m, x, ldx, B, threshold, out, st);
}

int ef_model_flush_device(ef_model_t* m, ef_stream_t stream) {
  if (!m) return EF_ERR_INVALID;
  if (!m->queue.empty()) {
    cudaStream_t want = ef::as_stream(stream), owner = m->queue_stream;
    EF_TRY(queue_launch(m));
    if (want != owner) {
      if (!m->flush_ev) EF_CUDA(cudaEventCreateWithFlags(&m->flush_ev, cudaEventDisableTiming));
      EF_CUDA(cudaEventRecord(m->flush_ev, owner));
      EF_CUDA(cudaStreamWaitEvent(want, m->flush_ev, 0));
    }
  }
  if (m->pending.B <= 0) return EF_OK;
  // The matching launch reads the rows carried by the submit launch, so it runs on the stream of that submit; when the
  // caller flushes from another stream (e.g. a host-buffer call on the model's own stream), that stream is made to wait.
  cudaStream_t want = ef::as_stream(stream), owner = m->pending.stream;
  EF_TRY(pipe_launch(m, nullptr, 0, 0, 0.0, nullptr, owner));
  if (want != owner) {
    if (!m->flush_ev) EF_CUDA(cudaEventCreateWithFlags(&m->flush_ev, cudaEventDisableTiming));
    EF_CUDA(cudaEventRecord(m->flush_ev, owner));
    EF_CUDA(cudaStreamWaitEvent(want, m->flush_ev, 0));
  }
  return EF_OK;
}

int ef_model_status(ef_model_t* m, int32_t* tc_pipeline_timeouts) {
  if (!m || !tc_pipeline_timeouts) return EF_ERR_INVALID;
  *tc_pipeline_timeouts = 0;
  if (!m->status.p) return EF_OK;
  EF_CUDA(cudaMemcpy(tc_pipeline_timeouts, m->status.p, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return EF_OK;
}

static int host_reserve(ef_model_t* m, int32_t B) {
  EF_TRY(ef_model_reserve(m, B));
  if (B <= m->host_reserved) return EF_OK;
  m->x_ld = ef::round_up(m->D, 16);     // dense when D % 16 == 0: contiguous host batches go up as ONE 1-D copy per chunk
  EF_TRY(m->x_dev.ensure((size_t)B * m->x_ld));
  EF_TRY(m->resid_dev.ensure(sizeof(double) * (size_t)B));
  EF_TRY(m->index32_dev.ensure(sizeof(int32_t) * (size_t)B));
  EF_TRY(m->label_dev.ensure(sizeof(int32_t) * (size_t)B));
  if (!m->bad_dev.p) {
    EF_TRY(m->bad_dev.ensure(16));
    EF_CUDA(cudaMemset(m->bad_dev.p, 0, 16));
    EF_CUDA(cudaDeviceSynchronize());
  }
  m->host_reserved = B;
  return EF_OK;
}

// The caller's arrays are ordinary (pageable) host memory: a device->host copy straight into them is staged and
// synchronised by the driver one array at a time.  Everything goes into one page-locked block instead (asynchronous, back
// to back); after ONE synchronisation the block is scattered with memcpy.
struct ResultBlock {
  double *proj, *score, *resid;
  int32_t *index, *label, *flag;
};

static size_t result_block_bytes(const ef_model_t* m, size_t nB) {
  return sizeof(double) * nB * (size_t)m->k + sizeof(double) * nB * 2 + sizeof(int32_t) * nB * 2 + 64;
}

static ResultBlock carve(const ef_model_t* m, void* pinned, size_t nB) {
  char* p = reinterpret_cast<char*>(pinned);
  ResultBlock r;
  r.proj = reinterpret_cast<double*>(p);   p += sizeof(double) * nB * (size_t)m->k;
  r.score = reinterpret_cast<double*>(p);  p += sizeof(double) * nB;
  r.resid = reinterpret_cast<double*>(p);  p += sizeof(double) * nB;
  r.index = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * nB;
  r.label = reinterpret_cast<int32_t*>(p); p += sizeof(int32_t) * nB;
  r.flag = reinterpret_cast<int32_t*>(p);
  return r;
}

static int ensure_pinned(void** pinned, size_t* have, size_t need) {
  if (need <= *have) return EF_OK;
  if (*pinned) cudaFreeHost(*pinned);
  *pinned = nullptr;
  *have = 0;
  EF_CUDA(cudaMallocHost(pinned, need));
  *have = need;
  return EF_OK;
}

static int enqueue_results(ef_model_t* m, int32_t B, bool proj, bool resid, bool label, const ef_result_t& dev,
                           void* pinned) {
  cudaStream_t st = m->stream;
  const size_t nB = (size_t)B;
  const ResultBlock h = carve(m, pinned, nB);
  if (proj) EF_CUDA(cudaMemcpyAsync(h.proj, dev.proj, sizeof(double) * nB * m->k, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(h.score, dev.score, sizeof(double) * nB, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(h.index, dev.index, sizeof(int32_t) * nB, cudaMemcpyDeviceToHost, st));
  if (label) EF_CUDA(cudaMemcpyAsync(h.label, dev.label, sizeof(int32_t) * nB, cudaMemcpyDeviceToHost, st));
  if (resid) EF_CUDA(cudaMemcpyAsync(h.resid, dev.resid2, sizeof(double) * nB, cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemcpyAsync(h.flag, m->status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return EF_OK;
}

static int scatter_results(const ef_model_t* m, int32_t B, const ef_result_t* out, const void* pinned) {
  const size_t nB = (size_t)B;
  const ResultBlock h = carve(m, const_cast<void*>(pinned), nB);
  if (out->proj) memcpy(out->proj, h.proj, sizeof(double) * nB * m->k);
  if (out->score) memcpy(out->score, h.score, sizeof(double) * nB);
  if (out->index) memcpy(out->index, h.index, sizeof(int32_t) * nB);
  if (out->label) memcpy(out->label, h.label, sizeof(int32_t) * nB);
  if (out->resid2) memcpy(out->resid2, h.resid, sizeof(double) * nB);
  if (*h.flag) {
    ef::set_error_detail("tcgen05 projection pipeline timed out (mbarrier wait > 2 s)", cudaErrorLaunchTimeout);
    return EF_ERR_CUDA;
  }
  return EF_OK;
}

static int copy_results_back(ef_model_t* m, int32_t B, const ef_result_t* out, const ef_result_t& dev) {
  EF_TRY(ensure_pinned(&m->pinned, &m->pinned_bytes, result_block_bytes(m, (size_t)B)));
  EF_TRY(enqueue_results(m, B, out->proj != nullptr, out->resid2 != nullptr, out->label != nullptr, dev, m->pinned));
  EF_CUDA(cudaStreamSynchronize(m->stream));
  return scatter_results(m, B, out, m->pinned);
}

// Crops of a host batch -> xbuf in chunks on the copy stream, every chunk recognised as soon as it has landed (only the
// last chunk's kernel is exposed behind the PCIe transfer, which dominates this path).
static int enqueue_host_batch(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold, bool proj,
                              bool resid, uint8_t* xbuf, std::vector<cudaEvent_t>& evs, ef_result_t* dev_out) {
  cudaStream_t st = m->stream;
  ef_result_t dev;
  dev.proj = proj ? m->proj.as<double>() : nullptr;
  dev.score = m->score.as<double>();
  dev.index = m->index32_dev.as<int32_t>();
  dev.label = m->label_dev.as<int32_t>();
  dev.resid2 = resid ? m->resid_dev.as<double>() : nullptr;
  constexpr int kChunk = 1024;
  const int n_chunks = B >= 2 * kChunk ? (B + kChunk - 1) / kChunk : 1;
  if (!m->copy_stream) EF_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  while ((int)evs.size() < n_chunks + 1) {
    cudaEvent_t e;
    EF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    evs.push_back(e);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int b0 = c * kChunk, rows = n_chunks == 1 ? B : std::min(kChunk, B - b0);
    cudaStream_t cs = m->copy_stream;
    uint8_t* xd = xbuf + (size_t)b0 * m->x_ld;
    const uint8_t* xh = x + (size_t)b0 * ldx;
    if (ldx == m->x_ld) {
      EF_CUDA(cudaMemcpyAsync(xd, xh, (size_t)rows * ldx, cudaMemcpyHostToDevice, cs));
    } else {
      EF_CUDA(cudaMemcpy2DAsync(xd, m->x_ld, xh, ldx, m->D, rows, cudaMemcpyHostToDevice, cs));
    }
    EF_CUDA(cudaEventRecord(evs[c], cs));
    EF_CUDA(cudaStreamWaitEvent(st, evs[c], 0));
    ef_result_t part;
    part.proj = dev.proj ? dev.proj + (size_t)b0 * m->k : nullptr;
    part.score = dev.score + b0;
    part.index = dev.index + b0;
    part.label = dev.label + b0;
    part.resid2 = dev.resid2 ? dev.resid2 + b0 : nullptr;
    EF_TRY(ef_model_recognize_device(m, xd, m->x_ld, rows, threshold, &part, st));
  }
  *dev_out = dev;
  return EF_OK;
}

int ef_model_recognize_host(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold,
                            const ef_result_t* out) {
  if (m && B == 0) return EF_OK;
  if (!m || !x || !out || B < 0 || ldx < m->D) return EF_ERR_INVALID;
  if (out->resid2 && !m->with_residual) return EF_ERR_INVALID;
  if (m->hslot[0].busy || m->hslot[1].busy) return EF_ERR_INVALID;     // asynchronous batches in flight: wait first
  EF_TRY(host_reserve(m, B));
  ef_result_t dev;
  EF_TRY(enqueue_host_batch(m, x, ldx, B, threshold, out->proj != nullptr, out->resid2 != nullptr,
                            m->x_dev.as<uint8_t>(), m->chunk_ev, &dev));
  return copy_results_back(m, B, out, dev);
}

int ef_model_submit_host(ef_model_t* m, const uint8_t* x, int64_t ldx, int32_t B, double threshold, int32_t want,
                         int32_t* ticket) {
  if (!m || !x || !ticket || B <= 0 || ldx < m->D) return EF_ERR_INVALID;
  const bool proj = want & 1, resid = want & 2, label = want & 4;
  if (resid && !m->with_residual) return EF_ERR_INVALID;
  auto& hs = m->hslot[m->next_slot];
  if (hs.busy) return EF_ERR_INVALID;                                  // both slots in flight: ef_model_wait_host first
  EF_TRY(host_reserve(m, B));
  EF_TRY(hs.x.ensure((size_t)B * m->x_ld));
  EF_TRY(ensure_pinned(&hs.pinned, &hs.pinned_bytes, result_block_bytes(m, (size_t)B)));
  if (!hs.done) EF_CUDA(cudaEventCreateWithFlags(&hs.done, cudaEventDisableTiming));
  ef_result_t dev;
  EF_TRY(enqueue_host_batch(m, x, ldx, B, threshold, proj, resid, hs.x.as<uint8_t>(), hs.chunk_ev, &dev));
  EF_TRY(enqueue_results(m, B, proj, resid, label, dev, hs.pinned));
  EF_CUDA(cudaEventRecord(hs.done, m->stream));
  hs.B = B;
  hs.busy = true;
  hs.want_proj = proj;
  hs.want_resid = resid;
  hs.want_label = label;
  *ticket = m->next_slot;
  m->next_slot ^= 1;
  return EF_OK;
}

int ef_model_wait_host(ef_model_t* m, int32_t ticket, const ef_result_t* out) {
  if (!m || !out || ticket < 0 || ticket > 1) return EF_ERR_INVALID;
  auto& hs = m->hslot[ticket];
  if (!hs.busy) return EF_ERR_INVALID;
  if ((out->proj && !hs.want_proj) || (out->resid2 && !hs.want_resid) || (out->label && !hs.want_label))
    return EF_ERR_INVALID;
  EF_CUDA(cudaEventSynchronize(hs.done));
  hs.busy = false;
  return scatter_results(m, hs.B, out, hs.pinned);
}

int ef_model_recognize_boxes_device(ef_model_t* m, const uint8_t* frames, int64_t frame_stride, int32_t pitch,
                                    int32_t width, int32_t height, int32_t channels, int32_t n_frames,
                                    const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh, double threshold,
                                    const ef_result_t* out, ef_stream_t stream) {
  if (!m || !out || n_boxes < 0) return EF_ERR_INVALID;
  if ((int64_t)dw * dh != m->D) return EF_ERR_INVALID;
  if (n_boxes == 0) return EF_OK;
  EF_TRY(host_reserve(m, n_boxes));
  // boxes outside their frame (or with w, h <= 0, or a bad frame index) become all-zero crops AND are counted: the host
  // entry point turns a non-zero count into EF_ERR_INVALID, device callers read it with ef_model_bad_boxes
  EF_TRY(ef_preprocess(frames, frame_stride, pitch, width, height, channels, n_frames, boxes, n_boxes, dw, dh,
                       m->x_dev.as<uint8_t>(), m->x_ld, m->bad_dev.as<int32_t>(), stream));
  return ef_model_recognize_device(m, m->x_dev.as<uint8_t>(), m->x_ld, n_boxes, threshold, out, stream);
}

int ef_model_bad_boxes(ef_model_t* m, ef_stream_t stream, int32_t* count) {
  if (!m || !count) return EF_ERR_INVALID;
  *count = 0;
  if (!m->bad_dev.p) return EF_OK;
  cudaStream_t st = ef::as_stream(stream);
  EF_CUDA(cudaMemcpyAsync(count, m->bad_dev.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemsetAsync(m->bad_dev.p, 0, sizeof(int32_t), st));
  EF_CUDA(cudaStreamSynchronize(st));
  return EF_OK;
}

int ef_model_recognize_boxes_host(ef_model_t* m, const uint8_t* frames, int64_t frame_stride, int32_t pitch,
                                  int32_t width, int32_t height, int32_t channels, int32_t n_frames,
                                  const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh, double threshold,
                                  const ef_result_t* out) {
  if (!m || !frames || !boxes || !out || n_boxes < 0 || n_frames <= 0) return EF_ERR_INVALID;
  if (n_boxes == 0) return EF_OK;
  EF_TRY(host_reserve(m, n_boxes));
  const size_t frame_bytes = (size_t)frame_stride * n_frames;
  EF_TRY(m->frames_dev.ensure(frame_bytes));
  EF_TRY(m->boxes_dev.ensure(sizeof(ef_box_t) * (size_t)n_boxes));
  cudaStream_t st = m->stream;
  EF_CUDA(cudaMemcpyAsync(m->frames_dev.p, frames, frame_bytes, cudaMemcpyHostToDevice, st));
  EF_CUDA(cudaMemcpyAsync(m->boxes_dev.p, boxes, sizeof(ef_box_t) * (size_t)n_boxes, cudaMemcpyHostToDevice, st));
  ef_result_t dev;
  dev.proj = out->proj ? m->proj.as<double>() : nullptr;
  dev.score = m->score.as<double>();
  dev.index = m->index32_dev.as<int32_t>();
  dev.label = m->label_dev.as<int32_t>();
  dev.resid2 = out->resid2 ? m->resid_dev.as<double>() : nullptr;
  EF_TRY(ef_model_recognize_boxes_device(m, m->frames_dev.as<uint8_t>(), frame_stride, pitch, width, height, channels,
                                         n_frames, m->boxes_dev.as<ef_box_t>(), n_boxes, dw, dh, threshold, &dev, st));
  EF_TRY(ensure_pinned(&m->pinned, &m->pinned_bytes, result_block_bytes(m, (size_t)n_boxes)));
  EF_TRY(enqueue_results(m, n_boxes, out->proj != nullptr, out->resid2 != nullptr, out->label != nullptr, dev, m->pinned));
  int32_t* bad_host = carve(m, m->pinned, (size_t)n_boxes).flag + 1;
  EF_CUDA(cudaMemcpyAsync(bad_host, m->bad_dev.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaMemsetAsync(m->bad_dev.p, 0, sizeof(int32_t), st));
  EF_CUDA(cudaStreamSynchronize(st));
  if (*bad_host != 0) {
    // the reference would clip the slice or raise inside cv2.resize; a black crop with a normal-looking label is never
    // returned silently
    char msg[96];
    snprintf(msg, sizeof(msg), "%d of %d boxes are not inside their frame", *bad_host, n_boxes);
    ef::set_error_detail(msg, cudaErrorInvalidValue);
    return EF_ERR_INVALID;
  }
  return scatter_results(m, n_boxes, out, m->pinned);
}

// M3 in one call: every person's model on the same boxes (recognize_face_all_models, scan-template-v4.py:289-319, calls
// extract_face_features + recognize_face_with_model once per model and per face).  Frames and boxes go up once, K1 runs
// once, every model's K2 writes its (score, index, label) rows into ONE device block, ONE device->host copy and ONE
// synchronisation bring everything back: the per-face latency is a handful of launches, not a Python loop of tensors.
int ef_models_recognize_boxes_host(ef_model_t* const* models, int32_t n_models, const uint8_t* frames,
                                   int64_t frame_stride, int32_t pitch, int32_t width, int32_t height, int32_t channels,
                                   int32_t n_frames, const ef_box_t* boxes, int32_t n_boxes, int32_t dw, int32_t dh,
                                   double threshold, double* score, int32_t* index, int32_t* label) {
  if (!models || n_models <= 0 || !frames || !boxes || !score || !index || !label || n_boxes < 0 || n_frames <= 0)
    return EF_ERR_INVALID;
  if (n_boxes == 0) return EF_OK;
  ef_model_t* m0 = models[0];
  for (int i = 0; i < n_models; ++i) {
    if (!models[i] || (int64_t)dw * dh != models[i]->D) return EF_ERR_INVALID;
    if (models[i]->pending.B > 0 || !models[i]->queue.empty()) return EF_ERR_INVALID;     // flush the serving queue first
    EF_TRY(host_reserve(models[i], n_boxes));
    if (i > 0) EF_CUDA(cudaStreamSynchronize(models[i]->stream));     // its own (normally idle) stream owes us nothing
  }
  cudaStream_t st = m0->stream;
  const size_t nB = (size_t)n_boxes, rows = (size_t)n_models * nB;
  const size_t frame_bytes = (size_t)frame_stride * n_frames;
  const size_t block = rows * (sizeof(double) + 2 * sizeof(int32_t)) + 16;
  EF_TRY(m0->frames_dev.ensure(frame_bytes));
  EF_TRY(m0->boxes_dev.ensure(sizeof(ef_box_t) * nB));
  EF_TRY(m0->multi_dev.ensure(block));
  const size_t status_off = ef::round_up((int64_t)block, 16);
  EF_TRY(ensure_pinned(&m0->pinned, &m0->pinned_bytes,
                       std::max(status_off + sizeof(int32_t) * (size_t)n_models, result_block_bytes(m0, nB))));
  EF_CUDA(cudaMemcpyAsync(m0->frames_dev.p, frames, frame_bytes, cudaMemcpyHostToDevice, st));
  EF_CUDA(cudaMemcpyAsync(m0->boxes_dev.p, boxes, sizeof(ef_box_t) * nB, cudaMemcpyHostToDevice, st));
  char* d = m0->multi_dev.as<char>();
  double* d_score = reinterpret_cast<double*>(d);
  int32_t* d_index = reinterpret_cast<int32_t*>(d + rows * sizeof(double));
  int32_t* d_label = d_index + rows;
  int32_t* d_bad = d_label + rows;                                    // 4-byte aligned tail word
  EF_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int32_t), st));
  EF_TRY(ef_preprocess(m0->frames_dev.as<uint8_t>(), frame_stride, pitch, width, height, channels, n_frames,
                       m0->boxes_dev.as<ef_box_t>(), n_boxes, dw, dh, m0->x_dev.as<uint8_t>(), m0->x_ld, d_bad, st));
  // A handful of crops (the reference's pattern: ONE face per call) leaves the GPU almost empty and every model's chain of
  // 3-4 kernels is latency: the models then run side by side, each on its own stream, between two events (crops ready /
  // chain done).  Large batches fill the GPU model by model and stay on one stream.
  const bool fan = n_models > 1 && n_boxes <= 256 && !getenv("EF_NO_FANOUT_STREAMS");
  if (fan) {
    if (!m0->fan_ev) EF_CUDA(cudaEventCreateWithFlags(&m0->fan_ev, cudaEventDisableTiming));
    EF_CUDA(cudaEventRecord(m0->fan_ev, st));
  }
  for (int i = 0; i < n_models; ++i) {
    ef_result_t dev{};
    dev.score = d_score + (size_t)i * nB;
    dev.index = d_index + (size_t)i * nB;
    dev.label = d_label + (size_t)i * nB;
    cudaStream_t si = (fan && i > 0) ? models[i]->stream : st;
    if (si != st) EF_CUDA(cudaStreamWaitEvent(si, m0->fan_ev, 0));
    EF_TRY(ef_model_recognize_device(models[i], m0->x_dev.as<uint8_t>(), m0->x_ld, n_boxes, threshold, &dev, si));
    if (si != st) {
      if (!models[i]->fan_ev) EF_CUDA(cudaEventCreateWithFlags(&models[i]->fan_ev, cudaEventDisableTiming));
      EF_CUDA(cudaEventRecord(models[i]->fan_ev, si));
    }
  }
  if (fan)
    for (int i = 1; i < n_models; ++i)
      if (models[i]->stream != st) EF_CUDA(cudaStreamWaitEvent(st, models[i]->fan_ev, 0));
  EF_CUDA(cudaMemcpyAsync(m0->pinned, d, block, cudaMemcpyDeviceToHost, st));
  int32_t* h_status = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(m0->pinned) + status_off);
  for (int i = 0; i < n_models; ++i)                                  // tcgen05 pipeline-timeout flags of the models
    EF_CUDA(cudaMemcpyAsync(h_status + i, models[i]->status.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  EF_CUDA(cudaStreamSynchronize(st));
  const char* h = reinterpret_cast<const char*>(m0->pinned);
  const int32_t bad = *reinterpret_cast<const int32_t*>(h + rows * (sizeof(double) + 2 * sizeof(int32_t)));
  if (bad != 0) {
    char msg[96];
    snprintf(msg, sizeof(msg), "%d of %d boxes are not inside their frame", bad, n_boxes);
    ef::set_error_detail(msg, cudaErrorInvalidValue);
    return EF_ERR_INVALID;
  }
  for (int i = 0; i < n_models; ++i)
    if (h_status[i] != 0) {                                           // a pipeline timed out: never garbage as a result
      ef::set_error_detail("tcgen05 pipeline timeout (ef_models_recognize_boxes_host)", cudaErrorLaunchTimeout);
      return EF_ERR_CUDA;
    }
  memcpy(score, h, rows * sizeof(double));
  memcpy(index, h + rows * sizeof(double), rows * sizeof(int32_t));
  memcpy(label, h + rows * (sizeof(double) + sizeof(int32_t)), rows * sizeof(int32_t));
  return EF_OK;
}

}  // extern "C"
