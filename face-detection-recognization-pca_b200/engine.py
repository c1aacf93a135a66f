"""Batched recognition engine and PCA fit on top of the C ABI (host-side mirror, no arithmetic here).

`Recognizer` wraps one ef_model handle: a model dict of either reference generation goes in once, batches of
crops go through `recognize` (host numpy buffers, copies inside) or `recognize_device` (torch CUDA tensors,
zero-copy).  `fit_gen1` / `fit_gen2` wrap the device PCA fits.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import METRIC_COSINE_G1, METRIC_COSINE_SK, METRIC_L2, Box, FitInfo, Gen2Fit, ModelDesc, Result, check


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RecognitionResult:
    """Arrays for a batch: features [B,k], score [B], index [B], label [B], resid2 [B] or None."""
    __slots__ = ("features", "score", "index", "label", "resid2")

    def __init__(self, features, score, index, label, resid2):
        self.features, self.score, self.index, self.label, self.resid2 = features, score, index, label, resid2


class Recognizer:
    """Device-resident recognition model.

    basis: [D, k] (any strides; the Gen-1 pickle stores it Fortran-ordered) or, with basis_is_components=True,
    sklearn's components_ [k, D].  mean [D]; scale / pca_mean [D] or None; gallery [Ng, k]; labels [Ng] or None.
    """

    def __init__(self, basis, mean, gallery, *, scale=None, pca_mean=None, labels=None, metric=METRIC_COSINE_G1,
                 basis_is_components=False, n_slices=0, with_residual=True):
        L = _lib.lib()
        basis = np.asarray(basis)
        if basis.dtype != np.float64:
            basis = basis.astype(np.float64)           # float32 pickles upcast exactly (numpy does the same)
        if basis_is_components:
            basis = basis.T                             # view [D, k]
        if basis.ndim != 2:
            raise ValueError("basis must be 2-D")
        D, k = basis.shape
        if not (basis.flags["C_CONTIGUOUS"] or basis.flags["F_CONTIGUOUS"]):
            basis = np.asfortranarray(basis)
        sd, sk = basis.strides[0] // 8, basis.strides[1] // 8
        self._keep = [basis, _f64(mean), _f64(gallery)]
        mean_a, gal = self._keep[1], self._keep[2]
        if mean_a.shape != (D,) or gal.ndim != 2 or gal.shape[1] != k:
            raise ValueError(f"shape mismatch: basis {basis.shape}, mean {mean_a.shape}, gallery {gal.shape}")
        scale_a = _f64(scale) if scale is not None else None
        pm_a = _f64(pca_mean) if pca_mean is not None else None
        lab_a = np.ascontiguousarray(np.asarray(labels, dtype=np.int32)) if labels is not None else None
        desc = ModelDesc(D=D, k=k, basis=_ptr(basis), basis_stride_d=sd, basis_stride_k=sk, mean=_ptr(mean_a),
                         scale=_ptr(scale_a), pca_mean=_ptr(pm_a), gallery=_ptr(gal), gallery_ld=k,
                         n_gallery=gal.shape[0], labels=_ptr(lab_a), metric=metric, n_slices=n_slices,
                         with_residual=1 if with_residual else 0)
        handle = C.c_void_p()
        check(L.ef_model_create(C.byref(handle), C.byref(desc)), "ef_model_create")
        self._h, self._L = handle, L
        self._held = []
        self.D, self.k, self.n_gallery, self.metric = D, k, gal.shape[0], metric
        self.with_residual = with_residual
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._L.ef_model_destroy(self._h)
            self._h = None

    __del__ = close

    def reserve(self, max_batch):
        check(self._L.ef_model_reserve(self._h, int(max_batch)), "ef_model_reserve")

    def use_tensor_cores(self, mode=2):
        """2 (default): single cluster kernel; 1 / True: tcgen05 stream-K projection + epilogue kernels; 0 / False: dp4a."""
        check(self._L.ef_model_set_tensor_cores(self._h, int(mode)), "ef_model_set_tensor_cores")

    def set_serving(self, kernel=0, queue_depth=0):
        """submit_device's kernel: 0 = persistent queue kernel (default), 1 = pipelined kernel (one launch per submit);
        queue_depth 1..32 batches per persistent launch (0 keeps the current value, -d = fixed depth d without the
        adaptive early launch)."""
        check(self._L.ef_model_set_serving(self._h, int(kernel), int(queue_depth)), "ef_model_set_serving")

    def pipeline_timeouts(self):
        """Non-zero when the tcgen05 projection kernel hit a bounded-wait timeout (synchronous read)."""
        n = C.c_int32()
        check(self._L.ef_model_status(self._h, C.byref(n)), "ef_model_status")
        return n.value

    def serving_path(self):
        """Kernel path the last recognise / submit call took: 4 persistent stream kernel, 3 pipelined kernel, 2 single
        cluster kernel, 1 tcgen05 projection + epilogue kernels, 0 dp4a."""
        return self.kernel_timing_read()[2]

    def kernel_timing(self, enable=True):
        check(self._L.ef_model_kernel_timing(self._h, 1 if enable else 0), "ef_model_kernel_timing")

    def kernel_timing_read(self):
        """(calls, mean projection-kernel milliseconds, used_tensor_cores) since kernel_timing(True)."""
        n, ms, tc = C.c_int32(), C.c_double(), C.c_int32()
        check(self._L.ef_model_kernel_timing_read(self._h, C.byref(n), C.byref(ms), C.byref(tc)),
              "ef_model_kernel_timing_read")
        return n.value, ms.value, tc.value

    # ------------------------------------------------------------------ host buffers (numpy in, numpy out)
    def recognize(self, crops, threshold=0.7, want_features=True, want_residual=None):
        """crops: uint8 [B, D] (row pitch arbitrary).  Copies in, runs K2, copies results out (synchronous)."""
        x = np.asarray(crops)
        if x.dtype != np.uint8 or x.ndim != 2 or x.shape[1] != self.D:
            raise ValueError(f"crops must be uint8 [B, {self.D}], got {x.dtype} {x.shape}")
        if x.strides[1] != 1:
            x = np.ascontiguousarray(x)
        B = x.shape[0]
        want_residual = self.with_residual if want_residual is None else want_residual
        feats = np.empty((B, self.k), dtype=np.float64) if want_features else None
        score = np.empty(B, dtype=np.float64)
        index = np.empty(B, dtype=np.int32)
        label = np.empty(B, dtype=np.int32)
        resid = np.empty(B, dtype=np.float64) if want_residual else None
        res = Result(_ptr(feats), _ptr(score), _ptr(index), _ptr(label), _ptr(resid))
        ldx = x.strides[0] if B > 1 else self.D          # numpy gives length-1 axes arbitrary strides
        check(self._L.ef_model_recognize_host(self._h, _ptr(x), ldx, B, float(threshold), C.byref(res)),
              "ef_model_recognize_host")
        return RecognitionResult(feats, score, index, label, resid)

    def submit(self, crops, threshold=0.7, want_features=True, want_residual=None):
        """Asynchronous recognize(): enqueues copy-in, kernels and copy-out and returns a ticket for wait().  At most
        two batches may be in flight; the crops array must stay alive (and should be page-locked, e.g. a pinned torch
        tensor's .numpy()) until the matching wait().  The upload of batch i+1 overlaps the kernels and the result copy
        of batch i."""
        x = np.asarray(crops)
        if x.dtype != np.uint8 or x.ndim != 2 or x.shape[1] != self.D:
            raise ValueError(f"crops must be uint8 [B, {self.D}], got {x.dtype} {x.shape}")
        if x.strides[1] != 1:
            x = np.ascontiguousarray(x)
        B = x.shape[0]
        want_residual = self.with_residual if want_residual is None else want_residual
        want = (1 if want_features else 0) | (2 if want_residual else 0) | 4
        ticket = C.c_int32(-1)
        ldx = x.strides[0] if B > 1 else self.D
        check(self._L.ef_model_submit_host(self._h, _ptr(x), ldx, B, float(threshold), want, C.byref(ticket)),
              "ef_model_submit_host")
        return (ticket.value, B, bool(want_features), bool(want_residual), x)      # x kept alive by the ticket

    def wait(self, ticket):
        """Blocks until the batch behind `ticket` is done; returns its RecognitionResult (same values as recognize())."""
        slot, B, want_features, want_residual, _x = ticket
        feats = np.empty((B, self.k), dtype=np.float64) if want_features else None
        score = np.empty(B, dtype=np.float64)
        index = np.empty(B, dtype=np.int32)
        label = np.empty(B, dtype=np.int32)
        resid = np.empty(B, dtype=np.float64) if want_residual else None
        res = Result(_ptr(feats), _ptr(score), _ptr(index), _ptr(label), _ptr(resid))
        check(self._L.ef_model_wait_host(self._h, slot, C.byref(res)), "ef_model_wait_host")
        return RecognitionResult(feats, score, index, label, resid)

    def recognize_boxes(self, frames, boxes, side, threshold=0.7, want_features=True, want_residual=None):
        """frames: uint8 [F, H, W] gray or [F, H, W, 3] BGR (or a single frame); boxes: int [B, 4] (x, y, w, h) or
        [B, 5] (frame, x, y, w, h).  K1 + K2 through host buffers."""
        fr = np.ascontiguousarray(frames)
        if fr.dtype != np.uint8:
            raise ValueError("frames must be uint8")
        if fr.ndim == 2 or (fr.ndim == 3 and fr.shape[2] == 3):
            fr = fr[None]
        channels = 3 if fr.ndim == 4 else 1
        F, H, W = fr.shape[:3]
        bx = np.asarray(boxes, dtype=np.int32).reshape(-1, np.asarray(boxes).shape[-1])
        if bx.shape[1] == 4:
            bx = np.concatenate([np.zeros((len(bx), 1), np.int32), bx], axis=1)
        bx = np.ascontiguousarray(bx)
        B = len(bx)
        if side * side != self.D:
            raise ValueError("side*side must equal the model dimension")
        want_residual = self.with_residual if want_residual is None else want_residual
        feats = np.empty((B, self.k), dtype=np.float64) if want_features else None
        score = np.empty(B, dtype=np.float64)
        index = np.empty(B, dtype=np.int32)
        label = np.empty(B, dtype=np.int32)
        resid = np.empty(B, dtype=np.float64) if want_residual else None
        res = Result(_ptr(feats), _ptr(score), _ptr(index), _ptr(label), _ptr(resid))
        check(self._L.ef_model_recognize_boxes_host(self._h, _ptr(fr), H * W * channels, W * channels, W, H, channels,
                                                    F, _ptr(bx), B, side, side, float(threshold), C.byref(res)),
              "ef_model_recognize_boxes_host")
        return RecognitionResult(feats, score, index, label, resid)

    # ------------------------------------------------------------------ device buffers (torch CUDA tensors)
    def recognize_device(self, x, threshold=0.7, out=None, want_residual=None):
        """x: torch uint8 CUDA tensor [B, ldx>=D] with 16-byte aligned rows.  Enqueues on torch's current stream and
        returns a dict of CUDA tensors (features, score, index, label, resid2)."""
        import torch
        if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 2 and x.stride(1) == 1):
            raise ValueError("x must be a 2-D uint8 CUDA tensor with unit inner stride")
        B = x.shape[0]
        want_residual = self.with_residual if want_residual is None else want_residual
        if out is None:
            out = {
                "features": torch.empty((B, self.k), dtype=torch.float64, device=x.device),
                "score": torch.empty(B, dtype=torch.float64, device=x.device),
                "index": torch.empty(B, dtype=torch.int32, device=x.device),
                "label": torch.empty(B, dtype=torch.int32, device=x.device),
                "resid2": torch.empty(B, dtype=torch.float64, device=x.device) if want_residual else None,
            }
        res = Result(out["features"].data_ptr(), out["score"].data_ptr(), out["index"].data_ptr(),
                     out["label"].data_ptr(), out["resid2"].data_ptr() if out.get("resid2") is not None else None)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        check(self._L.ef_model_recognize_device(self._h, x.data_ptr(), x.stride(0), B, float(threshold), C.byref(res),
                                                C.c_void_p(stream)), "ef_model_recognize_device")
        return out

    def submit_device(self, x, threshold=0.7, out=None, want_residual=None, stream=None):
        """Queued form of recognize_device for a stream of batches (ef_model_submit_device): the batch joins the model's
        queue; one persistent kernel recognises the queued batches back to back when the queue is full, when the
        previous launch has finished, or at flush_device().  x and the returned dict of output tensors must stay alive
        and untouched until then (the Recognizer keeps references to them until the next flush_device).
        stream: raw CUDA stream handle (default: torch's current stream; the lookup costs ~1.5 us per call)."""
        if out is None:
            import torch
            if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 2 and x.stride(1) == 1):
                raise ValueError("x must be a 2-D uint8 CUDA tensor with unit inner stride")
            B = x.shape[0]
            want_residual = self.with_residual if want_residual is None else want_residual
            out = {
                "features": torch.empty((B, self.k), dtype=torch.float64, device=x.device),
                "score": torch.empty(B, dtype=torch.float64, device=x.device),
                "index": torch.empty(B, dtype=torch.int32, device=x.device),
                "label": torch.empty(B, dtype=torch.int32, device=x.device),
                "resid2": torch.empty(B, dtype=torch.float64, device=x.device) if want_residual else None,
            }
        res = out.get("_res")                       # the ctypes view of a re-used output dict is built once
        if res is None:
            import torch
            if not (x.is_cuda and x.dtype == torch.uint8 and x.dim() == 2 and x.stride(1) == 1):
                raise ValueError("x must be a 2-D uint8 CUDA tensor with unit inner stride")
            res = C.byref(Result(out["features"].data_ptr(), out["score"].data_ptr(), out["index"].data_ptr(),
                                 out["label"].data_ptr(),
                                 out["resid2"].data_ptr() if out.get("resid2") is not None else None))
            out["_res"] = res
        if stream is None:
            import torch
            stream = torch.cuda.current_stream(x.device).cuda_stream
        status = self._L.ef_model_submit_device(self._h, x.data_ptr(), x.stride(0), x.shape[0], threshold, res, stream)
        if status:
            check(status, "ef_model_submit_device")
        # the launch may come later (queue): keep the tensors away from torch's caching allocator until the flush
        held = self._held
        held.append((x, out))
        if len(held) > 64:
            del held[:-32]                 # older batches were launched long ago (queue depth <= 32, stream ordered)
        return out

    def flush_device(self, device=None):
        """Completes the batch left pending by submit_device (enqueues on torch's current stream)."""
        import torch
        stream = torch.cuda.current_stream(device).cuda_stream
        check(self._L.ef_model_flush_device(self._h, C.c_void_p(stream)), "ef_model_flush_device")
        self._held = []                    # everything queued has been launched; stream order protects the tensors now

    def bad_boxes(self, device=None):
        """Boxes outside their frame seen by recognize_boxes_device since the last call (synchronises torch's current
        stream; clears the counter).  Their results are those of an all-zero crop and must be discarded."""
        import torch
        n = C.c_int32()
        stream = torch.cuda.current_stream(device).cuda_stream
        check(self._L.ef_model_bad_boxes(self._h, C.c_void_p(stream), C.byref(n)), "ef_model_bad_boxes")
        return n.value

    def check_device_results(self, device=None):
        """Call after synchronising on results of the device entry points: raises when a box lay outside its frame or
        when a tcgen05 pipeline wait timed out (the kernels drain instead of hanging and only set a flag)."""
        bad = self.bad_boxes(device)
        if bad:
            raise _lib.EigenfacesError(_lib.EF_ERR_INVALID, "recognize_boxes_device", f"{bad} boxes are not inside their frame")
        if self.pipeline_timeouts():
            raise _lib.EigenfacesError(_lib.EF_ERR_CUDA, "recognize_device", "tcgen05 pipeline timed out")

    def recognize_boxes_device(self, frames, boxes, side, threshold=0.7, out=None, want_residual=None):
        """frames: torch uint8 CUDA [F, H, W] / [F, H, W, 3]; boxes: torch int32 CUDA [B, 5] (frame, x, y, w, h).
        Asynchronous; boxes outside their frame are counted, see check_device_results()."""
        import torch
        channels = 3 if frames.dim() == 4 else 1
        F, H, W = frames.shape[:3]
        B = boxes.shape[0]
        want_residual = self.with_residual if want_residual is None else want_residual
        if out is None:
            dev = frames.device
            out = {
                "features": torch.empty((B, self.k), dtype=torch.float64, device=dev),
                "score": torch.empty(B, dtype=torch.float64, device=dev),
                "index": torch.empty(B, dtype=torch.int32, device=dev),
                "label": torch.empty(B, dtype=torch.int32, device=dev),
                "resid2": torch.empty(B, dtype=torch.float64, device=dev) if want_residual else None,
            }
        res = Result(out["features"].data_ptr(), out["score"].data_ptr(), out["index"].data_ptr(),
                     out["label"].data_ptr(), out["resid2"].data_ptr() if out.get("resid2") is not None else None)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        check(self._L.ef_model_recognize_boxes_device(
            self._h, frames.data_ptr(), frames.stride(0), frames.stride(1), W, H, channels, F, boxes.data_ptr(), B,
            side, side, float(threshold), C.byref(res), C.c_void_p(stream)), "ef_model_recognize_boxes_device")
        return out


def recognize_boxes_all_models(recognizers, frames, boxes, side, threshold=0.7):
    """The same boxes against several models in ONE C-ABI call (ef_models_recognize_boxes_host): frames / boxes as in
    Recognizer.recognize_boxes.  Returns (score [M, B] float64, index [M, B] int32, label [M, B] int32) as numpy arrays.
    The single-crop pattern of the reference (recognize_face_all_models, scan-template-v4.py:289-319) costs one upload,
    K1 once, one download and one synchronisation instead of a round trip per model."""
    recs = list(recognizers)
    if not recs:
        raise ValueError("no models")
    fr = np.ascontiguousarray(frames)
    if fr.dtype != np.uint8:
        raise ValueError("frames must be uint8")
    if fr.ndim == 2 or (fr.ndim == 3 and fr.shape[2] == 3):
        fr = fr[None]
    channels = 3 if fr.ndim == 4 else 1
    F, H, W = fr.shape[:3]
    bx = np.asarray(boxes, dtype=np.int32).reshape(-1, np.asarray(boxes).shape[-1])
    if bx.shape[1] == 4:
        bx = np.concatenate([np.zeros((len(bx), 1), np.int32), bx], axis=1)
    bx = np.ascontiguousarray(bx)
    B, M = len(bx), len(recs)
    if any(side * side != r.D for r in recs):
        raise ValueError("side*side must equal every model's dimension")
    score = np.empty((M, B), dtype=np.float64)
    index = np.empty((M, B), dtype=np.int32)
    label = np.empty((M, B), dtype=np.int32)
    handles = (C.c_void_p * M)(*[r._h for r in recs])
    L = recs[0]._L
    check(L.ef_models_recognize_boxes_host(handles, M, _ptr(fr), H * W * channels, W * channels, W, H, channels, F,
                                           _ptr(bx), B, side, side, float(threshold), _ptr(score), _ptr(index),
                                           _ptr(label)), "ef_models_recognize_boxes_host")
    return score, index, label


def preprocess_device(frames, boxes, side, out=None, bad=None):
    """K1 alone on torch CUDA tensors: returns uint8 [B, ld] with ld = side*side rounded up to 128.
    bad: optional int32 CUDA tensor [1] that accumulates the number of boxes outside their frame (asynchronous use: the
    caller checks it when it synchronises anyway).  Without it the call checks for itself -- one synchronisation -- and
    raises: a black crop is never returned silently."""
    import torch
    channels = 3 if frames.dim() == 4 else 1
    F, H, W = frames.shape[:3]
    B = boxes.shape[0]
    ld = (side * side + 127) // 128 * 128
    if out is None:
        out = torch.zeros((B, ld), dtype=torch.uint8, device=frames.device)
    counter = bad if bad is not None else torch.zeros(1, dtype=torch.int32, device=frames.device)
    stream = torch.cuda.current_stream(frames.device).cuda_stream
    check(_lib.lib().ef_preprocess(frames.data_ptr(), frames.stride(0), frames.stride(1), W, H, channels, F,
                                   boxes.data_ptr(), B, side, side, out.data_ptr(), out.stride(0), counter.data_ptr(),
                                   C.c_void_p(stream)), "ef_preprocess")
    if bad is None and B:
        n_bad = int(counter.item())
        if n_bad:
            raise _lib.EigenfacesError(_lib.EF_ERR_INVALID, "ef_preprocess", f"{n_bad} of {B} boxes are not inside their frame")
    return out


def _check_u8(X):
    X = np.asarray(X)
    if X.dtype != np.uint8:
        Xi = np.rint(X)
        if not (np.array_equal(Xi, X) and Xi.min() >= 0 and Xi.max() <= 255):
            raise ValueError("the device PCA fit takes 8-bit pixel data (integral values in [0, 255])")
        X = Xi.astype(np.uint8)
    if X.ndim != 2:
        raise ValueError("data matrix must be [N, D]")
    return np.ascontiguousarray(X)


def fit_gen1(X, n_components=None):
    """Device manual_pca (useless/train.py:56-128).  Returns (eigenfaces [D,k] F-order, mean, projected, eigenvalues, info)."""
    X = _check_u8(X)
    N, D = X.shape
    n = N if N < D else D
    k = min(N - 1, D) if n_components is None else int(n_components)
    k = min(k, n)
    eigenfaces = np.empty((D, k), dtype=np.float64, order="F")
    mean = np.empty(D, dtype=np.float64)
    projected = np.empty((N, k), dtype=np.float64)
    eigenvalues = np.empty(k, dtype=np.float64)
    info = FitInfo()
    check(_lib.lib().ef_fit_gen1_host(_ptr(X), X.strides[0], N, D, k, _ptr(eigenfaces), _ptr(mean), _ptr(projected),
                                      _ptr(eigenvalues), C.byref(info)), "ef_fit_gen1_host")
    return eigenfaces, mean, projected, eigenvalues, {"sweeps": info.sweeps, "branch": info.branch,
                                                      "off_norm": info.off_norm, "gpu_ms": info.gpu_ms}


def _fit_gen2_like(entry, X, n_components):
    X = _check_u8(X)
    N, D = X.shape
    k = int(n_components)
    out = {
        "mean_face": np.empty(D), "scaler_mean": np.empty(D), "scaler_var": np.empty(D), "scaler_scale": np.empty(D),
        "pca_mean": np.empty(D), "components": np.empty((k, D)), "explained_variance": np.empty(k),
        "explained_variance_ratio": np.empty(k), "singular_values": np.empty(k), "noise_variance": np.empty(1),
        "features": np.empty((N, k)),
    }
    g = Gen2Fit(**{name: _ptr(arr) for name, arr in out.items()})
    info = FitInfo()
    check(getattr(_lib.lib(), entry)(_ptr(X), X.strides[0], N, D, k, C.byref(g), C.byref(info)), entry)
    out["noise_variance"] = float(out["noise_variance"][0])
    out["info"] = {"sweeps": info.sweeps, "branch": info.branch, "off_norm": info.off_norm, "gpu_ms": info.gpu_ms}
    return out


def fit_gen2(X, n_components):
    """Device StandardScaler + PCA(full) (train-v5.py:349-385).  Returns a dict of float64 arrays + info."""
    return _fit_gen2_like("ef_fit_gen2_host", X, n_components)


def fit_manual(X, n_components):
    """Device ManualStandardScaler + ManualPCA (scripts/manual/train-v2.py:9-72, :189-193).  Same dict as fit_gen2."""
    return _fit_gen2_like("ef_fit_manual_host", X, n_components)


def pca_fit_f64(Z, n_components):
    """PCA(k, full).fit_transform / ManualPCA.fit_transform of a float64 matrix on the device (ef_pca_fit_f64_host)."""
    Z = np.ascontiguousarray(np.asarray(Z, dtype=np.float64))
    if Z.ndim != 2:
        raise ValueError("data matrix must be [N, D]")
    N, D = Z.shape
    k = int(n_components)
    out = {"pca_mean": np.empty(D), "components": np.empty((k, D)), "explained_variance": np.empty(k),
           "explained_variance_ratio": np.empty(k), "singular_values": np.empty(k), "noise_variance": np.empty(1),
           "features": np.empty((N, k))}
    g = Gen2Fit(**{name: _ptr(arr) for name, arr in out.items()})
    info = FitInfo()
    check(_lib.lib().ef_pca_fit_f64_host(_ptr(Z), Z.strides[0] // 8, N, D, k, C.byref(g), C.byref(info)), "ef_pca_fit_f64_host")
    out["noise_variance"] = float(out["noise_variance"][0])
    out["info"] = {"sweeps": info.sweeps, "branch": info.branch, "off_norm": info.off_norm, "gpu_ms": info.gpu_ms}
    return out


def scaler_fit_u8(X, flavour=0):
    """StandardScaler.fit (flavour 0) / ManualStandardScaler.fit (1) on 8-bit crops: (mean, var, scale)."""
    X = _check_u8(X)
    N, D = X.shape
    mean, var, scale = np.empty(D), np.empty(D), np.empty(D)
    check(_lib.lib().ef_scaler_fit_u8_host(_ptr(X), X.strides[0], N, D, int(flavour), _ptr(mean), _ptr(var), _ptr(scale)),
          "ef_scaler_fit_u8_host")
    return mean, var, scale


def standardize_u8(X, mean, scale=None):
    """(X - mean) / scale of 8-bit crops on the device (ef_standardize_u8_device) -> float64 [N, D] numpy."""
    import torch
    X = _check_u8(X)
    N, D = X.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    xd = torch.from_numpy(X).to(dev)
    md = torch.from_numpy(_f64(mean)).to(dev)
    sd = torch.from_numpy(_f64(scale)).to(dev) if scale is not None else None
    z = torch.empty((N, D), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().ef_standardize_u8_device(xd.data_ptr(), xd.stride(0), N, D, md.data_ptr(),
                                              sd.data_ptr() if sd is not None else None, None, z.data_ptr(), D,
                                              C.c_void_p(stream)), "ef_standardize_u8_device")
    return z.cpu().numpy()


def project_f64(Z, components, mean):
    """(Z - mean) . components^T on the device (two ef_dgemm_device calls: Z C^T, then the rank-one mean term)."""
    import torch
    Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
    Cm = _f64(components)
    k, D = Cm.shape
    N = Z.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    zd = torch.from_numpy(np.ascontiguousarray(Z)).to(dev)
    cd = torch.from_numpy(Cm).to(dev)
    md = torch.from_numpy(_f64(mean)).reshape(1, D).to(dev)
    ones = torch.ones((N, 1), dtype=torch.float64, device=dev)
    bias = torch.empty((1, k), dtype=torch.float64, device=dev)
    out = torch.empty((N, k), dtype=torch.float64, device=dev)
    L = _lib.lib()
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    check(L.ef_dgemm_device(1, k, D, 1.0, md.data_ptr(), D, 1, cd.data_ptr(), 1, D, 0.0, bias.data_ptr(), k, stream), "dgemm")
    check(L.ef_dgemm_device(N, k, D, 1.0, zd.data_ptr(), D, 1, cd.data_ptr(), 1, D, 0.0, out.data_ptr(), k, stream), "dgemm")
    check(L.ef_dgemm_device(N, k, 1, -1.0, ones.data_ptr(), 1, 1, bias.data_ptr(), k, 1, 1.0, out.data_ptr(), k, stream), "dgemm")
    return out.cpu().numpy()
