"""ctypes binding of include/eigenfaces_b200.h.  No CPU fallback: a missing library or device raises."""
import ctypes as C
import os

from . import _build

c_i32, c_i64, c_dbl, c_void = C.c_int32, C.c_int64, C.c_double, C.c_void_p
p_u8, p_dbl, p_i32, p_i64 = C.POINTER(C.c_uint8), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)

EF_OK, EF_ERR_INVALID, EF_ERR_UNSUPPORTED, EF_ERR_CUDA, EF_ERR_NOMEM, EF_ERR_NOCONVERGE = 0, -1, -2, -3, -4, -5
METRIC_COSINE_SK, METRIC_COSINE_G1, METRIC_L2 = 0, 1, 2


class EigenfacesError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: {_LIB.ef_error_string(status).decode() if _LIB else status}"
                         + (f" [{detail}]" if detail else ""))


class Box(C.Structure):
    _fields_ = [("frame", c_i32), ("x", c_i32), ("y", c_i32), ("w", c_i32), ("h", c_i32)]


class ModelDesc(C.Structure):
    _fields_ = [("D", c_i32), ("k", c_i32), ("basis", c_void), ("basis_stride_d", c_i64), ("basis_stride_k", c_i64),
                ("mean", c_void), ("scale", c_void), ("pca_mean", c_void), ("gallery", c_void), ("gallery_ld", c_i64),
                ("n_gallery", c_i32), ("labels", c_void), ("metric", c_i32), ("n_slices", c_i32),
                ("with_residual", c_i32)]


class Result(C.Structure):
    _fields_ = [("proj", c_void), ("score", c_void), ("index", c_void), ("label", c_void), ("resid2", c_void)]


class FitInfo(C.Structure):
    _fields_ = [("sweeps", c_i32), ("branch", c_i32), ("off_norm", c_dbl), ("gpu_ms", c_dbl)]


class Gen2Fit(C.Structure):
    _fields_ = [(n, c_void) for n in ("mean_face", "scaler_mean", "scaler_var", "scaler_scale", "pca_mean",
                                      "components", "explained_variance", "explained_variance_ratio",
                                      "singular_values", "noise_variance", "features")]


# name -> (restype, argtypes); this table is also what tests/test_abi.py checks against the header.
SIGNATURES = {
    "ef_version": (C.c_int, []),
    "ef_error_string": (C.c_char_p, [C.c_int]),
    "ef_last_error_detail": (C.c_char_p, []),
    "ef_launch_count": (c_i64, []),
    "ef_device_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ef_preprocess": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_void, c_i32, c_i32, c_i32, c_void,
                                c_i64, c_void, c_void]),
    "ef_model_create": (C.c_int, [C.POINTER(c_void), C.POINTER(ModelDesc)]),
    "ef_model_destroy": (None, [c_void]),
    "ef_model_reserve": (C.c_int, [c_void, c_i32]),
    "ef_model_dims": (C.c_int, [c_void, p_i32, p_i32, p_i32, p_i32]),
    "ef_model_set_tensor_cores": (C.c_int, [c_void, c_i32]),
    "ef_model_status": (C.c_int, [c_void, p_i32]),
    "ef_model_kernel_timing": (C.c_int, [c_void, c_i32]),
    "ef_model_kernel_timing_read": (C.c_int, [c_void, p_i32, p_dbl, p_i32]),
    "ef_model_recognize_device": (C.c_int, [c_void, c_void, c_i64, c_i32, c_dbl, C.POINTER(Result), c_void]),
    "ef_model_submit_device": (C.c_int, [c_void, c_void, c_i64, c_i32, c_dbl, C.POINTER(Result), c_void]),
    "ef_model_flush_device": (C.c_int, [c_void, c_void]),
    "ef_model_set_serving": (C.c_int, [c_void, c_i32, c_i32]),
    "ef_model_recognize_host": (C.c_int, [c_void, c_void, c_i64, c_i32, c_dbl, C.POINTER(Result)]),
    "ef_model_submit_host": (C.c_int, [c_void, c_void, c_i64, c_i32, c_dbl, c_i32, p_i32]),
    "ef_model_wait_host": (C.c_int, [c_void, c_i32, C.POINTER(Result)]),
    "ef_model_recognize_boxes_device": (C.c_int, [c_void, c_void, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_void,
                                                  c_i32, c_i32, c_i32, c_dbl, C.POINTER(Result), c_void]),
    "ef_model_recognize_boxes_host": (C.c_int, [c_void, c_void, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_void,
                                                c_i32, c_i32, c_i32, c_dbl, C.POINTER(Result)]),
    "ef_models_recognize_boxes_host": (C.c_int, [c_void, c_i32, c_void, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_void,
                                                 c_i32, c_i32, c_i32, c_dbl, c_void, c_void, c_void]),
    "ef_model_bad_boxes": (C.c_int, [c_void, c_void, p_i32]),
    "ef_match_work_bytes": (C.c_size_t, [c_i32, c_i64]),
    "ef_gallery_prepare_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_i32, c_void, c_i64, c_void, c_void]),
    "ef_match_device": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_void, c_i64, c_void, c_i64, c_i64, c_i32, c_void,
                                  c_void, c_void, c_void]),
    "ef_match_tc_image_bytes": (C.c_size_t, [c_i64, c_i32]),
    "ef_match_tc_image_bytes_metric": (C.c_size_t, [c_i64, c_i32, c_i32]),
    "ef_match_tc_prepare_device": (C.c_int, [c_void, c_i64, c_void, c_i64, c_i32, c_i32, c_void, c_void]),
    "ef_match_tc_work_bytes": (C.c_size_t, [c_i32, c_i64, c_i32]),
    "ef_match_tc_device": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_void, c_i64, c_void, c_void, c_i64, c_i64, c_i32,
                                     c_void, c_void, c_void, C.c_size_t, c_void]),
    "ef_match_tc_flags": (C.c_int, [c_void, p_i32]),
    "ef_template_match_work_bytes": (C.c_size_t, [c_i32, c_i32, c_i32, c_void, c_void]),
    "ef_template_match_device": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_void, c_void, c_void, c_void, c_i32, c_void,
                                           c_void, c_void, c_void, c_void, C.c_size_t, c_void]),
    "ef_match_reduce_device": (C.c_int, [c_void, c_void, c_i32, c_i32, c_i32, c_void, c_void, c_void]),
    "ef_fit_gen1_host": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, c_void, c_void, c_void, c_void,
                                   C.POINTER(FitInfo)]),
    "ef_fit_gen2_host": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, C.POINTER(Gen2Fit), C.POINTER(FitInfo)]),
    "ef_fit_manual_host": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, C.POINTER(Gen2Fit), C.POINTER(FitInfo)]),
    "ef_pca_fit_f64_host": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, C.POINTER(Gen2Fit), C.POINTER(FitInfo)]),
    "ef_scaler_fit_u8_host": (C.c_int, [c_void, c_i64, c_i32, c_i32, c_i32, c_void, c_void, c_void]),
    "ef_colsum_u8_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_void, c_void]),
    "ef_gram_u8_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_void, c_void]),
    "ef_gram_u8_tc_work_bytes": (C.c_size_t, [c_i64, c_i32, c_i32]),
    "ef_gram_u8_tc_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_void, c_void, C.c_size_t,
                                       c_void]),
    "ef_gram_u8_tc_store_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_i32, c_i32, c_i32, c_void, c_void, C.c_size_t,
                                             c_void]),
    "ef_gram_center_work_bytes": (C.c_size_t, [c_i32]),
    "ef_gram_center_device": (C.c_int, [c_void, c_i32, c_i32, c_void, c_i64, c_dbl, c_void, c_void, c_void]),
    "ef_eigh_work_bytes": (C.c_size_t, [c_i32]),
    "ef_eigh_jacobi_device": (C.c_int, [c_void, c_i32, c_void, c_void, c_void, c_i32, c_dbl, p_i32, p_dbl, c_void]),
    "ef_chol_inverse_device": (C.c_int, [c_void, c_i32, c_void, c_void, c_void]),
    "ef_dgemm_device": (C.c_int, [c_i32, c_i32, c_i32, c_dbl, c_void, c_i64, c_i64, c_void, c_i64, c_i64, c_dbl, c_void,
                                  c_i64, c_void]),
    "ef_dgemm_tc_work_bytes": (C.c_size_t, [c_i32, c_i32, c_i32]),
    "ef_dgemm_tc_device": (C.c_int, [c_i32, c_i32, c_i32, c_dbl, c_void, c_i64, c_i64, c_void, c_i64, c_i64, c_dbl, c_void,
                                     c_i64, c_i32, c_void, c_void]),
    "ef_standardize_u8_device": (C.c_int, [c_void, c_i64, c_i64, c_i32, c_void, c_void, c_void, c_void, c_i64, c_void]),
}

_LIB = None


def lib():
    """The loaded library (built on first use when the .so is missing and nvcc is available)."""
    global _LIB
    if _LIB is None:
        # content-hash check against the build manifest: a library older than csrc/ or the header is rebuilt (the
        # build is incremental and serialised across processes by a file lock), never loaded stale
        path = _build.build()
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError here == header / library drift
            fn.restype = res
            fn.argtypes = args
        if handle.ef_version() != 1:
            raise RuntimeError(f"libeigenfaces_b200.so ABI {handle.ef_version()} != 1")
        _LIB = handle
    return _LIB


def check(status, where):
    if status != EF_OK:
        with_detail = status == EF_ERR_CUDA or (status == EF_ERR_INVALID and "boxes" in where)
        detail = lib().ef_last_error_detail().decode() if with_detail else ""
        raise EigenfacesError(status, where, detail)


def launch_count():
    return int(lib().ef_launch_count())
