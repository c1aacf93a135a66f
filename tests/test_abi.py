"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
binds with the documented signatures, and fails loudly (never falls back) without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import eigenfaces_b200 as ef
from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "eigenfaces_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ef_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 30
    lib = C.CDLL(ef._lib._build.build())                  # incremental: rebuilds only what changed
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_binding_table_matches_header():
    assert sorted(ef._lib.SIGNATURES) == _declared()
    L = ef._lib.lib()
    assert L.ef_version() == 1
    assert L.ef_error_string(0) == b"ok" and b"CUDA" in L.ef_error_string(-3)


def test_struct_layouts_match_the_header():
    assert C.sizeof(ef._lib.Box) == 20
    assert C.sizeof(ef._lib.Result) == 5 * C.sizeof(C.c_void_p)
    assert C.sizeof(ef._lib.ModelDesc) == 104
    assert C.sizeof(ef._lib.FitInfo) == 24
    assert C.sizeof(ef._lib.Gen2Fit) == 11 * C.sizeof(C.c_void_p)


def test_argument_validation_needs_no_gpu():
    L = ef._lib.lib()
    assert L.ef_model_create(None, None) == ef._lib.EF_ERR_INVALID
    assert L.ef_preprocess(None, 0, 0, 0, 0, 1, 1, None, 1, 64, 64, None, 0, None, None) == ef._lib.EF_ERR_INVALID
    assert L.ef_eigh_work_bytes(100) >= 8 * (100 * 100 + 100)
    assert L.ef_eigh_work_bytes(0) == 0


def test_matcher_image_sizes_are_host_arithmetic():
    """Sizing functions of the tensor-core matchers are pure host arithmetic (callable without a device): 256-row tiles x
    K slabs of 64 halfs (3 k halfs per row, 3 (k + 1) for the Euclidean metric) x 32 KB + a 256-byte trailer."""
    L = ef._lib.lib()
    tiles = -(-1_000_000 // 256)
    assert L.ef_match_tc_image_bytes(1_000_000, 128) == tiles * 6 * 32768 + 256
    assert L.ef_match_tc_image_bytes_metric(1_000_000, 128, ef.METRIC_COSINE_G1) == tiles * 6 * 32768 + 256
    assert L.ef_match_tc_image_bytes_metric(1_000_000, 128, ef.METRIC_L2) == tiles * 7 * 32768 + 256
    assert L.ef_match_tc_image_bytes_metric(229, 50, ef.METRIC_COSINE_SK) == 1 * 3 * 32768 + 256
    assert L.ef_match_tc_image_bytes_metric(590, 590, ef.METRIC_COSINE_SK) == 3 * 28 * 32768 + 256
    assert L.ef_match_tc_image_bytes_metric(0, 50, ef.METRIC_COSINE_SK) == 0
    assert L.ef_match_tc_prepare_device(None, 0, None, 0, 0, 0, None, None) == ef._lib.EF_ERR_INVALID
    assert L.ef_match_tc_device(None, 0, 0, 0, None, 0, None, None, 0, 0, 0, None, None, None, 0, None) == ef._lib.EF_ERR_INVALID


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(ef.EigenfacesError) as ei:
        ef.Recognizer(np.eye(16, 4), np.zeros(16), np.zeros((3, 4)))
    assert ei.value.status == ef._lib.EF_ERR_CUDA
    with pytest.raises(ef.EigenfacesError):
        ef.fit_gen1(np.zeros((4, 16), np.uint8), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "face-detection-recognization-pca_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|import_module\(.oracle|#include\s+.*oracle/|dlopen\(.*oracle",
                             src, flags=re.M):
                    offenders.append(f)
    assert not offenders
