"""Micro-benchmark of ef_dgemm_tc_device (FP64 tensor-core path) beside ef_dgemm_device (CUDA cores) and torch.matmul
(cuBLAS) at the subspace-solver shapes.  Not a bench line."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

L = ef._lib.lib()
work = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
shapes = ((8192, 8192, 8192, 1), (10000, 320, 10000, 1), (10000, 256, 10000, 1), (10000, 192, 10000, 1), (10000, 128, 10000, 1),
          (320, 320, 10000, 16), (10000, 320, 320, 1), (100000, 256, 10000, 1))
for (M, N, K, splits) in shapes:
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((K, N), dtype=torch.float64, device="cuda")
    Cm = torch.empty((M, N), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run_tc():
        ef._lib.check(L.ef_dgemm_tc_device(M, N, K, 1.0, A.data_ptr(), K, 1, B.data_ptr(), N, 1, 0.0, Cm.data_ptr(), N, splits,
                                           work.data_ptr(), st), "dgemm_tc")

    def run_cc():
        ef._lib.check(L.ef_dgemm_device(M, N, K, 1.0, A.data_ptr(), K, 1, B.data_ptr(), N, 1, 0.0, Cm.data_ptr(), N, st), "dgemm")

    def run_blas():
        torch.matmul(A, B, out=Cm)

    line = f"M={M} N={N} K={K} splits={splits}:"
    for name, fn in (("tensor-core", run_tc), ("cuda-core", run_cc), ("cuBLAS", run_blas)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        line += f"  {name} {ms:.3f} ms = {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s"
    print(line, flush=True)
