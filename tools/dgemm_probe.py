"""Micro-benchmark of ef_dgemm_device at the subspace-solver shape (10 000 x 10 000 by 10 000 x 288).  Not a bench line."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

L = ef._lib.lib()
for (M, N, K) in ((10000, 288, 10000), (288, 288, 10000), (10000, 288, 288), (12500, 256, 10000)):
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((K, N), dtype=torch.float64, device="cuda")
    Cm = torch.empty((M, N), dtype=torch.float64, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        ef._lib.check(L.ef_dgemm_device(M, N, K, 1.0, A.data_ptr(), K, 1, B.data_ptr(), N, 1, 0.0, Cm.data_ptr(), N, st), "dgemm")
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    err = float((Cm - A @ B).abs().max())
    print(f"M={M} N={N} K={K}: {ms:.3f} ms  {2.0 * M * N * K / ms / 1e9:.2f} TFLOP/s  max|diff vs torch| {err:.2e}", flush=True)
