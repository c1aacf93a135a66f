"""Micro-benchmark of ef_eigh_jacobi_device (cluster-resident vs global-memory kernel).  Not a bench line."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

L = ef._lib.lib()
for n in [int(a) for a in sys.argv[1:]] or [178, 229, 308, 512]:
    rng = np.random.default_rng(n)
    Z = rng.normal(size=(n, n - 1))
    A = Z @ Z.T
    for mode in ("cluster", "global"):
        if mode == "global":
            os.environ["EF_NO_CLUSTER_JACOBI"] = "1"
        else:
            os.environ.pop("EF_NO_CLUSTER_JACOBI", None)
        evals = torch.empty(n, dtype=torch.float64, device="cuda")
        evecs = torch.empty((n, n), dtype=torch.float64, device="cuda")
        work = torch.empty(L.ef_eigh_work_bytes(n), dtype=torch.uint8, device="cuda")
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ts = []
        for rep in range(3):
            a = torch.from_numpy(A.copy()).cuda()
            sweeps, off = C.c_int32(), C.c_double()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ef._lib.check(L.ef_eigh_jacobi_device(a.data_ptr(), n, evals.data_ptr(), evecs.data_ptr(), work.data_ptr(), 0,
                                                  0.0, C.byref(sweeps), C.byref(off), st), "jacobi")
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        w = evals.cpu().numpy(); V = evecs.cpu().numpy()
        w_ref = np.linalg.eigvalsh(A)[::-1]
        print(f"n={n} {mode}: {min(ts):8.3f} ms  sweeps {sweeps.value}  max|dw|/w0 {np.abs(w - w_ref).max() / w_ref[0]:.2e}  "
              f"|VV^T-I| {np.abs(V @ V.T - np.eye(n)).max():.2e}", flush=True)
