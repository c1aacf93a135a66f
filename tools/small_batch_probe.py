"""Small batches through the shipped shapes: tensor-core matcher (default) against the float64 kernels
(EF_NO_MATCH_SMALL_TC=1), host call and device-resident call.  Not a bench line."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(0)
for name, D, k, ng, metric, scaled in (("k=50 n=229", 10000, 50, 229, ef.METRIC_COSINE_G1, False),
                                       ("k=178 n=178", 4096, 178, 178, ef.METRIC_COSINE_SK, True),
                                       ("k=50 n=590", 4096, 50, 590, ef.METRIC_COSINE_SK, True)):
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    kw = dict(scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D)) if scaled else {}
    rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=metric, **kw)
    ld = (D + 127) // 128 * 128
    for B in (1, 8, 64, 512):
        xh = rng.integers(0, 256, (B, D), dtype=np.uint8)
        xd = torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda")
        line = f"{name} B={B:4d}:"
        for env in ({}, {"EF_NO_MATCH_SMALL_TC": "1"}):
            os.environ.update(env)
            out = rec.recognize_device(xd, 0.8)
            for _ in range(10):
                rec.recognize(xh, 0.8)
                rec.recognize_device(xd, 0.8, out=out)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                rec.recognize(xh, 0.8)
            host_us = (time.perf_counter() - t0) / 200 * 1e6
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                rec.recognize_device(xd, 0.8, out=out)
            e1.record(); torch.cuda.synchronize()
            dev_us = e0.elapsed_time(e1) / 200 * 1e3
            line += f"  {'float64 kernels' if env else 'tensor-core matcher'}: host call {host_us:7.1f} us, device call {dev_us:6.1f} us;"
            for kname in env:
                os.environ.pop(kname)
        print(line, flush=True)
    rec.close()
