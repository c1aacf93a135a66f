"""Debugging aid: phase timestamps of the stream-K projection kernel on the shipped Gen-1 shape (k = 50)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B = 4096
rng = np.random.default_rng(0)
for D, k, ng, scaled in ((10000, 50, 229, False), (4096, 178, 178, True)):
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    kw = dict(scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D)) if scaled else {}
    rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=ef.METRIC_COSINE_G1, **kw)
    ld = (D + 127) // 128 * 128
    xs = [torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda") for _ in range(6)]
    out = rec.recognize_device(xs[0], 0.8)
    for env in ({}, {"EF_TC_STAGES": "2"}, {"EF_TC_GRID": "128"}, {"EF_TC_GRID": "64"}, {"EF_TC_GRID": "32"}):
        for kk in ("EF_TC_STAGES", "EF_TC_GRID", "EF_TC_PROBE"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        for i in range(5):
            rec.recognize_device(xs[i % 6], 0.8, out=out)
        torch.cuda.synchronize()
        rec.kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(30):
            rec.recognize_device(xs[i % 6], 0.8, out=out)
        e1.record()
        torch.cuda.synchronize()
        calls, ms, tc = rec.kernel_timing_read()
        rec.kernel_timing(False)
        print(f"D={D} k={k} {env}: step {1e3 * e0.elapsed_time(e1) / 30:.1f} us, projection kernel {1e3 * ms:.1f} us", flush=True)
        os.environ["EF_TC_PROBE"] = "1"
        rec.recognize_device(xs[1], 0.8, out=out)
        torch.cuda.synchronize()
        os.environ.pop("EF_TC_PROBE")
    rec.close()
