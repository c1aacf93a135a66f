"""Debugging aid: phase timestamps of recognize_cluster_kernel on a C2-shaped batch (EF_TC_PROBE=1).  Not a benchmark."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B, D, K = 4096, 10000, 10
rng = np.random.default_rng(0)
E = np.linalg.qr(rng.normal(size=(D, K)))[0]
mu = rng.uniform(60, 200, D)
G = rng.normal(size=(1024, K)) * 1000
xs = [torch.randint(0, 256, (B, 10112), dtype=torch.uint8, device="cuda") for _ in range(6)]
for resid in (True, False):
    rec = ef.Recognizer(E, mu, G, metric=ef.METRIC_COSINE_G1, with_residual=resid)
    out = rec.recognize_device(xs[0], 0.8)
    torch.cuda.synchronize()
    for i in range(5):
        rec.recognize_device(xs[i % 6], 0.8, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        rec.recognize_device(xs[i % 6], 0.8, out=out)
    e1.record()
    torch.cuda.synchronize()
    print(f"with_residual={resid}: step {1e3 * e0.elapsed_time(e1) / 50:.1f} us", flush=True)
    for mode in os.environ.get("PROBE_MODES", "1").split(","):
        os.environ["EF_TC_PROBE"] = mode
        print("probe mode", mode, flush=True)
        for i in range(2):
            rec.recognize_device(xs[i + 1], 0.8, out=out)
        torch.cuda.synchronize()
        for i in range(3):
            rec.submit_device(xs[i + 1], 0.8, out=out)
        rec.flush_device()
        torch.cuda.synchronize()
        os.environ.pop("EF_TC_PROBE")
    for i in range(5):
        rec.submit_device(xs[i % 6], 0.8, out=out)
    e0.record()
    for i in range(50):
        rec.submit_device(xs[i % 6], 0.8, out=out)
    rec.flush_device()
    e1.record()
    torch.cuda.synchronize()
    print(f"with_residual={resid}: pipelined step {1e3 * e0.elapsed_time(e1) / 51:.1f} us", flush=True)
