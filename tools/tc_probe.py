"""Debugging aid: run C2-shaped batches through the tensor-core projection with per-CTA phase timestamps
(EF_TC_PROBE=1) and a few pipeline-shape overrides (EF_TC_STAGES, EF_TC_GRID).  Not a benchmark."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B, D, K = 4096, 10000, 10
rng = np.random.default_rng(0)
E = np.linalg.qr(rng.normal(size=(D, K)))[0]
mu = rng.uniform(60, 200, D)
G = rng.normal(size=(1024, K)) * 1000
rec = ef.Recognizer(E, mu, G, metric=ef.METRIC_COSINE_G1, with_residual=True)
xs = [torch.randint(0, 256, (B, 10112), dtype=torch.uint8, device="cuda") for _ in range(6)]
out = rec.recognize_device(xs[0], 0.8)
torch.cuda.synchronize()


def timed(tag, n=40):
    for i in range(5):
        rec.recognize_device(xs[i % 6], 0.8, out=out)
    torch.cuda.synchronize()
    rec.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        rec.recognize_device(xs[i % 6], 0.8, out=out)
    e1.record()
    torch.cuda.synchronize()
    calls, ms, tc = rec.kernel_timing_read()
    rec.kernel_timing(False)
    print(f"{tag}: step {1e3 * e0.elapsed_time(e1) / n:.1f} us, projection kernel {1e3 * ms:.1f} us (tc={tc})", flush=True)


for mode in (2, 1):
    rec.use_tensor_cores(mode)
    timed(f"mode {mode}")
rec.use_tensor_cores(1)
for env in ({}, {"EF_TC_STAGES": "4"}, {"EF_TC_STAGES": "2"}, {"EF_TC_GRID": "128"}, {"EF_TC_GRID": "74"},
            {"EF_TC_GRID": "296"}):
    for k in ("EF_TC_STAGES", "EF_TC_GRID", "EF_TC_PROBE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    timed(str(env))
    os.environ["EF_TC_PROBE"] = "1"
    rec.recognize_device(xs[1], 0.8, out=out)
    torch.cuda.synchronize()
    os.environ.pop("EF_TC_PROBE")
# without residual (no sum-of-squares pass in the epilogue warps)
for k in ("EF_TC_STAGES", "EF_TC_GRID"):
    os.environ.pop(k, None)
rec2 = ef.Recognizer(E, mu, G, metric=ef.METRIC_COSINE_G1, with_residual=False)
o2 = rec2.recognize_device(xs[0], 0.8)
rec = rec2
out = o2
timed("no residual")
os.environ["EF_TC_PROBE"] = "1"
rec.recognize_device(xs[1], 0.8, out=out)
torch.cuda.synchronize()
