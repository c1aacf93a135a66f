"""Warm per-kernel durations (CUPTI through torch.profiler) of ef_model_recognize_device for the shipped model shapes.
ncu's launch list serialises kernels and flushes caches; this one times them inside the steady-state loop."""
import os
import sys
from collections import defaultdict

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(0)
B = 4096
only = sys.argv[1] if len(sys.argv) > 1 else ""
for name, D, k, ng, metric, scaled in (("Gen-1 shipped (100x100, k=50, 229 rows)", 10000, 50, 229, ef.METRIC_COSINE_G1, False),
                                       ("Gen-2 train-v5 (64x64, k=178, 178 rows)", 4096, 178, 178, ef.METRIC_COSINE_SK, True),
                                       ("Gen-2 train-v4 (64x64, k=50, 590 rows)", 4096, 50, 590, ef.METRIC_COSINE_SK, True),
                                       ("Gen-2 full-k (64x64, k=590, 590 rows)", 4096, 590, 590, ef.METRIC_COSINE_SK, True),
                                       ("bench C2 (100x100, k=10, 1024 rows)", 10000, 10, 1024, ef.METRIC_COSINE_G1, False)):
    if only and only not in name:
        continue
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    kw = dict(scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D)) if scaled else {}
    rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=metric, **kw)
    ld = (D + 127) // 128 * 128
    xs = [torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda") for _ in range(8)]
    out = rec.recognize_device(xs[0], 0.8)
    for i in range(10):
        rec.recognize_device(xs[i % 8], 0.8, out=out)
    torch.cuda.synchronize()
    n = 40
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(n):
            rec.recognize_device(xs[i % 8], 0.8, out=out)
        torch.cuda.synchronize()
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for ev in prof.events():
        if ev.device_type.name == "CUDA" or "cuda" in str(ev.device_type).lower():
            tot[ev.name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            cnt[ev.name] += 1
    print(name)
    s = 0.0
    for kname, t in sorted(tot.items(), key=lambda kv: -kv[1]):
        short = kname.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        print(f"    {short[:60]:60s} {t / n:8.2f} us per call ({cnt[kname] / n:.1f} launches)")
        s += t / n
    print(f"    {'sum of kernels':60s} {s:8.2f} us", flush=True)
    rec.close()
