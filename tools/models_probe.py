"""Device-resident recognition throughput for the shipped model shapes (paths chosen by ef_model_recognize_device).
Not a bench line."""
import os
import sys

import numpy as np
import torch

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
    for i in range(5):
        rec.recognize_device(xs[i % 8], 0.8, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ef.launch_count()
    e0.record()
    for i in range(50):
        rec.recognize_device(xs[i % 8], 0.8, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    n_call, _, path = (0, 0, 0)
    print(f"{name}: {ms * 1e3:8.1f} us per 4096 crops = {B / ms / 1e3:7.1f} M crops/s, {(ef.launch_count() - l0) / 50:.0f} launches per call, "
          f"{B * D / ms / 1e6:7.1f} GB/s of crops", flush=True)
    rec.close()
