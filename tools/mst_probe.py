"""Debugging aid: phase clock stamps of match_small_tc_kernel (EF_MST_TRACE) on the shipped model shapes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(0)
B = 4096
for name, D, k, ng, metric, scaled in (("k=50 n=229", 10000, 50, 229, ef.METRIC_COSINE_G1, False),
                                       ("k=178 n=178", 4096, 178, 178, ef.METRIC_COSINE_SK, True),
                                       ("k=50 n=590", 4096, 50, 590, ef.METRIC_COSINE_SK, True),
                                       ("k=590 n=590", 4096, 590, 590, ef.METRIC_COSINE_SK, True)):
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    kw = dict(scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D)) if scaled else {}
    rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(ng, k)) * 100, metric=metric, **kw)
    ld = (D + 127) // 128 * 128
    x = torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda")
    out = rec.recognize_device(x, 0.8)
    torch.cuda.synchronize()
    print(name, flush=True)
    os.environ["EF_MST_TRACE"] = "1"
    for _ in range(3):
        rec.recognize_device(x, 0.8, out=out)
    torch.cuda.synchronize()
    os.environ.pop("EF_MST_TRACE")
    rec.close()
