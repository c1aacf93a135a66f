"""The L2 metric (north-star extra X1) on the bench shape: time per 4096-crop batch through recognize_device and through
the submit / flush queue, beside the Gen-1 cosine metric.  Not a bench line."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B, D, k, ng = 4096, 10000, 10, 1024
rng = np.random.default_rng(0)
E = np.linalg.qr(rng.normal(size=(D, k)))[0]
mean = rng.uniform(60, 200, D)
gal = rng.normal(size=(ng, k)) * 100
ld = (D + 127) // 128 * 128
xs = [torch.randint(0, 256, (B, ld), dtype=torch.uint8, device="cuda") for _ in range(8)]
for name, metric in (("cosine (Gen-1)", ef.METRIC_COSINE_G1), ("L2", ef.METRIC_L2)):
    rec = ef.Recognizer(E, mean, gal, metric=metric, labels=np.arange(ng) % 4)
    thr = 0.8 if metric != ef.METRIC_L2 else 1e12
    outs = [rec.recognize_device(x, thr) for x in xs]
    torch.cuda.synchronize()
    for label, fn in (("recognize_device", lambda i: rec.recognize_device(xs[i % 8], thr, out=outs[i % 8])),
                      ("submit_device queue", lambda i: rec.submit_device(xs[i % 8], thr, out=outs[i % 8]))):
        for i in range(40):
            fn(i)
        rec.flush_device(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            fn(i)
        rec.flush_device()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        print(f"{name:16s} {label:22s}: {us:7.2f} us per batch, path {rec.serving_path()}", flush=True)
    rec.close()
