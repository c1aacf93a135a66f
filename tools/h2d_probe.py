"""Debugging aid: raw pinned H2D bandwidth for one C2 batch vs the host-buffer recognise call.  Not a bench line."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

B, D = 4096, 10000
h = torch.randint(0, 256, (B, D), dtype=torch.uint8).pin_memory()
d = torch.empty((B, D), dtype=torch.uint8, device="cuda")
for chunks in (1, 2, 4, 8):
    rows = B // chunks
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        for c in range(chunks):
            d[c * rows:(c + 1) * rows].copy_(h[c * rows:(c + 1) * rows], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(f"H2D {B * D / 1e6:.1f} MB in {chunks} chunk(s): {dt * 1e3:.3f} ms = {B * D / dt / 1e9:.1f} GB/s", flush=True)
rng = np.random.default_rng(0)
E = np.linalg.qr(rng.normal(size=(D, 10)))[0]
rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(1024, 10)) * 1000, metric=ef.METRIC_COSINE_G1)
hn = h.numpy()
for _ in range(5):
    rec.recognize(hn, 0.8, want_features=False)
t0 = time.perf_counter()
for _ in range(50):
    rec.recognize(hn, 0.8, want_features=False)
dt = (time.perf_counter() - t0) / 50
print(f"recognize (host buffers): {dt * 1e3:.3f} ms per batch = {B / dt / 1e6:.2f} M crops/s", flush=True)
