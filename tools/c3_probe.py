"""Config 3 on one GPU: n gallery identities x k = 128, B = 4096 queries -- tensor-core matcher vs the float64 scan on a
query subsample.  Not a bench line."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k, B = 128, 4096
dev = torch.device("cuda")
gen = torch.Generator(device=dev); gen.manual_seed(1_000_003)
lam = (torch.arange(1, k + 1, device=dev, dtype=torch.float64)) ** -2.0
G = torch.randn((n, k), generator=gen, device=dev, dtype=torch.float64) * lam.sqrt()
truth = torch.randint(0, n, (B,), generator=gen, device=dev)
P = G[truth] + 0.05 * torch.randn((B, k), generator=gen, device=dev, dtype=torch.float64) * lam.sqrt()
metrics = {"sk": ef.METRIC_COSINE_SK, "g1": ef.METRIC_COSINE_G1, "l2": ef.METRIC_L2}
for metric in [metrics[m] for m in os.environ.get("EF_C3_METRICS", "sk").split(",")]:
    t0 = time.perf_counter()
    sg = ef.dist.ShardedGallery(G, 0, metric)
    torch.cuda.synchronize()
    print(f"gallery prepare + float16 image: {time.perf_counter() - t0:.3f} s, image {sg.image.numel() / 1e6:.0f} MB", flush=True)
    for _ in range(2):
        sg.match_local(P)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        s, i = sg.match_local(P)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * B * n * 384 * (2 if os.environ.get("EF_MATCH_TC_TWO_PASS") else 1)   # K = 3k float16 products per pass
    print(f"metric {metric} n={n} k={k} B={B}: {ms:.2f} ms per batch = {B / ms * 1e3 / 1e6:.2f} M queries/s; filter {flops / ms / 1e9:.0f} TFLOP/s (f16); "
          f"flags {sg.last_flags}; accuracy vs planted {float((i == truth).double().mean()):.4f}", flush=True)
    # float64 scan on a subsample for the timing comparison and an equality check
    sub = 64
    sg64 = ef.dist.ShardedGallery(G, 0, metric, use_tensor_cores=False)
    sg64.match_local(P[:sub]); torch.cuda.synchronize()
    e0.record(); s64, i64 = sg64.match_local(P[:sub]); e1.record(); torch.cuda.synchronize()
    print(f"float64 scan: {e0.elapsed_time(e1):.1f} ms for {sub} queries -> {e0.elapsed_time(e1) * B / sub:.0f} ms per 4096 (extrapolated); "
          f"equal: {bool(torch.equal(i64, i[:sub]) and torch.equal(s64, s[:sub]))}", flush=True)
