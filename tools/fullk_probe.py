import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import eigenfaces_b200 as ef
rng = np.random.default_rng(0)
B = 4096
for k in (272, 308, 590):
    D = 4096
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    rec = ef.Recognizer(E, rng.uniform(60, 200, D), rng.normal(size=(k, k)) * 100, metric=ef.METRIC_COSINE_SK,
                        scale=rng.uniform(20, 60, D), pca_mean=rng.normal(0, 1e-3, D))
    xs = [torch.randint(0, 256, (B, D), dtype=torch.uint8, device="cuda") for _ in range(4)]
    out = rec.recognize_device(xs[0], 0.8)
    for i in range(5):
        rec.recognize_device(xs[i % 4], 0.8, out=out)
    torch.cuda.synchronize()
    rec.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        rec.recognize_device(xs[i % 4], 0.8, out=out)
    e1.record(); torch.cuda.synchronize()
    calls, ms, tc = rec.kernel_timing_read()
    rec.kernel_timing(False)
    e0.record()
    for i in range(30):
        rec.recognize_device(xs[i % 4], 0.8, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"k=N={k}: {e0.elapsed_time(e1) / 30 * 1e3:.1f} us per 4096 crops; projection kernel {ms * 1e3:.1f} us (NC = {8 * (k + 1)})", flush=True)
    rec.close()
