import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import eigenfaces_b200 as ef
N, D, R, K = 25000, 10000, 300, 256
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(4242)
F = torch.linalg.qr(torch.randn((D, R), generator=g, device=dev, dtype=torch.float32))[0]
sig = 40.0 * torch.arange(1, R + 1, device=dev, dtype=torch.float32) ** -0.7
X = torch.empty((N, D), dtype=torch.uint8, device=dev)
for i in range(0, N, 5000):
    L = torch.randn((5000, R), generator=g, device=dev) * sig
    X[i:i + 5000] = (128 + L @ F.T + 4.0 * torch.randn((5000, D), generator=g, device=dev)).round_().clamp_(0, 255).to(torch.uint8)
Xf = X.double(); mean = Xf.mean(0); Xc = Xf - mean
cov = (Xc.T @ Xc) / (N - 1)
del Xf, Xc
for deg in (8, 12, 16, 24):
    for blk in (288, 320):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lam, Q, info = ef.dist.eigh_topk_device(cov, K, degree=deg, block=blk)
        torch.cuda.synchronize()
        print(f"degree {deg} block {blk}: {time.perf_counter() - t0:.3f} s", info, flush=True)
