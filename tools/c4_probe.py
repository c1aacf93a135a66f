"""Config 4 on one GPU (all rows local): N x 10 000 uint8 with a planted rank-300 structure -> tensor-core Gram ->
exact centring -> Chebyshev-filtered subspace iteration (k = 256) -> projection.  Prints phase times and the principal
angle to the planted subspace.  Not a bench line."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
D, R, K = 10_000, 300, 256
dev = torch.device("cuda")
g = torch.Generator(device=dev); g.manual_seed(4242)
F = torch.linalg.qr(torch.randn((D, R), generator=g, device=dev, dtype=torch.float32))[0]
sig = 40.0 * torch.arange(1, R + 1, device=dev, dtype=torch.float32) ** -0.7
X = torch.empty((N, D), dtype=torch.uint8, device=dev)
for i in range(0, N, 10_000):
    n = min(10_000, N - i)
    L = torch.randn((n, R), generator=g, device=dev) * sig
    X[i:i + n] = (128 + L @ F.T + 4.0 * torch.randn((n, D), generator=g, device=dev)).round_().clamp_(0, 255).to(torch.uint8)
torch.cuda.synchronize()
configs = [("chol", "8", "0"), ("chol", "", "1"), ("chol", "8", "1"), ("chol", "8,16", "1"), ("chol", "8,8,24", "1"), ("jacobi", "", "1")]
if len(sys.argv) > 2:
    configs = [tuple(c.split(":")) for c in sys.argv[2:]]
for orth, degs, lock in configs:
    os.environ["EF_SUBSPACE_ORTH"], os.environ["EF_SUBSPACE_DEGREES"], os.environ["EF_SUBSPACE_LOCK"] = orth, degs, lock
    for rep in range(2):
        t0 = time.perf_counter()
        E, mean, proj, ev = ef.dist.fit_gen1_sharded(X, N, K, timings=True)
        torch.cuda.synchronize()
        print(f"orth {orth} degrees {degs or 'adaptive'} lock {lock} rep {rep}: fit_gen1_sharded N={N} D={D} k={K}: {time.perf_counter() - t0:.3f} s",
              {k: round(v, 4) for k, v in ef.dist.fit_gen1_sharded.last_timings.items()}, ef.dist.fit_gen1_sharded.last_solver_info, flush=True)
    print("   orthonormality:", float((E.T @ E - torch.eye(K, device=dev, dtype=torch.float64)).abs().max()), " ev[0], ev[255]:", float(ev[0]), float(ev[255]), flush=True)
# principal angles between the fitted top-K space and the planted factors (the top-K planted directions dominate)
cos = torch.linalg.svdvals(F[:, :20].double().T @ E)
print("smallest cosine of the 20 leading planted factors to the fitted space:", float(cos.min()), " eigenvalues[:3]", ev[:3].tolist(), " ev[255]", float(ev[255]))
print("orthonormality:", float((E.T @ E - torch.eye(K, device=dev, dtype=torch.float64)).abs().max()))
