import numpy as np
import pytest


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


def face_like(rng, X, n, noise=8.0):
    """Perturbed training crops: clip(round(x + N(0, noise^2)))."""
    base = X[rng.integers(0, len(X), n)].astype(np.float64)
    return np.clip(np.rint(base + rng.normal(0, noise, base.shape)), 0, 255).astype(np.uint8)
