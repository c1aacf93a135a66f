"""Oracle for the north-star extras -- test infrastructure only.

PARITY UNPINNED: the reference has no L2 matcher and no reconstruction error anywhere (grep for
reconstruct|euclid|cdist|argmin over all .py finds only the abandoned useless/scan-enhanced.py:311-315
ensemble score).  BASELINE.json's north_star asks for both, so the oracle is the textbook float64
formula, stated in its direct (not expanded) form so that it is independent of the engine's algebra.
"""
import numpy as np


def l2_nearest(p, gallery):
    """argmin_j ||p_b - g_j||^2 by explicit differences.  Returns (dist2[B], idx[B]); ties -> lowest j."""
    p = np.asarray(p, dtype=np.float64)
    gallery = np.asarray(gallery, dtype=np.float64)
    best = np.empty(len(p))
    idx = np.empty(len(p), dtype=np.int64)
    for b in range(len(p)):
        d = gallery - p[b]
        d2 = np.einsum('ij,ij->i', d, d)
        idx[b] = int(np.argmin(d2))
        best[b] = d2[idx[b]]
    return best, idx


def reconstruction_error2(v, basis):
    """Squared distance from face space: || v - E (E^T v) ||^2 for rows of v [B, D], basis E [D, k].

    Gen-1: v = x - mean_face, E = eigenfaces (useless/train.py:94-95 normalises the columns).
    Gen-2: v = (x - scaler.mean_) / scaler.scale_ - pca.mean_, E = pca.components_.T.
    """
    v = np.asarray(v, dtype=np.float64)
    p = v @ basis
    r = v - p @ basis.T
    return np.einsum('ij,ij->i', r, r)
