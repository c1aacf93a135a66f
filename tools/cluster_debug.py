"""Debugging aid: compare the cluster kernel (mode 2) against the stream-K + epilogue path (mode 1) on a C2-shaped batch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(20250820)
D, K, NG, B = 10000, 10, int(os.environ.get("NG", 1024)), int(os.environ.get("B", 4096))
E = np.linalg.qr(rng.normal(size=(D, K)))[0]
mu = rng.uniform(60, 200, D)
lam = (np.arange(K) + 1.0) ** -1.0 * 4e5
G = rng.normal(0, 1, (NG, K)) * np.sqrt(lam)
rec = ef.Recognizer(E, mu, G, metric=ef.METRIC_COSINE_G1)
c = rng.normal(0, 1, (B, K)) * np.sqrt(lam)
Q = np.clip(np.rint(mu + c @ E.T + rng.normal(0, 8, (B, D))), 0, 255).astype(np.uint8)
outs = []
for mode in (1, 2, 2):
    rec.use_tensor_cores(mode)
    outs.append(rec.recognize(Q, 0.5))
    print("mode", mode, "timeouts", rec.pipeline_timeouts())
a, b, c2 = outs
for f in ("features", "score", "index", "label", "resid2"):
    x, y = getattr(a, f), getattr(b, f)
    bad = np.nonzero((x != y).reshape(B, -1).any(1))[0]
    print(f, "mismatch rows:", len(bad), bad[:20], "repeat-equal:", np.array_equal(y, getattr(c2, f)))
    if len(bad) and f in ("index", "score"):
        for r in bad[:5]:
            print("   row", r, "mode1", x[r], "mode2", y[r], "lane", r % 32, "rank", (r % 128) // 32, "tile", r // 128)
