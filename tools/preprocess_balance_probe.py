"""Is ef_preprocess limited by the imbalance of crop sizes?  Same pixel volume three ways: random sides 100..300 in queue
order, the same boxes sorted by decreasing area, and all boxes 208 x 208.  Not a bench line."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(5150)
F, H, W, nb = 8, 1080, 1920, 4096
frames = torch.randint(0, 256, (F, H, W), dtype=torch.uint8, device="cuda")


def boxes_for(side):
    return np.stack([rng.integers(0, F, nb), (rng.random(nb) * (W - side)).astype(np.int64),
                     (rng.random(nb) * (H - side)).astype(np.int64), side, side], axis=1).astype(np.int32)


side = rng.integers(100, 301, nb)
variants = {"random sides 100..300, queue order": boxes_for(side)}
b = variants["random sides 100..300, queue order"]
variants["same boxes, largest first"] = b[np.argsort(-b[:, 3].astype(np.int64) * b[:, 4], kind="stable")]
variants["all 208 x 208"] = boxes_for(np.full(nb, 208))
for name, bx in variants.items():
    boxes = torch.from_numpy(np.ascontiguousarray(bx)).cuda()
    out = ef.engine.preprocess_device(frames, boxes, 100)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ef.engine.preprocess_device(frames, boxes, 100, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    by = float((bx[:, 3].astype(np.int64) * bx[:, 4]).sum()) + nb * 10000
    print(f"{name:40s}: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.1f} GB/s algorithmic = {by / ms / 1e6 / 6550.1:.3f} of HBM peak", flush=True)
