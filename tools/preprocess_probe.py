"""Micro-benchmark of ef_preprocess on the resize-active workload of bench.py (gray and BGR).  Not a bench line."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(5150)
F, H, W, nb = 8, 1080, 1920, 4096
side = rng.integers(100, 301, nb)
bx = np.stack([rng.integers(0, F, nb), (rng.random(nb) * (W - side)).astype(np.int64),
               (rng.random(nb) * (H - side)).astype(np.int64), side, side], axis=1).astype(np.int32)
boxes = torch.from_numpy(bx).cuda()
for channels in (1, 3):
    frames = torch.randint(0, 256, (F, H, W) if channels == 1 else (F, H, W, 3), dtype=torch.uint8, device="cuda")
    for out_side in (100, 64):
        out = ef.engine.preprocess_device(frames, boxes, out_side)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ef.engine.preprocess_device(frames, boxes, out_side, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        by = float((side.astype(np.int64) ** 2).sum()) * channels + nb * out_side * out_side
        print(f"channels {channels} -> {out_side}x{out_side}: {ms * 1e3:7.1f} us  {nb / ms / 1e3:6.2f} M crops/s  {by / ms / 1e6:7.1f} GB/s algorithmic",
              flush=True)
