"""Warm per-kernel durations of the template-matching detector (bench shape: 640x480 frame, 60 jobs)."""
import os
import sys
from collections import defaultdict

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

rng = np.random.default_rng(640480)
frame = rng.integers(0, 256, (480, 640), dtype=np.uint8)
tmpls = [rng.integers(0, 256, (int(rng.integers(80, 121)), int(rng.integers(80, 121))), dtype=np.uint8) for _ in range(20)]
m = ef.template.TemplateMatcher(tmpls)
fd = torch.from_numpy(frame).cuda()
for _ in range(3):
    m.match(fd)
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        m.match(fd)
    torch.cuda.synchronize()
tot, cnt = defaultdict(float), defaultdict(int)
for ev in prof.events():
    if "cuda" in str(ev.device_type).lower():
        tot[ev.name] += ev.device_time
        cnt[ev.name] += 1
for k, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    short = k.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
    print(f"{short[:50]:50s} {t / n:9.1f} us per frame ({cnt[k] / n:.1f} launches)")
