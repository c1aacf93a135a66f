"""The B = 1 call pattern of the reference (recognize_face_all_models over 4 person models) alone: for ncu launch lists
and host-side timing.  Not a bench line."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras  # noqa: E402
import eigenfaces_b200 as ef  # noqa: E402

print(json.dumps(bench_extras.latency_b1_section(ef, torch, torch.device("cuda"))))
