import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import eigenfaces_b200 as ef
g = np.load("/root/repo/tests/golden/preprocess.npz")
h, w, c, dw, dh, seed = [int(v) for v in g["specs"][11]]
img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
fr = torch.from_numpy(img[None]).cuda(); bx = torch.tensor([[0, 0, 0, w, h]], dtype=torch.int32, device="cuda")
for flags in ("0", "1", "2", "3"):
    os.environ["EF_PRE_DEBUG"] = flags
    got = ef.preprocess_device(fr, bx, dw)[0, :dw * dh].cpu().numpy().reshape(dh, dw)
    bad = np.argwhere(got != g["out_11"])
    print("flags", flags, "mismatches", len(bad), bad[:6].tolist(), [(int(got[y, x]), int(g["out_11"][y, x])) for y, x in bad[:6]])
