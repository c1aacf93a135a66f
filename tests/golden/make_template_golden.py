"""Generates tests/golden/template_match.npz by running OpenCV itself (cv2.resize + cv2.matchTemplate
TM_CCOEFF_NORMED + cv2.minMaxLoc, the calls of scan-template-v4.py:167-171) on seeded synthetic frames.
Run in the build container (cv2 4.13.0): python tests/golden/make_template_golden.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(4131)
out = {}
cases = []
# (frame W, H, [(template w, h)])
for ci, (W, H, tmpls) in enumerate([(160, 120, [(30, 26), (41, 37)]), (131, 97, [(25, 33)]), (200, 90, [(64, 40), (21, 20)])]):
    # smooth background + noise, templates cut from a face-like blob pasted into the frame at known places
    yy, xx = np.mgrid[0:H, 0:W]
    frame = 110 + 40 * np.sin(xx / 17.0) * np.cos(yy / 11.0) + rng.normal(0, 12, (H, W))
    templates = []
    for (w, h) in tmpls:
        ty, tx = np.mgrid[0:h, 0:w]
        blob = 128 + 90 * np.exp(-(((tx - w / 2) / (w / 4)) ** 2 + ((ty - h / 2) / (h / 3)) ** 2)) - 60 * (np.abs(tx - w / 2) < w / 10)
        blob = np.clip(np.round(blob + rng.normal(0, 6, (h, w))), 0, 255).astype(np.uint8)
        templates.append(blob)
        px, py = int(rng.integers(8, W - w - 8)), int(rng.integers(8, H - h - 8))
        frame[py:py + h, px:px + w] = 0.7 * blob + 0.3 * frame[py:py + h, px:px + w]
    frame = np.clip(np.round(frame), 0, 255).astype(np.uint8)
    out[f"c{ci}_frame"] = frame
    for ti, t in enumerate(templates):
        out[f"c{ci}_t{ti}"] = t
        for scale in (0.8, 1.0, 1.2):
            nw, nh = int(t.shape[1] * scale), int(t.shape[0] * scale)
            if nw < 20 or nh < 20 or nw > W or nh > H:
                continue
            st = cv2.resize(t, (nw, nh))
            res = cv2.matchTemplate(frame, st, cv2.TM_CCOEFF_NORMED)
            _, max_val, _, max_loc = cv2.minMaxLoc(res)
            tag = f"c{ci}_t{ti}_s{int(scale * 10)}"
            out[tag + "_templ"] = st
            out[tag + "_map"] = res
            out[tag + "_best"] = np.array([max_val, max_loc[0], max_loc[1]], dtype=np.float64)
            cases.append(tag)
# degenerate inputs: flat template (all ones map), flat frame region (zero scores)
frame = out["c0_frame"].copy()
frame[10:60, 20:90] = 77
flat_t = np.full((22, 24), 140, np.uint8)
tex_t = out["c0_t0"]
out["d_frame"] = frame
out["d_flat_templ"] = flat_t
out["d_flat_map"] = cv2.matchTemplate(frame, flat_t, cv2.TM_CCOEFF_NORMED)
out["d_tex_templ"] = tex_t
out["d_tex_map"] = cv2.matchTemplate(frame, tex_t, cv2.TM_CCOEFF_NORMED)
out["cases"] = np.array(cases)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "template_match.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes;", len(cases), "cases")
