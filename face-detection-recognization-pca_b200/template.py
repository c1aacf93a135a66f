"""Template-matching detector on the device (K6): cv2.matchTemplate(TM_CCOEFF_NORMED) + cv2.minMaxLoc for every
(template image, scale) of every person against one gray frame in a single batch of launches.

Mirrors scan-template-v4.py:
  TemplateMatcher(templates, scales)            the per-frame loop body :147-174 (resize :167, matchTemplate :170,
                                                minMaxLoc :171); templates are scaled once, not once per frame
  is_detection_in_corner / calculate_overlap / non_max_suppression      :75-127, :199-251 (host logic, verbatim rules)
  template_match_models(models, frame)          template_match_all_models :129-197
The product path needs the CUDA library (no CPU fallback); the oracle lives in oracle/template_match.py.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

SCALES = (0.8, 1.0, 1.2)          # scan-template-v4.py:159
MAX_JOBS = 64                     # per ef_template_match_device call


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("eigenfaces_b200.template needs a CUDA device (no CPU fallback)")
    return torch


def resize_gray_device(img, new_w, new_h):
    """cv2.resize(img, (new_w, new_h)) (INTER_LINEAR, bit exact: K1) for one gray uint8 image; returns a CUDA tensor
    [new_h, new_w]."""
    torch = _torch()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    src = torch.from_numpy(img).to(dev)
    boxes = torch.tensor([[0, 0, 0, w, h]], dtype=torch.int32, device=dev)
    out = torch.zeros((1, new_w * new_h), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().ef_preprocess(src.data_ptr(), src.numel(), src.stride(0), w, h, 1, 1, boxes.data_ptr(), 1,
                                   int(new_w), int(new_h), out.data_ptr(), out.stride(0), None, C.c_void_p(stream)),
          "ef_preprocess")
    return out.view(new_h, new_w)


class TemplateMatcher:
    """Templates of one or more persons, pre-scaled on the device.

    templates: list of 2-D uint8 arrays.  jobs = every (template, scale) with int(w * scale) >= 20 and
    int(h * scale) >= 20 (scan-template-v4.py:161-165), in the reference's loop order (template outer, scale inner)."""

    def __init__(self, templates, scales=SCALES):
        torch = _torch()
        self.jobs = []                     # (template index, scale, w, h)
        parts = []
        off = 0
        self.offsets = []
        for ti, t in enumerate(templates):
            t = np.ascontiguousarray(t, dtype=np.uint8)
            if t.ndim != 2:
                raise ValueError("templates must be gray (2-D uint8) images")
            for scale in scales:
                new_w, new_h = int(t.shape[1] * scale), int(t.shape[0] * scale)
                if new_w < 20 or new_h < 20 or new_w > 1024:
                    continue
                parts.append(resize_gray_device(t, new_w, new_h).reshape(-1))
                self.jobs.append((ti, scale, new_w, new_h))
                self.offsets.append(off)
                off += new_w * new_h
        dev = torch.device("cuda", torch.cuda.current_device())
        self.bytes = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.uint8, device=dev)
        self._work = None

    def scaled_template(self, job):
        ti, scale, w, h = self.jobs[job]
        o = self.offsets[job]
        return self.bytes[o:o + w * h].view(h, w)

    def match(self, frame, want_maps=False):
        """frame: gray uint8 [H, W] (numpy or CUDA tensor).  Returns a list with one entry per job: None when the
        scaled template does not fit the frame (:164), else dict(max_val, x, y, width, height, scale, template[, map])."""
        torch = _torch()
        dev = torch.device("cuda", torch.cuda.current_device())
        if not torch.is_tensor(frame):
            frame = torch.from_numpy(np.ascontiguousarray(frame, dtype=np.uint8)).to(dev)
        if frame.dim() != 2 or frame.dtype != torch.uint8 or frame.stride(1) != 1:
            raise ValueError("frame must be a gray uint8 image with unit column stride")
        H, W = frame.shape
        live = [i for i, (_, _, w, h) in enumerate(self.jobs) if w <= W and h <= H]
        out = [None] * len(self.jobs)
        lib = _lib.lib()
        stream = torch.cuda.current_stream(dev).cuda_stream
        for c0 in range(0, len(live), MAX_JOBS):
            chunk = live[c0:c0 + MAX_JOBS]
            n = len(chunk)
            t_off = (C.c_int64 * n)(*[self.offsets[i] for i in chunk])
            tw = (C.c_int32 * n)(*[self.jobs[i][2] for i in chunk])
            th = (C.c_int32 * n)(*[self.jobs[i][3] for i in chunk])
            sizes = [(W - self.jobs[i][2] + 1) * (H - self.jobs[i][3] + 1) for i in chunk]
            r_offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
            r_off = (C.c_int64 * n)(*[int(v) for v in r_offs[:-1]])
            maps = torch.empty(int(r_offs[-1]), dtype=torch.float32, device=dev) if want_maps else None
            need = lib.ef_template_match_work_bytes(W, H, n, tw, th)
            if need == 0:
                raise RuntimeError("ef_template_match_work_bytes rejected the job list")
            if self._work is None or self._work.numel() < need:
                self._work = torch.empty(need, dtype=torch.uint8, device=dev)
            best_val = torch.empty(n, dtype=torch.float64, device=dev)
            best_xy = torch.empty((n, 2), dtype=torch.int32, device=dev)
            check(lib.ef_template_match_device(frame.data_ptr(), frame.stride(0), W, H, self.bytes.data_ptr(), t_off, tw,
                                               th, n, maps.data_ptr() if want_maps else None, r_off,
                                               best_val.data_ptr(), best_xy.data_ptr(), self._work.data_ptr(),
                                               self._work.numel(), C.c_void_p(stream)), "ef_template_match_device")
            vals = best_val.cpu().numpy()
            xy = best_xy.cpu().numpy()
            for q, i in enumerate(chunk):
                ti, scale, w, h = self.jobs[i]
                d = dict(max_val=float(vals[q]), x=int(xy[q, 0]), y=int(xy[q, 1]), width=w, height=h, scale=scale,
                         template=ti)
                if want_maps:
                    d["map"] = maps[int(r_offs[q]):int(r_offs[q + 1])].view(H - h + 1, W - w + 1)
                out[i] = d
        return out


# ------------------------------------------------------------------------------------------- host rules (verbatim)
def is_detection_in_corner(detection, frame_width, frame_height, corner_threshold=0.15, border_threshold=0.05):
    """scan-template-v4.py:75-127."""
    x, y, w, h = detection['x'], detection['y'], detection['width'], detection['height']
    corner_w, corner_h = int(frame_width * corner_threshold), int(frame_height * corner_threshold)
    border_w, border_h = int(frame_width * border_threshold), int(frame_height * border_threshold)
    center_x, center_y = x + w // 2, y + h // 2
    if x < border_w or y < border_h or (x + w) > (frame_width - border_w) or (y + h) > (frame_height - border_h):
        return True
    if center_x < corner_w and center_y < corner_h:
        return True
    if center_x > (frame_width - corner_w) and center_y < corner_h:
        return True
    if center_x < corner_w and center_y > (frame_height - corner_h):
        return True
    if center_x > (frame_width - corner_w) and center_y > (frame_height - corner_h):
        return True
    return False


def calculate_overlap(det1, det2):
    """scan-template-v4.py:224-251 (intersection over union)."""
    x1_min, y1_min = det1['x'], det1['y']
    x1_max, y1_max = x1_min + det1['width'], y1_min + det1['height']
    x2_min, y2_min = det2['x'], det2['y']
    x2_max, y2_max = x2_min + det2['width'], y2_min + det2['height']
    ix0, iy0, ix1, iy1 = max(x1_min, x2_min), max(y1_min, y2_min), min(x1_max, x2_max), min(y1_max, y2_max)
    if ix1 <= ix0 or iy1 <= iy0:
        return 0.0
    inter = (ix1 - ix0) * (iy1 - iy0)
    union = det1['width'] * det1['height'] + det2['width'] * det2['height'] - inter
    return inter / union if union > 0 else 0.0


def non_max_suppression(detections, overlap_threshold=0.3):
    """scan-template-v4.py:199-222."""
    if len(detections) == 0:
        return []
    detections = sorted(detections, key=lambda d: d['confidence'], reverse=True)
    keep = []
    while detections:
        current = detections.pop(0)
        keep.append(current)
        detections = [d for d in detections if calculate_overlap(current, d) < overlap_threshold]
    return keep


def select_best_match(person_name, results, frame_width, frame_height):
    """The selection of scan-template-v4.py:144-191 over one person's job results (reference loop order): the highest
    max_val wins (strict >), candidates in a corner / border area are skipped, the winner must exceed 0.6."""
    best_match, best_score = None, 0.0
    for r in results:
        if r is None:
            continue
        if r['max_val'] > best_score:
            cand = {'x': r['x'], 'y': r['y'], 'width': r['width'], 'height': r['height'], 'person_name': person_name,
                    'confidence': r['max_val'], 'scale': r['scale']}
            if not is_detection_in_corner(cand, frame_width, frame_height):
                best_score = r['max_val']
                best_match = cand
    if best_match and best_score > 0.6:
        return best_match
    return None
