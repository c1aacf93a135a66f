"""Oracle for the crop preprocess -- test infrastructure only.

The reference calls un-vendored OpenCV for this step:
  * cv2.cvtColor(face_img, cv2.COLOR_BGR2GRAY)   scan-template-v4.py:257-260, train-v5.py:254,329
  * cv2.resize(gray, (64, 64)) / (sqrt(D), sqrt(D))  scan-template-v4.py:262, train-v5.py:255,330,
                                                      useless/scan.py:198-199,248-249
  * .flatten()                                    scan-template-v4.py:263, useless/scan.py:202,252
OpenCV (author: opencv-python 4.8.1.78 per useless/requirements.txt:3; container: 4.13.0) is not
in /root/reference, so its published fixed-point algorithm is restated here in integer numpy and
pinned against cv2 itself (tests/golden/preprocess.npz + live comparison when cv2 is importable).

Fixed-point spec (OpenCV imgproc: color_rgb.simd.hpp RGB2Gray<uchar>, resize.cpp
HResizeLinear/VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>):
  gray   = (3735*B + 19235*G + 9798*R + (1 << 14)) >> 15
  resize : scale = 1.0 / (dst / src) in double; per destination index d
             f = float32((d + 0.5) * scale - 0.5); s = floor(f); f -= s          (float32)
           x axis: s < 0 -> (s, f) = (0, 0); s >= w-1 -> (s, f) = (w-1, 0)
                   a0 = rint((1 - f) * 2048), a1 = rint(f * 2048)                 (float32, half-even)
                   H[r, d] = src[r, s] * a0 + src[r, min(s+1, w-1)] * a1          (int32)
           y axis: weights from the UNCLAMPED f, row indices clipped into [0, h-1]
                   out = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2
           special case: src == 2*dst on both axes -> INTER_AREA 2x2 box (a+b+c+d+2) >> 2
           special case: src == dst -> copy
"""
import numpy as np

GRAY_B, GRAY_G, GRAY_R, GRAY_SHIFT = 3735, 19235, 9798, 15
COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def bgr_to_gray(img):
    """cv2.cvtColor(img, COLOR_BGR2GRAY) for uint8 [h, w, 3] (P1)."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
    b = img[..., 0].astype(np.int32)
    g = img[..., 1].astype(np.int32)
    r = img[..., 2].astype(np.int32)
    y = (GRAY_B * b + GRAY_G * g + GRAY_R * r + (1 << (GRAY_SHIFT - 1))) >> GRAY_SHIFT
    return y.astype(np.uint8)


def linear_coeffs(src, dst, clamp_frac):
    """Per-destination source index and 11-bit weights along one axis.

    clamp_frac=True is the x-axis rule (fraction zeroed when the index is clamped);
    clamp_frac=False is the y-axis rule (weights from the unclamped fraction; the caller clips rows).
    Returns (s0, s1, w0, w1) as int32 arrays of length dst.
    """
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_frac:
        lo = s < 0
        s = np.where(lo, 0, s)
        f = np.where(lo, np.float32(0), f)
        hi = s >= src - 1
        s = np.where(hi, src - 1, s)
        f = np.where(hi, np.float32(0), f)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int32)
    s0 = np.clip(s, 0, src - 1).astype(np.int32)
    s1 = np.clip(s + 1, 0, src - 1).astype(np.int32)
    return s0, s1, w0, w1


def resize_linear_u8(gray, dw, dh):
    """cv2.resize(gray, (dw, dh)) (default INTER_LINEAR) for uint8 [h, w] (P2)."""
    gray = np.asarray(gray)
    assert gray.dtype == np.uint8 and gray.ndim == 2
    h, w = gray.shape
    if w == dw and h == dh:
        return gray.copy()
    if w == 2 * dw and h == 2 * dh:
        g = gray.astype(np.int32)
        return ((g[0::2, 0::2] + g[0::2, 1::2] + g[1::2, 0::2] + g[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = linear_coeffs(w, dw, clamp_frac=True)
    y0, y1, b0, b1 = linear_coeffs(h, dh, clamp_frac=False)
    g = gray.astype(np.int32)
    hbuf = g[:, x0] * a0[None, :] + g[:, x1] * a1[None, :]          # [h, dw] int32
    h0 = hbuf[y0, :] >> 4
    h1 = hbuf[y1, :] >> 4
    out = (((b0[:, None] * h0) >> 16) + ((b1[:, None] * h1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess_crop(face_img, dw, dh):
    """gray -> resize -> flatten, as scan-template-v4.py:257-263 / useless/scan.py:245-252 (P1-P3).

    face_img: uint8 [h, w, 3] BGR or uint8 [h, w] gray.  Returns uint8 [dw*dh].
    """
    face_img = np.asarray(face_img)
    gray = bgr_to_gray(face_img) if face_img.ndim == 3 else face_img
    return resize_linear_u8(gray, dw, dh).reshape(-1)


def preprocess_boxes(frame, boxes, dw, dh):
    """Batched form: crop frame[y:y+h, x:x+w] for each (x, y, w, h) then preprocess_crop.

    Mirrors the per-detection loop of scan-template-v4.py:360/:390 and useless/scan.py:243-252.
    Returns uint8 [B, dw*dh].
    """
    out = np.empty((len(boxes), dw * dh), dtype=np.uint8)
    for i, (x, y, w, h) in enumerate(boxes):
        out[i] = preprocess_crop(frame[y:y + h, x:x + w], dw, dh)
    return out
