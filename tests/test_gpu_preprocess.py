"""K1 parity: the CUDA preprocess kernel against the oracle (== cv2, see test_oracle_golden) -- bit exact."""
import numpy as np
import pytest

import eigenfaces_b200 as ef
from gpu_util import require_gpu
from oracle import preprocess

pytestmark = pytest.mark.gpu


def _run(frames, boxes, side):
    torch = require_gpu()
    fr = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    bx = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.int32)).cuda()
    out = ef.preprocess_device(fr, bx, side)
    torch.cuda.synchronize()
    return out[:, :side * side].cpu().numpy()


def test_golden_single_crops(golden):
    g = golden("preprocess.npz")
    for i, (h, w, c, dw, dh, seed) in enumerate(g["specs"]):
        shape = (h, w, 3) if c == 3 else (h, w)
        img = np.random.default_rng(int(seed)).integers(0, 256, shape, dtype=np.uint8)
        got = _run(img[None], [[0, 0, 0, w, h]], int(dw))[0].reshape(dh, dw)
        assert np.array_equal(got, g[f"out_{i:02d}"]), f"case {i}: {shape} -> {dw}x{dh}"


def test_golden_rois_in_bgr_frame(golden):
    g = golden("preprocess.npz")
    frame = np.random.default_rng(int(g["frame_seed"])).integers(0, 256, tuple(g["frame_shape"]), dtype=np.uint8)
    boxes = np.concatenate([np.zeros((len(g["boxes"]), 1), np.int32), g["boxes"]], axis=1)
    assert np.array_equal(_run(frame[None], boxes, 100), g["roi_out_100"])


@pytest.mark.parametrize("channels,side", [(1, 64), (3, 64), (1, 100), (3, 100)])
def test_random_boxes_match_oracle(channels, side):
    rng = np.random.default_rng(100 * channels + side)
    F, H, W = 3, 540, 960
    frames = rng.integers(0, 256, (F, H, W, 3) if channels == 3 else (F, H, W), dtype=np.uint8)
    boxes = []
    for i in range(400):
        w = int(rng.integers(1, 400)); h = w if i % 2 else int(rng.integers(1, 400))
        if i % 40 == 0:
            w = h = 2 * side           # INTER_AREA special case
        if i % 40 == 1:
            w = h = side               # identity
        x = int(rng.integers(0, W - w + 1)); y = int(rng.integers(0, H - h + 1))
        boxes.append((int(rng.integers(0, F)), x, y, w, h))
    got = _run(frames, boxes, side)
    for i, (f, x, y, w, h) in enumerate(boxes):
        want = preprocess.preprocess_crop(frames[f, y:y + h, x:x + w], side, side)
        assert np.array_equal(got[i], want), f"box {i}: {(f, x, y, w, h)}"


def test_full_size_frame_1080p_and_bad_box():
    """BASELINE config 5 shape: 1080p BGR frames; an out-of-frame box yields zeros instead of reading out of bounds."""
    rng = np.random.default_rng(5150)
    frame = rng.integers(0, 256, (1, 1080, 1920, 3), dtype=np.uint8)
    boxes = [(0, 1920 - 300, 1080 - 300, 300, 300), (0, 0, 0, 1920, 1080), (0, 1900, 1000, 64, 128), (0, 5, 5, 0, 10)]
    torch = require_gpu()
    fr = torch.from_numpy(frame).cuda()
    bx = torch.tensor(boxes, dtype=torch.int32, device="cuda")
    # asynchronous form: the caller owns the counter of bad boxes; their crops are all zero
    bad = torch.zeros(1, dtype=torch.int32, device="cuda")
    got = ef.preprocess_device(fr, bx, 64, bad=bad)[:, :4096].cpu().numpy()
    assert int(bad.item()) == 2
    assert np.array_equal(got[0], preprocess.preprocess_crop(frame[0, 780:, 1620:], 64, 64))
    assert np.array_equal(got[1], preprocess.preprocess_crop(frame[0], 64, 64))
    assert not got[2].any() and not got[3].any()
    # checking form (no counter passed): a black crop is never returned silently
    with pytest.raises(ef.EigenfacesError, match="2 of 4 boxes"):
        ef.preprocess_device(fr, bx, 64)
    # host entry point of a model: EF_ERR_INVALID instead of a normal-looking label for a black crop
    rng2 = np.random.default_rng(1)
    E = np.linalg.qr(rng2.normal(size=(4096, 6)))[0]
    rec = ef.Recognizer(E, rng2.uniform(0, 255, 4096), rng2.normal(size=(9, 6)))
    with pytest.raises(ef.EigenfacesError, match="not inside their frame"):
        rec.recognize_boxes(frame[0], [b[1:] for b in boxes], 64)
    ok = rec.recognize_boxes(frame[0], [b[1:] for b in boxes[:2]], 64)
    assert ok.index.shape == (2,)
    # device entry point: counted, read (and cleared) by bad_boxes()
    rec.recognize_boxes_device(fr, bx, 64)
    assert rec.bad_boxes() == 2 and rec.bad_boxes() == 0
    with pytest.raises(ef.EigenfacesError):
        rec.recognize_boxes_device(fr, bx, 64)
        rec.check_device_results()
    rec.close()


def test_empty_batch():
    torch = require_gpu()
    fr = torch.zeros((1, 8, 8), dtype=torch.uint8, device="cuda")
    bx = torch.zeros((0, 5), dtype=torch.int32, device="cuda")
    assert ef.preprocess_device(fr, bx, 64).shape == (0, 4096)
