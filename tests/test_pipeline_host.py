"""Host-side callers of the hot path (no GPU): crop discovery, detection-JSON regeneration, script shims parse."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import eigenfaces_b200 as ef
from conftest import ROOT

cv2 = pytest.importorskip("cv2")


def _write_person(base, name, n, rng, side=80):
    d = os.path.join(base, name)
    os.makedirs(d, exist_ok=True)
    for i in range(n):
        img = rng.integers(0, 256, (side, side + 7, 3), dtype=np.uint8)
        cv2.imwrite(os.path.join(d, f"face_{i:06d}_frame_{3 * i:06d}.jpg"), img)
    cv2.imwrite(os.path.join(d, "multi_person_eigenface_01.jpg"), np.zeros((64, 64), np.uint8))   # must be skipped
    cv2.imwrite(os.path.join(d, "multi_person_mean_face.jpg"), np.zeros((64, 64), np.uint8))
    return d


def test_crop_discovery_and_detection_json(tmp_path):
    rng = np.random.default_rng(0)
    base = str(tmp_path / "lock_version")
    _write_person(base, "alice", 5, rng)
    _write_person(base, "bob", 3, rng)
    assert ef.pipeline.count_face_images(base) == 8                       # base directory: sum over persons
    assert ef.pipeline.count_face_images(os.path.join(base, "alice")) == 5
    assert ef.pipeline.count_face_images(str(tmp_path / "nope")) == 0
    path = ef.pipeline.generate_detection_json_for_person("alice", os.path.join(base, "alice"))
    data = json.load(open(path))
    assert os.path.basename(path) == "alice_faces_detection.json"
    assert data["total_faces_detected"] == 5 and data["fps"] == 30.0 and data["total_frames"] == 13
    row = data["faces"][2]
    # keys of detection-v4.py:71-84 / train-v5.py:108-121
    assert set(row) == {"face_id", "frame_number", "timestamp", "x", "y", "width", "height", "center_x", "center_y",
                        "area", "image_path", "image_filename"}
    assert row["frame_number"] == 6 and row["width"] == 87 and row["height"] == 80 and row["area"] == 87 * 80
    assert row["image_filename"] == "face_000002_frame_000006.jpg"
    assert ef.pipeline.generate_detection_json_for_person("carol", str(tmp_path)) is None


def test_script_shims_expose_the_reference_cli():
    for script, flags in (("train-v4.py", ["--person"]), ("detection-v4.py", ["--video", "--person"]),
                          ("scan-template-v4.py", ["--video"]), ("run_pipeline.py", ["--video", "--person"])):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), "--help"], capture_output=True,
                             text=True, timeout=120)
        assert res.returncode == 0, res.stderr[-500:]
        for f in flags:
            assert f in res.stdout
    assert os.path.exists(os.path.join(ROOT, "scripts", "train-v5.py"))
