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


def test_scanner_loads_templates_like_the_reference(tmp_path):
    """load_all_models (scan-template-v4.py:17-74) without a GPU: the model pickle, the detection JSON and the first
    five face crops as gray templates -- found through image_path, through image_path with normalised separators
    (detection-v4.py writes Windows paths) or through image_filename inside the person's directory."""
    import pickle
    rng = np.random.default_rng(1)
    base = str(tmp_path / "faces" / "lock_version")
    d = _write_person(base, "carol", 7, rng)
    assert ef.pipeline.generate_detection_json_for_person("carol", d)
    info_path = os.path.join(d, "carol_faces_detection.json")
    info = json.load(open(info_path))
    for face in info["faces"]:                                    # what the reference's Windows run would have written
        face["image_path"] = "faces\\lock_version\\carol\\" + face["image_filename"]
    json.dump(info, open(info_path, "w"))
    pickle.dump({"face_features": np.zeros((7, 3)), "face_labels": np.zeros(7, int), "person_id_map": {"carol": 0}},
                open(os.path.join(d, "face_model.pkl"), "wb"))
    sc = ef.gen2.MultiModelFaceScanner()
    assert sc.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    m = sc.models["carol"]
    assert len(m["template_images"]) == 5 and len(m["detection_data"]["faces"]) == 7
    t = m["template_images"][0]
    assert t["image"].ndim == 2 and t["image"].dtype == np.uint8 and (t["width"], t["height"]) == (87, 80)
    # no detection JSON -> no templates, like the reference; the detector then reports nothing for that person
    os.remove(info_path)
    sc2 = ef.gen2.MultiModelFaceScanner()
    assert sc2.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    assert sc2.models["carol"]["template_images"] == [] and sc2.models["carol"]["detection_data"] is None
    assert sc2.template_match_all_models(np.zeros((120, 160), np.uint8)) == []
    assert not ef.gen2.MultiModelFaceScanner().load_all_models(str(tmp_path / "nothing" / "*.pkl"))
