"""End to end through the script-level callers on the device: crops on disk -> train_all_persons -> face_model.pkl ->
(a) the reference's own consumer code path (plain sklearn .transform + cosine_similarity on the unpickled objects) and
(b) the batched device scanner must agree; then the video loop on a synthetic clip."""
import json
import os
import pickle

import numpy as np
import pytest

import eigenfaces_b200 as ef
from gpu_util import require_gpu

pytestmark = pytest.mark.gpu
cv2 = pytest.importorskip("cv2")


def _persons(tmp_path, golden):
    """Two 'persons' cut from the golden light crops (100x100 gray), written as PNG-quality JPEGs."""
    X = golden("gen1_light.npz")["X_u8"].reshape(-1, 100, 100)
    base = str(tmp_path / "faces" / "lock_version")
    for name, rows in (("anna", range(0, 40)), ("bert", range(100, 130))):
        d = os.path.join(base, name)
        os.makedirs(d)
        for i, r in enumerate(rows):
            cv2.imwrite(os.path.join(d, f"face_{i:06d}_frame_{i:06d}.jpg"), cv2.cvtColor(X[r], cv2.COLOR_GRAY2BGR),
                        [cv2.IMWRITE_JPEG_QUALITY, 100])
    return base, X


def test_train_scripts_write_models_the_reference_consumer_can_use(tmp_path, golden):
    require_gpu()
    from sklearn.metrics.pairwise import cosine_similarity
    base, X = _persons(tmp_path, golden)
    ok, bad = ef.pipeline.train_all_persons(base)
    assert (ok, bad) == (2, 0)
    for name, n in (("anna", 40), ("bert", 30)):
        d = os.path.join(base, name)
        for f in ("face_model.pkl", "multi_person_mean_face.jpg", "multi_person_eigenface_01.jpg",
                  "multi_person_eigenface_10.jpg", "multi_person_model_info.json", f"{name}_faces_detection.json"):
            assert os.path.exists(os.path.join(d, f)), f
        info = json.load(open(os.path.join(d, "multi_person_model_info.json")))
        assert info["total_faces"] == n and info["n_components"] == n and abs(info["explained_variance_ratio"] - 1) < 1e-9
        m = pickle.load(open(os.path.join(d, "face_model.pkl"), "rb"))
        assert set(m) >= {"pca", "scaler", "face_features", "face_labels", "face_info", "person_id_map", "n_components",
                          "mean_face", "eigenfaces", "face_shape", "training_date"}
        # the reference's consumer (scan-template-v4.py:257-287) on the unpickled sklearn objects
        crops = [cv2.imread(os.path.join(d, f"face_{i:06d}_frame_{i:06d}.jpg")) for i in range(n)]
        flat = np.stack([cv2.resize(cv2.cvtColor(c, cv2.COLOR_BGR2GRAY), (64, 64)).flatten() for c in crops])
        feats = m["pca"].transform(m["scaler"].transform(flat))
        np.testing.assert_allclose(feats, m["face_features"], atol=1e-6)      # training crops reproduce the gallery
        sims = cosine_similarity(feats, m["face_features"])
        # the device scanner on the same crops
        scanner = ef.gen2.MultiModelFaceScanner()
        scanner.models[name] = {"model_data": m, "model_path": d}
        for i in (0, n // 2, n - 1):
            f_dev = scanner.extract_face_features(crops[i], m)
            np.testing.assert_allclose(f_dev, feats[i], rtol=1e-7, atol=1e-7)
            pid, pname, conf = scanner.recognize_face_all_models(crops[i], 0.8)
            assert pname == name and pid == 0 and abs(conf - sims[i].max()) < 1e-9
    # both models loaded: every crop is attributed to its own person
    scanner = ef.gen2.MultiModelFaceScanner()
    assert scanner.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    frame = np.zeros((100, 300, 3), np.uint8)
    frame[:, :100] = cv2.cvtColor(X[5], cv2.COLOR_GRAY2BGR)
    frame[:, 200:] = cv2.cvtColor(X[110], cv2.COLOR_GRAY2BGR)
    ids, names, confs = scanner.recognize_faces_all_models(frame, np.array([[0, 0, 100, 100], [200, 0, 100, 100]]), 0.8)
    assert names == ["anna", "bert"] and min(confs) > 0.9


def test_video_loop_on_a_synthetic_clip(tmp_path, golden):
    require_gpu()
    base, X = _persons(tmp_path, golden)
    assert ef.pipeline.train_person_model("anna", base, 20)               # train-v4 style: fixed component count
    scanner = ef.gen2.MultiModelFaceScanner()
    assert scanner.load_all_models(os.path.join(base, "*", "face_model.pkl"))
    clip = str(tmp_path / "clip.avi")
    vw = cv2.VideoWriter(clip, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (320, 240))
    if not vw.isOpened():
        pytest.skip("no video encoder in this OpenCV build")
    rng = np.random.default_rng(3)
    for t in range(6):
        fr = rng.integers(90, 110, (240, 320, 3), dtype=np.uint8)
        face = cv2.resize(cv2.cvtColor(X[t], cv2.COLOR_GRAY2BGR), (120, 120))
        fr[60:180, 100:220] = face
        vw.write(fr)
    vw.release()
    out = str(tmp_path / "results.json")
    res = ef.pipeline.process_video(clip, scanner, out, None, 0.5)
    assert res is not None and res["total_frames"] == 6 and os.path.exists(out)
    saved = json.load(open(out))
    # the reference's recognition_results.json schema (scripts/auto/scan-template-v2.py:442-502)
    assert {"video_path", "total_frames", "fps", "total_recognitions", "processing_date", "results"} <= set(saved)
    assert saved["total_recognitions"] == len(saved["results"])
    for d in saved["results"]:                                            # Haar may or may not fire on this clip
        assert {"frame_number", "timestamp", "x", "y", "width", "height", "person_id", "person_name", "confidence",
                "template_match_confidence", "ref_frame_diff"} <= set(d)
    # frames dealt over two ranks cover the same detections as one rank
    r0 = ef.pipeline.process_video(clip, scanner, None, None, 0.5, rank=0, world=2)
    r1 = ef.pipeline.process_video(clip, scanner, None, None, 0.5, rank=1, world=2)
    assert r0["total_recognitions"] + r1["total_recognitions"] == res["total_recognitions"]
