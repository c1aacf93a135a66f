"""GPU parity of the round-2 rows: the scripts/manual generation (F3), the Gen-1 model store with real reference
pickles, the Gen-1 per-frame recognition loop."""
import io
import os
import pickle

import cv2
import numpy as np
import pytest

import eigenfaces_b200 as ef
from conftest import GOLDEN
from gpu_util import require_gpu
from oracle import gen1 as ogen1
from oracle import gen2 as ogen2
from oracle import preprocess as opre

pytestmark = pytest.mark.gpu


def _align(A, B):
    s = np.sign(np.sum(A * B, axis=1))
    s[s == 0] = 1.0
    return s


def test_manual_fit_matches_the_reference(golden):
    """ef_fit_manual_host against ManualStandardScaler + ManualPCA of scripts/manual/train-v2.py (live outputs)."""
    require_gpu()
    g = golden("manual_joseph.npz")
    X = golden("gen2_joseph.npz")["X_u8"].copy()
    k = int(g["k"])
    fit = ef.engine.fit_manual(X, k)
    np.testing.assert_allclose(fit["scaler_mean"], g["ref_scaler_mean"], rtol=1e-14)
    np.testing.assert_allclose(fit["scaler_scale"], g["ref_scaler_scale"], rtol=1e-12)
    np.testing.assert_allclose(fit["mean_face"], g["ref_mean_face"], rtol=1e-14)
    s = _align(fit["components"], g["ref_components"])
    np.testing.assert_allclose(fit["components"] * s[:, None], g["ref_components"], atol=1e-8)
    np.testing.assert_allclose(fit["explained_variance_ratio"], g["ref_evr"], rtol=1e-9)
    np.testing.assert_allclose(fit["features"] * s[None, :], g["ref_features"], atol=1e-6)
    # the std == 0 -> 1 rule (train-v2.py:61-62): a constant pixel column
    X[:, 17] = 9
    m, v, sc = ef.engine.scaler_fit_u8(X, 1)
    assert sc[17] == 1.0 and v[17] == 0.0 and m[17] == 9.0
    mo, so = ogen2.manual_scaler_fit(X)
    np.testing.assert_allclose(sc, so, rtol=1e-12)


def test_manual_estimator_classes_on_float_data():
    """ManualPCA.fit_transform / transform on arbitrary float64 data (ef_pca_fit_f64_host + device GEMM)."""
    require_gpu()
    rng = np.random.default_rng(5)
    Z = rng.normal(size=(60, 300)) @ np.diag(np.linspace(3, 0.1, 300))
    pca = ef.manual.ManualPCA(n_components=7)
    F = pca.fit_transform(Z)
    ref = ogen2.manual_pca_fit(Z, 7)
    s = _align(pca.components_, ref["components"])
    np.testing.assert_allclose(pca.components_ * s[:, None], ref["components"], atol=1e-9)
    np.testing.assert_allclose(pca.explained_variance_ratio_, ref["explained_variance_ratio"], rtol=1e-9)
    np.testing.assert_allclose(F * s[None, :], (Z - ref["mean"]) @ ref["components"].T, atol=1e-9)
    np.testing.assert_allclose(pca.transform(Z[:5]), F[:5], atol=1e-9)
    Xu = rng.integers(0, 256, (40, 128), dtype=np.uint8)
    Xu[:, 3] = 200
    sc = ef.manual.ManualStandardScaler()
    Zs = sc.fit_transform(Xu)
    mo, so = ogen2.manual_scaler_fit(Xu)
    np.testing.assert_allclose(Zs, (Xu - mo) / so, atol=1e-12)


def test_manual_scanner_on_the_reference_pickle(golden):
    """A model the reference's train-v2.py wrote, recognised through FaceScanner: features / ids / names / confidences of
    scripts/manual/scan-template-v2.py's own extract_face_features + recognize_face."""
    require_gpu()
    g = golden("manual_joseph.npz")
    crops = golden("gen2_recog.npz")
    scanner = ef.manual.FaceScanner(os.path.join(GOLDEN, "manual_model.pkl"))
    assert scanner.load_model_and_data()
    for i in range(len(g["recog_conf"])):
        f = scanner.extract_face_features(crops[f"crop_{i:02d}"])
        np.testing.assert_allclose(f, g["recog_features"][i], rtol=1e-9, atol=1e-9)
        pid, name, conf = scanner.recognize_face(f, threshold=0.7)
        assert pid == int(g["recog_pid"][i]) and name == str(g["recog_name"][i])
        assert abs(conf - float(g["recog_conf"][i])) < 1e-12
    X = golden("gen2_joseph.npz")["X_u8"]
    rows = g["self_rows"]
    res = scanner.recognize_crops(X[rows].reshape(len(rows), 64, 64), [[i, 0, 0, 64, 64] for i in range(len(rows))], 0.7)
    np.testing.assert_allclose(res.score, g["self_conf"], atol=1e-12)
    assert np.array_equal(res.index, g["self_idx"])
    v = g["recog_features"][0]
    assert abs(scanner.manual_cosine_similarity(v, 2.5 * v) - 1.0) < 1e-12
    assert scanner.manual_cosine_similarity(v, np.zeros_like(v)) == 0.0


def test_manual_trainer_writes_a_reference_format_model(tmp_path, golden):
    require_gpu()
    g = golden("manual_joseph.npz")
    X = golden("gen2_joseph.npz")["X_u8"]
    tr = ef.manual.FaceTrainer(n_components=int(g["k"]))
    tr.face_images = X
    tr.face_info = [{"face_id": i} for i in range(len(X))]
    tr.assign_labels_interactive("Joseph_Lai")
    assert tr.train_pca_model()
    path = str(tmp_path / "face_model.pkl")
    assert tr.save_model(path) and tr.save_eigenfaces(str(tmp_path), "Joseph_Lai")
    raw = open(path, "rb").read()
    assert b"__main__" in raw and b"eigenfaces_b200" not in raw
    for f in ["Joseph_Lai_mean_face.jpg", "Joseph_Lai_eigenface_01.jpg", "Joseph_Lai_eigenface_10.jpg", "Joseph_Lai_model_info.json"]:
        assert os.path.exists(tmp_path / f), f
    # the consumer arithmetic of the reference (scan-template-v2.py:228-229, :244-258) on the written model
    model = ef.manual.load_manual_pickle(path)
    crops = golden("gen2_recog.npz")
    flat = opre.preprocess_crop(crops["crop_00"], 64, 64).reshape(1, -1)
    z = (flat - model["scaler"].mean_) / model["scaler"].scale_
    f_ref = ((z - model["pca"].mean_) @ model["pca"].components_.T)[0]
    sims = np.array([np.dot(f_ref, kf) / (np.linalg.norm(f_ref) * np.linalg.norm(kf)) for kf in model["face_features"]])
    scanner = ef.manual.FaceScanner(path)
    assert scanner.load_model_and_data()
    f = scanner.extract_face_features(crops["crop_00"])
    np.testing.assert_allclose(f, f_ref, rtol=1e-9, atol=1e-9)
    pid, name, conf = scanner.recognize_face(f, 0.0)
    assert abs(conf - sims.max()) < 1e-12 and name == "Joseph_Lai"
    np.testing.assert_allclose(conf, g["recog_conf"][0], atol=1e-9)   # same confidence as the reference's own model


def test_gen1_store_fit_recognise_and_renderings(tmp_path, golden):
    """Gen-1: recognition through a pickle the reference wrote; our fit written in that format; the JPEG renderings."""
    require_gpu()
    g = golden("gen1_store.npz")
    ref = ef.gen1.load_pca_model(os.path.join(GOLDEN, "gen1_store", "toy_v1_pca_model.pkl"))
    name, sims, ok, _ = ef.gen1.recognize_faces(g["queries_u8"], ref, 0.7)
    np.testing.assert_allclose(sims, g["ref_sims"], atol=1e-12)
    assert name == "toy"
    # renderings from the reference's arrays (the fit's own eigenvector signs are LAPACK's choice in the reference)
    ef.gen1.visualize_eigenfaces(ref["eigenfaces"], ref["mean_face"], str(tmp_path), "toy_v1")
    for key in g.files:
        if key.startswith("jpg_"):
            img = cv2.imread(str(tmp_path / (key[4:] + ".jpg")), cv2.IMREAD_GRAYSCALE)
            assert img is not None and np.abs(img.astype(int) - g[key].astype(int)).max() <= 2, key
    # our own fit in the same store format
    E, mean, proj, ev = ef.gen1.manual_pca(g["X_u8"], int(g["k"]))
    path = ef.gen1.save_pca_model(E, mean, proj, ev, ref["training_filenames"], "toy", str(tmp_path), "mine")
    mine = pickle.load(open(path, "rb"))
    assert mine["eigenfaces"].flags["F_CONTIGUOUS"] and mine["eigenfaces"].shape == ref["eigenfaces"].shape
    np.testing.assert_allclose(mine["eigenvalues"], ref["eigenvalues"], rtol=1e-9)
    assert np.array_equal(mine["mean_face"], ref["mean_face"])
    _, sims2, _, _ = ef.gen1.recognize_faces(g["queries_u8"], ef.gen1.load_pca_model(path), 0.7)
    np.testing.assert_allclose(sims2, g["ref_sims"], atol=1e-9)


def test_gen1_train_single_model_from_a_directory(tmp_path, golden):
    require_gpu()
    g = golden("gen1_store.npz")
    faces = tmp_path / "faces" / "Light_version"
    os.makedirs(faces)
    for i, row in enumerate(g["X_u8"]):
        cv2.imwrite(str(faces / f"face_{i:03d}.png"), row.reshape(32, 32))
    assert ef.gen1.train_single_model(str(faces), "toy", str(tmp_path / "models"), "light", n_components=6)
    for f in ["toy_light_pca_model.pkl", "toy_light_model_info.json", "toy_light_mean_face.jpg", "toy_light_eigenface_06.jpg"]:
        assert os.path.exists(tmp_path / "models" / f), f
    assert not ef.gen1.train_single_model(str(tmp_path / "nowhere"), "toy", str(tmp_path / "models"), "dark")
    model = ef.gen1.load_pca_model(str(tmp_path / "models" / "toy_light_pca_model.pkl"))
    E_ref, mean_ref, proj_ref, ev_ref = ogen1.manual_pca(g["X_u8"].astype(np.float64), 6)
    np.testing.assert_allclose(model["eigenvalues"], ev_ref, rtol=1e-9)


class _FixedCascade:
    """Stands in for cv2.CascadeClassifier: the boxes are given (Haar detection is host work and out of scope)."""

    def __init__(self, boxes):
        self.boxes = np.asarray(boxes, dtype=np.int32)

    def detectMultiScale(self, gray, scaleFactor=1.1, minNeighbors=5, minSize=(30, 30)):
        return self.boxes


def test_gen1_dual_model_frame_loop(light_model, dark_model, golden):
    """detect_and_recognize_faces_dual_model (useless/scan.py:217-268) for all boxes of a frame against the oracle's
    per-box cv2-exact resize + recognize_face_dual_model."""
    require_gpu()
    rng = np.random.default_rng(8)
    X = golden("gen1_light.npz")["X_u8"]
    frame = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    frame[20:120, 30:130] = cv2.cvtColor(X[3].reshape(100, 100), cv2.COLOR_GRAY2BGR)
    boxes = [(30, 20, 100, 100), (200, 100, 230, 230), (400, 50, 150, 180), (10, 200, 64, 64)]
    got = ef.gen1.detect_and_recognize_faces_dual_model(frame, _FixedCascade(boxes), dark_model, light_model, 0.7)
    gray = opre.bgr_to_gray(frame)
    assert len(got) == len(boxes)
    for (x, y, w, h), r in zip(boxes, got):
        vec = opre.resize_linear_u8(gray[y:y + h, x:x + w], 100, 100).reshape(-1).astype(np.float64)
        name, best, ok, ds, ls = ogen1.recognize_face_dual_model(vec, dark_model, light_model, 0.7)
        assert r[:4] == (x, y, w, h) and r[6] == ok
        assert abs(r[5] - best) < 1e-12
        if abs(ds - ls) > 1e-9:                       # (the pasted crop is in BOTH training sets: an exact tie up to rounding)
            assert r[4] == name
    assert got[0][6] and got[0][5] > 0.99
    single = ef.gen1.detect_and_recognize_faces(frame, _FixedCascade(boxes), light_model, 0.7)
    assert [s[4] for s in single] == [light_model["person_name"]] * 4 and single[0][6]
    assert ef.gen1.detect_and_recognize_faces(frame, _FixedCascade(np.zeros((0, 4))), light_model) == []
