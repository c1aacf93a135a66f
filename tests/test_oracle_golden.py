"""The oracle against every golden vector the reference offers for this path.

CPU only.  These tests are what makes the oracle trustworthy before the CUDA path is compared to it.
"""
import numpy as np
import pytest

from oracle import extras, gen1, gen2, preprocess


# ----------------------------------------------------------------------------- preprocess (P1-P3)
def test_preprocess_matches_cv2_golden(golden):
    g = golden("preprocess.npz")
    for i, (h, w, c, dw, dh, seed) in enumerate(g["specs"]):
        shape = (h, w, 3) if c == 3 else (h, w)
        img = np.random.default_rng(int(seed)).integers(0, 256, shape, dtype=np.uint8)
        got = preprocess.preprocess_crop(img, int(dw), int(dh)).reshape(dh, dw)
        assert np.array_equal(got, g[f"out_{i:02d}"]), f"case {i} {shape}->{dw}x{dh}"


def test_preprocess_boxes_matches_cv2_golden(golden):
    g = golden("preprocess.npz")
    frame = np.random.default_rng(int(g["frame_seed"])).integers(0, 256, tuple(g["frame_shape"]), dtype=np.uint8)
    got = preprocess.preprocess_boxes(frame, g["boxes"], 100, 100)
    assert np.array_equal(got, g["roi_out_100"])


def test_preprocess_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for trial in range(300):
        h, w = int(rng.integers(1, 330)), int(rng.integers(1, 330))
        if trial % 3 == 0:
            w = h
        if trial % 25 == 0:
            h = w = 2 * (64, 100)[trial % 2]
        c = trial % 2
        img = rng.integers(0, 256, (h, w, 3) if c else (h, w), dtype=np.uint8)
        side = (64, 100)[trial % 2]
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if c else img
        assert np.array_equal(preprocess.preprocess_crop(img, side, side), cv2.resize(gray, (side, side)).flatten())


# ----------------------------------------------------------------------------- Gen-1 fit (F1)
def test_manual_pca_reproduces_shipped_light_pickle(golden, light_model):
    g = golden("gen1_light.npz")
    assert [str(s) for s in g["pkl_filenames"]] == light_model["training_filenames"]
    assert np.array_equal(light_model["mean_face"], g["pkl_mean_face"])
    np.testing.assert_allclose(light_model["eigenvalues"], g["pkl_eigenvalues"], rtol=1e-12)
    # same LAPACK as the build container reproduces signs too; elsewhere compare up to sign
    ef, ref8 = light_model["eigenfaces"], g["pkl_eigenfaces_f64_first8"]
    sign = np.sign(np.sum(ef[:, :8] * ref8, axis=0))
    np.testing.assert_allclose(ef[:, :8] * sign, ref8, atol=1e-11)
    full_sign = np.sign(np.sum(ef * g["pkl_eigenfaces_f16"].astype(np.float64), axis=0))
    np.testing.assert_allclose(ef * full_sign, g["pkl_eigenfaces_f16"].astype(np.float64), atol=2e-4)
    np.testing.assert_allclose(light_model["projected_data"] * full_sign, g["pkl_projected"], atol=1e-8)


@pytest.mark.parametrize("version", ["light", "dark"])
def test_manual_pca_reproduces_model_info_json(golden, version, light_model, dark_model):
    g = golden(f"gen1_{version}.npz")
    model = light_model if version == "light" else dark_model
    np.testing.assert_allclose(gen1.explained_variance_ratio_info(model["eigenvalues"]), g["info_evr10"], rtol=1e-12)
    assert model["n_components"] == int(g["info_n_components"])
    assert model["face_dimensions"] == int(g["info_face_dimensions"])
    assert len(model["training_filenames"]) == int(g["info_n_training_images"])
    np.testing.assert_allclose(model["eigenvalues"], g["ref_eigenvalues"], rtol=1e-12)


def test_manual_pca_covariance_branch():
    """N >= D takes the np.cov branch (useless/train.py:97-103); both branches span the same space."""
    rng = np.random.default_rng(0)
    X = np.rint(rng.normal(128, 30, (60, 24)))
    ef, mean, proj, ev = gen1.manual_pca(X, 5)
    w = np.linalg.eigvalsh(np.cov((X - X.mean(0)).T))[::-1][:5]
    np.testing.assert_allclose(ev, w, rtol=1e-12)
    np.testing.assert_allclose(ef.T @ ef, np.eye(5), atol=1e-12)
    np.testing.assert_allclose(proj, (X - mean) @ ef, atol=1e-10)


# ----------------------------------------------------------------------------- Gen-1 recognition (J2, M2)
def test_gen1_recognition_matches_reference_outputs(golden, light_model, dark_model):
    g = golden("gen1_recog.npz")
    Q = g["queries_u8"]
    thr = float(g["threshold"])
    sign = np.sign(np.sum(light_model["projected_data"] * golden("gen1_light.npz")["pkl_projected"], axis=0))
    for i, row in enumerate(Q):
        v = row.astype(np.float64)
        p = gen1.project_face_to_eigenspace(v, light_model["eigenfaces"], light_model["mean_face"])
        np.testing.assert_allclose(p * sign, g["ref_proj_light"][i], atol=1e-8)
        _, sl, _ = gen1.recognize_face(v, light_model, thr)
        _, sd, _ = gen1.recognize_face(v, dark_model, thr)
        assert abs(sl - g["ref_sim_light"][i]) < 1e-12 and abs(sd - g["ref_sim_dark"][i]) < 1e-12
        name, best, rec, _, _ = gen1.recognize_face_dual_model(v, dark_model, light_model, thr)
        if abs(g["ref_sim_light"][i] - g["ref_sim_dark"][i]) > 1e-12:   # exact-duplicate crops tie at rounding level
            assert (name == "Joseph_Lai_dark") == bool(g["ref_dual_name_is_dark"][i])
        assert bool(rec) == bool(g["ref_dual_recognized"][i])
        assert abs(best - g["ref_dual_best"][i]) < 1e-12


def test_gen1_batched_equals_literal(golden, light_model):
    Q = golden("gen1_recog.npz")["queries_u8"]
    best, idx, rec = gen1.recognize_batch(Q, light_model, 0.8)
    for i, row in enumerate(Q):
        _, s, r = gen1.recognize_face(row.astype(np.float64), light_model, 0.8)
        assert abs(s - best[i]) < 1e-13 and bool(r) == bool(rec[i])


# ----------------------------------------------------------------------------- Gen-2 fit (F2)
def test_gen2_fit_matches_reference_train_v5(golden):
    g = golden("gen2_joseph.npz")
    X = g["X_u8"]
    assert X.shape == (int(g["info_total_faces"]), 4096) and int(g["info_n_components"]) == 178
    assert str(g["ref_solver"]) == "full"
    fit = gen2.train_pca_model(X, 178)
    np.testing.assert_allclose(fit["mean_face"], g["ref_mean_face"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(fit["scaler_mean"], g["ref_scaler_mean"], rtol=1e-14)
    np.testing.assert_allclose(fit["scaler_var"], g["ref_scaler_var"], rtol=1e-12)
    np.testing.assert_allclose(fit["scaler_scale"], g["ref_scaler_scale"], rtol=1e-12)
    np.testing.assert_allclose(fit["singular_values"][:170], g["ref_singular_values"][:170], rtol=1e-10)
    np.testing.assert_allclose(fit["explained_variance_ratio"][:170], g["ref_explained_variance_ratio"][:170], rtol=1e-10)
    assert abs(fit["explained_variance_ratio"].sum() - float(g["info_evr_sum"])) < 1e-9
    np.testing.assert_allclose(fit["eigenfaces"][:10], g["ref_components_first10"], atol=1e-10)
    np.testing.assert_allclose(fit["face_features"][:, :20], g["ref_face_features_first20"], atol=1e-8)
    assert float(g["ref_noise_variance"]) == 0.0 and fit["noise_variance"] == 0.0


def test_gen2_scaler_against_sklearn():
    sk = pytest.importorskip("sklearn.preprocessing")
    rng = np.random.default_rng(3)
    X = rng.integers(0, 256, (50, 300), dtype=np.uint8)
    X[:, 7] = 13                    # constant feature -> scale 1
    X[:, 8] = 0
    s = sk.StandardScaler().fit(X)
    mean, var, scale = gen2.scaler_fit(X)
    np.testing.assert_allclose(mean, s.mean_, rtol=1e-15)
    np.testing.assert_allclose(var, s.var_, rtol=1e-13, atol=1e-13)
    np.testing.assert_array_equal(scale == 1.0, s.scale_ == 1.0)
    np.testing.assert_allclose(scale, s.scale_, rtol=1e-13)
    np.testing.assert_array_equal(gen2.scaler_transform(X, mean, scale), s.transform(X))


# ----------------------------------------------------------------------------- Gen-2 recognition (N1, J1, M1, M3)
def _gen2_models(g):
    models = {}
    for person in [str(p) for p in g["persons"]]:
        models[person] = dict(
            scaler_mean=g[f"{person}_scaler_mean"], scaler_scale=g[f"{person}_scaler_scale"],
            components=g[f"{person}_components"], pca_mean=g[f"{person}_pca_mean"],
            face_features=g[f"{person}_face_features"], face_labels=g[f"{person}_face_labels"],
            person_id_map={person: 0})
    return models


def test_gen2_recognition_matches_reference_outputs(golden):
    g = golden("gen2_recog.npz")
    models = _gen2_models(g)
    for i in range(int(g["n_crops"])):
        flat = preprocess.preprocess_crop(g[f"crop_{i:02d}"], 64, 64)
        for j, (person, m) in enumerate(models.items()):
            f = gen2.extract_features(flat, m["scaler_mean"], m["scaler_scale"], m["components"], m["pca_mean"])[0]
            np.testing.assert_allclose(f, g["ref_features"][i, j], rtol=1e-10, atol=1e-9)
            pid, name, sim = gen2.recognize_with_model(f, m["face_features"], m["face_labels"], m["person_id_map"], 0.7)
            assert int(pid) == int(g["ref_single_pid"][i, j]) and name == str(g["ref_single_name"][i, j])
            assert abs(sim - g["ref_single_sim"][i, j]) < 1e-12
        pid, name, conf = gen2.recognize_all_models(flat, models, 0.8)
        assert int(pid) == int(g["ref_multi_pid"][i]) and name == str(g["ref_multi_name"][i])
        assert abs(conf - g["ref_multi_conf"][i]) < 1e-12


def test_gen2_shipped_pickle_recognition(golden):
    """The one shipped Gen-2 pickle (77 faces, k=76, float32 arrays): labels and argmax from the oracle equal the
    reference's arithmetic on it; stored face_features are reproduced to JPEG-decoder drift."""
    g = golden("gen2_shipped.npz")
    m = dict(scaler_mean=g["scaler_mean"], scaler_scale=g["scaler_scale"], components=g["components"],
             pca_mean=g["pca_mean"], face_features=g["face_features"], face_labels=g["face_labels"])
    best, idx, labels = gen2.recognize_batch(g["X_u8"], m, 0.7)
    assert np.array_equal(idx, g["ref_argmax"])
    assert np.array_equal(labels, g["ref_pid"])
    np.testing.assert_allclose(best, g["ref_sim"], atol=1e-12)
    feats = gen2.extract_features(g["X_u8"], m["scaler_mean"], m["scaler_scale"], m["components"], m["pca_mean"])
    np.testing.assert_allclose(feats, g["ref_features"], rtol=1e-9, atol=1e-7)
    rel = np.abs(feats - g["face_features"]).max() / np.abs(g["face_features"]).max()
    assert rel < 1e-4


def test_gen2_batched_equals_literal(golden):
    g = golden("gen2_recog.npz")
    models = _gen2_models(g)
    flats = np.stack([preprocess.preprocess_crop(g[f"crop_{i:02d}"], 64, 64) for i in range(int(g["n_crops"]))])
    for person, m in models.items():
        best, idx, labels = gen2.recognize_batch(flats, m, 0.7)
        for i, flat in enumerate(flats):
            f = gen2.extract_features(flat, m["scaler_mean"], m["scaler_scale"], m["components"], m["pca_mean"])[0]
            pid, _, sim = gen2.recognize_with_model(f, m["face_features"], m["face_labels"], m["person_id_map"], 0.7)
            assert int(pid) == int(labels[i]) and abs(sim - best[i]) < 1e-13


def test_manual_classes_variant():
    """scripts/manual ManualPCA (D x D covariance) spans the same leading subspace as the SVD route."""
    rng = np.random.default_rng(9)
    X = rng.integers(0, 256, (40, 64), dtype=np.uint8)
    mean, std = gen2.manual_scaler_fit(X)
    Z = (X - mean) / std
    a = gen2.manual_pca_fit(Z, 5)
    b = gen2.pca_fit_full(Z, 5)
    np.testing.assert_allclose(np.abs(a["components"] @ b["components"].T), np.eye(5), atol=1e-8)
    np.testing.assert_allclose(a["explained_variance"], b["explained_variance"], rtol=1e-10)


# ----------------------------------------------------------------------------- extras (X1, unpinned)
def test_extras_identities(light_model):
    rng = np.random.default_rng(11)
    x = rng.integers(0, 256, (5, 10000)).astype(np.float64)
    E, mu = light_model["eigenfaces"], light_model["mean_face"]
    v = x - mu
    p = v @ E
    err = extras.reconstruction_error2(v, E)
    np.testing.assert_allclose(err, np.einsum('ij,ij->i', v, v) - np.einsum('ij,ij->i', p, p), rtol=1e-9)
    d2, idx = extras.l2_nearest(p, light_model["projected_data"])
    G = light_model["projected_data"]
    full = ((p[:, None, :] - G[None, :, :]) ** 2).sum(-1)
    assert np.array_equal(idx, full.argmin(1))
