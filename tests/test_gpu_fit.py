"""PCA-fit parity through the C ABI: Gen-1 manual_pca and Gen-2 StandardScaler + PCA(full) against the reference's
shipped artefacts (goldens) and the oracle.

Bars: eigenvalues rtol 1e-9; eigenfaces compared after sign normalisation, principal angle of the retained
subspace < 1e-6 rad and per-component |cos| > 1 - 1e-10 where the spectrum is separated; projections atol 1e-6
(scale 3.5e3, i.e. 3e-10 relative).
"""
import numpy as np
import pytest

import eigenfaces_b200 as ef
from gpu_util import require_gpu
from oracle import gen1, gen2

pytestmark = pytest.mark.gpu


def _principal_angle(A, B):
    """Largest principal angle between the column spaces of two orthonormal bases."""
    s = np.linalg.svd(A.T @ B, compute_uv=False)
    return float(np.arccos(np.clip(s.min(), -1.0, 1.0)))


@pytest.mark.parametrize("version", ["light", "dark"])
def test_gen1_fit_matches_shipped_artifacts(golden, version, capsys):
    require_gpu()
    g = golden(f"gen1_{version}.npz")
    X = g["X_u8"]
    ef_, mean, proj, ev = ef.gen1.manual_pca(X.astype(np.float64), 50)
    assert ef_.shape == (10000, 50) and ef_.flags["F_CONTIGUOUS"] and proj.shape == (len(X), 50)
    assert np.array_equal(mean, g["ref_mean_face"])                       # exact integer sum / N
    np.testing.assert_allclose(ev, g["ref_eigenvalues"], rtol=1e-9)
    np.testing.assert_allclose(gen1.explained_variance_ratio_info(ev), g["info_evr10"], rtol=1e-9)   # *_model_info.json
    np.testing.assert_allclose(ef_.T @ ef_, np.eye(50), atol=1e-10)
    ref8 = g["ref_eigenfaces_f64_first8"]
    cosines = np.abs(np.sum(ef_[:, :8] * ref8, axis=0))
    assert cosines.min() > 1 - 1e-10
    sign = np.sign(np.sum(proj * g["ref_projected"], axis=0))
    np.testing.assert_allclose(proj * sign, g["ref_projected"], atol=1e-6)
    if version == "light":
        np.testing.assert_allclose(proj * sign, g["pkl_projected"], atol=1e-6)          # the shipped pickle itself
        np.testing.assert_allclose(ev, g["pkl_eigenvalues"], rtol=1e-9)
        full = g["pkl_eigenfaces_f16"].astype(np.float64)
        np.testing.assert_allclose(ef_ * np.sign(np.sum(ef_ * full, axis=0)), full, atol=2e-4)
    # subspace angle against the oracle's full basis
    E_ref, _, _, _ = gen1.manual_pca(X.astype(np.float64), 50)
    assert _principal_angle(E_ref[:, :40], ef_[:, :40]) < 1e-6


def test_gen1_fit_covariance_branch_and_defaults():
    require_gpu()
    rng = np.random.default_rng(0)
    X = rng.integers(0, 256, (300, 96), dtype=np.uint8)      # N >= D -> D x D covariance (useless/train.py:97-103)
    E, mean, proj, ev, info = ef.fit_gen1(X, 12)
    assert info["branch"] == 1
    E_ref, m_ref, p_ref, ev_ref = gen1.manual_pca(X.astype(np.float64), 12)
    np.testing.assert_allclose(ev, ev_ref, rtol=1e-9)
    assert np.abs(np.abs(np.sum(E * E_ref, axis=0)) - 1).max() < 1e-9
    sign = np.sign(np.sum(E * E_ref, axis=0))
    np.testing.assert_allclose(proj * sign, p_ref, atol=1e-8)
    # n_components=None -> min(N-1, D) (:111-112); k clamps to the number of eigenvalues (:114)
    Xs = rng.integers(0, 256, (20, 400), dtype=np.uint8)
    E2, _, p2, ev2, info2 = ef.fit_gen1(Xs, None)
    assert info2["branch"] == 0 and E2.shape == (400, 19) and p2.shape == (20, 19)
    E3, _, _, ev3, _ = ef.fit_gen1(Xs, 500)
    assert E3.shape == (400, 20)
    np.testing.assert_allclose(ev3[:19], gen1.manual_pca(Xs.astype(np.float64), 500)[3][:19], rtol=1e-9)
    with pytest.raises(ValueError):
        ef.fit_gen1(rng.normal(size=(10, 50)), 3)           # non-pixel data is refused, not silently mis-handled


def test_gen2_fit_matches_reference_train_v5(golden, tmp_path):
    require_gpu()
    g = golden("gen2_joseph.npz")
    X = g["X_u8"]
    tr = ef.gen2.MultiFaceTrainer(n_components=178)
    tr.face_images, tr.face_labels = X, np.zeros(len(X), dtype=int)
    tr.person_id_map = {"Joseph_Lai": 0}
    assert tr.train_pca_model() is True
    np.testing.assert_allclose(tr.mean_face, g["ref_mean_face"], atol=1e-12)
    np.testing.assert_allclose(tr.scaler.mean_, g["ref_scaler_mean"], rtol=1e-14)
    np.testing.assert_allclose(tr.scaler.var_, g["ref_scaler_var"], rtol=1e-11)
    np.testing.assert_allclose(tr.scaler.scale_, g["ref_scaler_scale"], rtol=1e-11)
    np.testing.assert_allclose(tr.pca.singular_values_[:170], g["ref_singular_values"][:170], rtol=1e-8)
    np.testing.assert_allclose(tr.pca.explained_variance_ratio_[:170], g["ref_explained_variance_ratio"][:170], rtol=1e-8)
    assert abs(tr.pca.explained_variance_ratio_.sum() - float(g["info_evr_sum"])) < 1e-9      # model_info.json scalar
    assert tr.pca.noise_variance_ == 0.0 and tr.pca.n_components_ == int(g["info_n_components"])
    # svd_flip sign convention reproduced -> components and features comparable without sign fixing
    np.testing.assert_allclose(tr.eigenfaces[:10], g["ref_components_first10"], atol=1e-8)
    np.testing.assert_allclose(tr.face_features[:, :20], g["ref_face_features_first20"], atol=1e-6)
    good = tr.pca.singular_values_ / tr.pca.singular_values_[0] > 1e-6
    C = tr.eigenfaces[good]
    np.testing.assert_allclose(C @ C.T, np.eye(good.sum()), atol=1e-9)
    # the pickle round-trips through genuine sklearn objects (what scan-template-v4.py calls .transform on)
    path = str(tmp_path / "face_model.pkl")
    assert tr.save_model(path)
    import pickle
    md = pickle.load(open(path, "rb"))
    assert sorted(md) == sorted(['pca', 'scaler', 'face_features', 'face_labels', 'face_info', 'person_id_map',
                                 'n_components', 'mean_face', 'eigenfaces', 'face_shape', 'training_date'])
    feats = md["pca"].transform(md["scaler"].transform(X[:5]))
    np.testing.assert_allclose(feats[:, :20], g["ref_face_features_first20"][:5], atol=1e-6)
    tr2 = ef.gen2.MultiFaceTrainer()
    assert tr2.load_model(path) and tr2.n_components == 178


def test_gen2_fit_truncated_k_and_tall_branch():
    require_gpu()
    rng = np.random.default_rng(4)
    base = rng.normal(0, 1, (120, 12)) @ rng.normal(0, 1, (12, 256))
    X = np.clip(np.rint(128 + 20 * base + rng.normal(0, 3, (120, 256))), 0, 255).astype(np.uint8)
    out = ef.fit_gen2(X, 10)
    ref = gen2.train_pca_model(X, 10)
    np.testing.assert_allclose(out["singular_values"], ref["singular_values"], rtol=1e-9)
    np.testing.assert_allclose(out["components"], ref["eigenfaces"], atol=1e-8)
    np.testing.assert_allclose(out["features"], ref["face_features"], atol=1e-7)
    np.testing.assert_allclose(out["noise_variance"], ref["noise_variance"], rtol=1e-9)
    np.testing.assert_allclose(out["explained_variance_ratio"], ref["explained_variance_ratio"], rtol=1e-9)
    Xt = np.clip(np.rint(128 + 25 * rng.normal(0, 1, (500, 8)) @ rng.normal(0, 1, (8, 64)) + rng.normal(0, 2, (500, 64))), 0, 255).astype(np.uint8)
    out = ef.fit_gen2(Xt, 6)                     # N > D: covariance side
    ref = gen2.train_pca_model(Xt, 6)
    assert out["info"]["branch"] == 1
    np.testing.assert_allclose(out["singular_values"], ref["singular_values"], rtol=1e-9)
    np.testing.assert_allclose(out["components"], ref["eigenfaces"], atol=1e-8)
    np.testing.assert_allclose(out["features"], ref["face_features"], atol=1e-7)


def test_trained_model_recognises_its_own_faces(golden):
    """BASELINE config 1: fit on the Joseph_Lai crops, then recognise the same crops (self match, label 0)."""
    require_gpu()
    X = golden("gen2_joseph.npz")["X_u8"]
    tr = ef.gen2.MultiFaceTrainer(n_components=10)
    tr.face_images, tr.face_labels, tr.person_id_map = X, np.zeros(len(X), dtype=int), {"Joseph_Lai": 0}
    assert tr.train_pca_model()
    md = dict(pca=tr.pca, scaler=tr.scaler, face_features=tr.face_features, face_labels=tr.face_labels,
              person_id_map=tr.person_id_map)
    res = ef.gen2.recognizer_for(md).recognize(X, 0.7)
    assert (res.label == 0).all() and (res.score > 1 - 1e-9).all()
    ref = gen2.train_pca_model(X, 10)
    m = dict(scaler_mean=ref["scaler_mean"], scaler_scale=ref["scaler_scale"], components=ref["eigenfaces"],
             pca_mean=ref["pca_mean"], face_features=ref["face_features"], face_labels=np.zeros(len(X), int))
    _, idx, _ = gen2.recognize_batch(X, m, 0.7)
    uniq = np.unique(X, axis=0, return_index=True)[1]
    assert np.array_equal(res.index[uniq], idx[uniq])


def test_subspace_eigensolver_matches_numpy_eigh():
    """Chebyshev-filtered subspace iteration (the config-4 solver: top-k of a covariance too large for Jacobi) against
    numpy eigh on a planted spectrum; eigenvalues 1e-9 rel, subspace by principal angles, orthonormal columns."""
    torch = require_gpu()
    rng = np.random.default_rng(4242)
    n, k = 1500, 40
    F = np.linalg.qr(rng.normal(size=(n, 300)))[0]
    sig = 40.0 * np.arange(1, 301) ** -0.7
    Cm = (F * sig ** 2) @ F.T + 16.0 * np.eye(n) + 1e-3 * np.diag(rng.random(n))
    Cm = (Cm + Cm.T) / 2
    w, V = np.linalg.eigh(Cm)
    w, V = w[::-1], V[:, ::-1]
    lam, Q, info = ef.dist.eigh_topk_device(torch.from_numpy(Cm).cuda(), k)
    lam, Q = lam.cpu().numpy(), Q.cpu().numpy()
    np.testing.assert_allclose(lam, w[:k], rtol=1e-9)
    np.testing.assert_allclose(Q.T @ Q, np.eye(k), atol=1e-10)
    # principal angles between span(Q) and span(V[:, :k]): cosines = singular values of V_k^T Q
    cosines = np.linalg.svd(V[:, :k].T @ Q, compute_uv=False)
    assert np.sqrt(np.maximum(0.0, 1.0 - cosines.min() ** 2)) < 1e-6, (cosines.min(), info)
    assert info["residual"] <= 1e-11 and info["outer"] < 60, info


def test_row_sharded_fit_subspace_branch_matches_oracle():
    """fit_gen1_sharded with the subspace solver (what D = 10 000 uses) on a size the oracle finishes quickly."""
    torch = require_gpu()
    rng = np.random.default_rng(7)
    N, D, k = 2500, 640, 16
    X = np.clip(np.rint(128 + 25 * rng.normal(size=(N, 24)) @ rng.normal(size=(24, D)) / np.sqrt(24) * np.linspace(2, 0.3, D)
                        + rng.normal(0, 3, (N, D))), 0, 255).astype(np.uint8)
    E, mean, proj, ev = ef.dist.fit_gen1_sharded(torch.from_numpy(X).cuda(), N, k, solver="subspace")
    E_ref, mean_ref, proj_ref, ev_ref = gen1.manual_pca(X.astype(np.float64), k)
    np.testing.assert_allclose(ev.cpu().numpy(), ev_ref, rtol=1e-9)
    assert np.array_equal(mean.cpu().numpy(), mean_ref)
    sign = np.sign(np.sum(E.cpu().numpy() * E_ref, axis=0))
    gaps = np.abs(np.diff(ev_ref)) / ev_ref[:-1]
    well = np.concatenate([[True], gaps > 1e-3]) & np.concatenate([gaps > 1e-3, [True]])   # well-separated components
    np.testing.assert_allclose((E.cpu().numpy() * sign)[:, well], E_ref[:, well], atol=1e-6)
    np.testing.assert_allclose((proj.cpu().numpy() * sign)[:, well], proj_ref[:, well], atol=1e-4)
