"""K2 parity: projection + nearest-gallery match through the C ABI against the oracle and the reference goldens.

Bars (BASELINE.json north_star): identity labels / arg-best indices bit exact; projections and scores within a
stated float tolerance (float64-equivalent path, S = 8 digit planes): |dp| <= 1e-9 * max(1, |p|), |dscore| <= 1e-12.
"""
import numpy as np
import pytest

import eigenfaces_b200 as ef
from gpu_util import face_like, require_gpu
from oracle import extras, gen1, gen2, preprocess

pytestmark = pytest.mark.gpu

PROJ_RTOL = 1e-9
SCORE_ATOL = 1e-12


def _close_proj(got, want):
    np.testing.assert_allclose(got, want, rtol=PROJ_RTOL, atol=PROJ_RTOL)


# ------------------------------------------------------------------------------------------------ Gen-1
def test_gen1_reference_golden(golden, light_model, dark_model):
    require_gpu()
    g = golden("gen1_recog.npz")
    Q, thr = g["queries_u8"], float(g["threshold"])
    sign = np.sign(np.sum(light_model["projected_data"] * golden("gen1_light.npz")["pkl_projected"], axis=0))
    _, sl, rl, res = ef.gen1.recognize_faces(Q, light_model, thr)
    _, sd, rd, _ = ef.gen1.recognize_faces(Q, dark_model, thr)
    np.testing.assert_allclose(res.features * sign, g["ref_proj_light"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(sl, g["ref_sim_light"], atol=SCORE_ATOL)
    np.testing.assert_allclose(sd, g["ref_sim_dark"], atol=SCORE_ATOL)
    names, best, rec, _, _ = ef.gen1.recognize_faces_dual_model(Q, dark_model, light_model, thr)
    np.testing.assert_allclose(best, g["ref_dual_best"], atol=SCORE_ATOL)
    assert np.array_equal(rec, g["ref_dual_recognized"])
    decided = np.abs(g["ref_sim_light"] - g["ref_sim_dark"]) > 1e-12
    assert np.array_equal((names == "Joseph_Lai_dark")[decided], g["ref_dual_name_is_dark"][decided])
    # single-crop signatures of the reference
    name, s, ok = ef.gen1.recognize_face(Q[17].astype(np.float64), light_model, thr)
    assert name == "Joseph_Lai" and abs(s - g["ref_sim_light"][17]) < SCORE_ATOL and ok == bool(s >= thr)
    p = ef.gen1.project_face_to_eigenspace(Q[3].astype(np.float64), light_model["eigenfaces"], light_model["mean_face"])
    _close_proj(p, gen1.project_face_to_eigenspace(Q[3].astype(np.float64), light_model["eigenfaces"], light_model["mean_face"]))


@pytest.mark.parametrize("use_tc", [0, 1, 2])
def test_gen1_batch_vs_oracle(golden, light_model, use_tc):
    require_gpu()
    X = golden("gen1_light.npz")["X_u8"]
    rng = np.random.default_rng(1)
    Q = np.concatenate([X, face_like(rng, X, 300), rng.integers(0, 256, (55, 10000), dtype=np.uint8)])
    rec = ef.gen1.recognizer_for(light_model)
    rec.use_tensor_cores(use_tc)
    res = rec.recognize(Q, 0.8)
    rec.use_tensor_cores(2)
    want_p = gen1.project_batch(Q, light_model["eigenfaces"], light_model["mean_face"])
    _close_proj(res.features, want_p)
    best, idx, ok = gen1.recognize_batch(Q, light_model, 0.8)
    np.testing.assert_allclose(res.score, best, atol=SCORE_ATOL)
    sims = gen1.cosine_matrix(want_p, light_model["projected_data"])
    # arg-best: identical wherever the reference's own top-2 margin is above its rounding noise
    srt = np.sort(sims, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-13
    assert decided.mean() > 0.5
    assert np.array_equal(res.index[decided], idx[decided])
    assert np.array_equal(res.label >= 0, ok)
    # residual: distance from face space, textbook form
    want_r = extras.reconstruction_error2(Q.astype(np.float64) - light_model["mean_face"], light_model["eigenfaces"])
    np.testing.assert_allclose(res.resid2, want_r, rtol=1e-9, atol=1e-6)


def test_gen1_self_match_duplicates_lowest_index(light_model, golden):
    """Every training crop returns its own gallery row, except exact duplicates where the LOWEST row wins
    (np.argmax rule; SURVEY.md section 8c golden 6)."""
    require_gpu()
    Q = golden("gen1_light.npz")["X_u8"]
    rng = np.random.default_rng(2)
    G = light_model["projected_data"].copy()
    dup_src = rng.choice(len(G), 20, replace=False)
    dup_dst = rng.choice(np.setdiff1d(np.arange(len(G)), dup_src), 20, replace=False)
    G[dup_dst] = G[dup_src]
    rec = ef.Recognizer(light_model["eigenfaces"], light_model["mean_face"], G, metric=ef.METRIC_COSINE_G1)
    res = rec.recognize(Q, 0.0)
    sims = gen1.cosine_matrix(gen1.project_batch(Q, light_model["eigenfaces"], light_model["mean_face"]), G)
    want = np.argmax(sims, axis=1)
    srt = np.sort(sims, axis=1)
    clear = (srt[:, -1] - srt[:, -2]) > 1e-13
    assert np.array_equal(res.index[clear], want[clear])
    # expected row for a training crop = lowest index among the gallery rows identical to its own row
    # (the shipped set itself contains exact duplicate crops, e.g. rows 74 and 131)
    first_of = np.array([np.flatnonzero((G == G[i]).all(axis=1))[0] for i in range(len(G))])
    for s, d in zip(dup_src, dup_dst):
        assert res.index[s] == first_of[s] <= min(s, d)
        assert not (res.index == max(s, d)).any(), "a duplicated gallery row must never beat its lower-index twin"
    untouched = np.setdiff1d(np.arange(len(G)), dup_dst)
    assert np.array_equal(res.index[untouched], first_of[untouched])


def test_gen1_slices_tolerance(light_model, golden):
    """Fewer digit planes trade accuracy for tensor work: S=4 is float32-class, S=8 float64-class."""
    require_gpu()
    X = golden("gen1_light.npz")["X_u8"][:64]
    want = gen1.project_batch(X, light_model["eigenfaces"], light_model["mean_face"])
    for S, tol in ((3, 2e-2), (4, 2e-4), (6, 1e-8), (8, 1e-9)):
        rec = ef.Recognizer(light_model["eigenfaces"], light_model["mean_face"], light_model["projected_data"],
                            metric=ef.METRIC_COSINE_G1, n_slices=S)
        got = rec.recognize(X, 0.8).features
        assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max() / 1e3), f"S={S}"


def test_l2_metric_and_edge_batches(light_model, golden):
    require_gpu()
    X = golden("gen1_light.npz")["X_u8"]
    rec = ef.Recognizer(light_model["eigenfaces"], light_model["mean_face"], light_model["projected_data"],
                        metric=ef.METRIC_L2)
    rng = np.random.default_rng(3)
    Q = face_like(rng, X, 130, noise=12.0)
    res = rec.recognize(Q, 1e12)
    p = gen1.project_batch(Q, light_model["eigenfaces"], light_model["mean_face"])
    d2, idx = extras.l2_nearest(p, light_model["projected_data"])
    assert np.array_equal(res.index, idx)
    np.testing.assert_allclose(res.score, d2, rtol=1e-9, atol=1e-6)
    # ragged / tiny batches
    for B in (0, 1, 31, 33, 129):
        r = rec.recognize(Q[:B], 1e12)
        assert r.index.shape == (B,) and np.array_equal(r.index, idx[:B])
    # zero query under the Gen-1 cosine rule: similarity 0.0 (useless/scan.py:73-74)
    rec_c = ef.gen1.recognizer_for(light_model)
    mean_img = np.clip(np.rint(light_model["mean_face"]), 0, 255).astype(np.uint8)[None]
    r = rec_c.recognize(mean_img, 0.8)
    assert r.score[0] <= 1.0 and r.label[0] == -1 or r.score[0] >= 0.8


# ------------------------------------------------------------------------------------------------ Gen-2
def _gen2_model_dict(g, person):
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    pca = PCA(n_components=len(g[f"{person}_components"]))
    pca.components_, pca.mean_ = g[f"{person}_components"], g[f"{person}_pca_mean"]
    pca.explained_variance_ = g[f"{person}_explained_variance"]
    pca.whiten = False
    sc = StandardScaler()
    sc.mean_, sc.scale_, sc.var_ = g[f"{person}_scaler_mean"], g[f"{person}_scaler_scale"], g[f"{person}_scaler_var"]
    return dict(pca=pca, scaler=sc, face_features=g[f"{person}_face_features"],
                face_labels=g[f"{person}_face_labels"], person_id_map={person: 0}, n_components=20,
                mean_face=g[f"{person}_mean_face"], eigenfaces=pca.components_, face_shape=(64, 64))


def test_gen2_reference_golden(golden):
    require_gpu()
    g = golden("gen2_recog.npz")
    persons = [str(p) for p in g["persons"]]
    scanner = ef.gen2.MultiModelFaceScanner()
    for p in persons:
        scanner.models[p] = {"model_data": _gen2_model_dict(g, p)}
    for i in range(int(g["n_crops"])):
        crop = g[f"crop_{i:02d}"]
        for j, p in enumerate(persons):
            md = scanner.models[p]["model_data"]
            f = scanner.extract_face_features(crop, md)
            np.testing.assert_allclose(f, g["ref_features"][i, j], rtol=1e-9, atol=1e-8)
            pid, name, sim = scanner.recognize_face_with_model(f, md, 0.7)
            assert int(pid) == int(g["ref_single_pid"][i, j]) and name == str(g["ref_single_name"][i, j])
            assert abs(sim - g["ref_single_sim"][i, j]) < SCORE_ATOL
        pid, name, conf = scanner.recognize_face_all_models(crop, 0.8)
        assert int(pid) == int(g["ref_multi_pid"][i]) and name == str(g["ref_multi_name"][i])
        assert abs(conf - g["ref_multi_conf"][i]) < SCORE_ATOL


def test_gen2_shipped_pickle_labels_bit_exact(golden):
    """The reference's one shipped Gen-2 pickle (float32 arrays, k=76, 77 faces): labels and arg-max rows equal the
    reference's arithmetic exactly; features within tolerance."""
    require_gpu()
    g = golden("gen2_shipped.npz")
    rec = ef.Recognizer(g["components"], g["scaler_mean"], g["face_features"], scale=g["scaler_scale"],
                        pca_mean=g["pca_mean"], labels=g["face_labels"], metric=ef.METRIC_COSINE_SK,
                        basis_is_components=True)
    res = rec.recognize(g["X_u8"], 0.7)
    assert np.array_equal(res.index, g["ref_argmax"])
    assert np.array_equal(res.label, g["ref_pid"])
    np.testing.assert_allclose(res.score, g["ref_sim"], atol=SCORE_ATOL)
    np.testing.assert_allclose(res.features, g["ref_features"], rtol=1e-9, atol=1e-7)
    m = dict(scaler_mean=g["scaler_mean"], scaler_scale=g["scaler_scale"], components=g["components"].astype(np.float64),
             pca_mean=g["pca_mean"].astype(np.float64))
    z = gen2.scaler_transform(g["X_u8"], m["scaler_mean"], m["scaler_scale"]) - m["pca_mean"]
    want_r = extras.reconstruction_error2(z, m["components"].T)
    # |z|^2 - |p|^2 equals the direct form only up to the orthonormality defect of the basis; this pickle stores
    # float32 components (defect ~1e-7), so the two textbook forms differ by ~1e-7 |z|^2 (unpinned extra, X1).
    E = m["components"]
    defect = np.abs(E @ E.T - np.eye(len(E))).max()
    znorm2 = np.einsum('ij,ij->i', z, z)
    assert 1e-9 < defect < 1e-5
    assert np.all(np.abs(res.resid2 - want_r) <= 4 * len(E) * defect * znorm2)


def test_gen2_full_k_model_vs_oracle(golden):
    """train-v5 sets k = N (178): wide basis (NC = 8 * 179 digit-plane columns) through several N tiles."""
    require_gpu()
    X = golden("gen2_joseph.npz")["X_u8"]
    fit = gen2.train_pca_model(X, 178)
    rec = ef.Recognizer(fit["eigenfaces"], fit["scaler_mean"], fit["face_features"], scale=fit["scaler_scale"],
                        pca_mean=fit["pca_mean"], labels=np.zeros(178, np.int32), metric=ef.METRIC_COSINE_SK,
                        basis_is_components=True)
    rng = np.random.default_rng(8)
    Q = np.concatenate([X, face_like(rng, X, 100)])
    res = rec.recognize(Q, 0.7)
    m = dict(scaler_mean=fit["scaler_mean"], scaler_scale=fit["scaler_scale"], components=fit["eigenfaces"],
             pca_mean=fit["pca_mean"], face_features=fit["face_features"], face_labels=np.zeros(178, int))
    want = gen2.extract_features(Q, m["scaler_mean"], m["scaler_scale"], m["components"], m["pca_mean"])
    np.testing.assert_allclose(res.features, want, rtol=1e-9, atol=1e-8)
    best, idx, labels = gen2.recognize_batch(Q, m, 0.7)
    np.testing.assert_allclose(res.score, best, atol=SCORE_ATOL)
    sims = gen2.sk_cosine_similarity(want, m["face_features"])
    srt = np.sort(sims, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-13
    assert np.array_equal(res.index[decided], idx[decided])
    assert np.array_equal(res.label, labels)
    assert np.array_equal(res.index[:178][decided[:178]], np.arange(178)[decided[:178]])   # self match


def test_all_three_kernel_paths_agree_bit_for_bit(light_model, golden):
    """dp4a (CUDA cores), tcgen05 stream-K + epilogue kernels, and the single cluster kernel compute the same integers,
    hence the same float64 features, scores, indices, labels and residuals."""
    require_gpu()
    X = golden("gen1_light.npz")["X_u8"]
    rng = np.random.default_rng(12)
    Q = np.concatenate([X, face_like(rng, X, 411)])            # 640 crops: 5 full tiles
    Q = np.concatenate([Q, Q[:37]])                            # + a ragged tail
    rec = ef.Recognizer(light_model["eigenfaces"][:, :10].copy(order="F"), light_model["mean_face"],
                        light_model["projected_data"][:, :10].copy(), metric=ef.METRIC_COSINE_G1)
    outs = []
    for mode in (0, 1, 2):
        rec.use_tensor_cores(mode)
        outs.append(rec.recognize(Q, 0.8))
        assert rec.pipeline_timeouts() == 0
    for o in outs[1:]:
        for f in ("features", "score", "index", "label", "resid2"):
            assert np.array_equal(getattr(o, f), getattr(outs[0], f)), f


def test_device_buffers_match_host_buffers(light_model, golden):
    torch = require_gpu()
    X = golden("gen1_light.npz")["X_u8"][:100]
    rec = ef.gen1.recognizer_for(light_model)
    host = rec.recognize(X, 0.8)
    xd = torch.zeros((100, 10112), dtype=torch.uint8, device="cuda")
    xd[:, :10000] = torch.from_numpy(X).cuda()
    out = rec.recognize_device(xd, 0.8)
    torch.cuda.synchronize()
    assert np.array_equal(out["features"].cpu().numpy(), host.features)
    assert np.array_equal(out["score"].cpu().numpy(), host.score)
    assert np.array_equal(out["index"].cpu().numpy(), host.index)
    assert np.array_equal(out["label"].cpu().numpy(), host.label)
    assert np.array_equal(out["resid2"].cpu().numpy(), host.resid2)


def test_baseline_config2_full_size_properties(light_model):
    """BASELINE config 2 at full size (B=4096, D=10000, k=10, Ng=1024): size-independent properties --
    linearity of the integer projection in the basis planes, permutation equivariance, determinism."""
    require_gpu()
    rng = np.random.default_rng(20250820)
    E = np.asfortranarray(light_model["eigenfaces"][:, :10])
    mu = light_model["mean_face"]
    lam = light_model["eigenvalues"][:10]
    G = rng.normal(0, 1, (1024, 10)) * np.sqrt(lam)
    labels = (np.arange(1024) % 4).astype(np.int32)
    rec = ef.Recognizer(E, mu, G, labels=labels, metric=ef.METRIC_COSINE_G1)
    c = rng.normal(0, 1, (4096, 10)) * np.sqrt(lam)
    Q = np.clip(np.rint(mu + c @ E.T + rng.normal(0, 8, (4096, 10000))), 0, 255).astype(np.uint8)
    r1 = rec.recognize(Q, 0.5)
    r2 = rec.recognize(Q, 0.5)
    assert np.array_equal(r1.features, r2.features) and np.array_equal(r1.index, r2.index)      # deterministic
    perm = rng.permutation(4096)
    r3 = rec.recognize(Q[perm], 0.5)
    assert np.array_equal(r3.features, r1.features[perm]) and np.array_equal(r3.index, r1.index[perm])
    sub = rng.choice(4096, 256, replace=False)
    want_p = gen1.project_batch(Q[sub], E, mu)
    _close_proj(r1.features[sub], want_p)
    sims = gen1.cosine_matrix(want_p, G)
    assert np.array_equal(r1.index[sub], np.argmax(sims, axis=1))
    assert np.array_equal(r1.label[sub], np.where(sims.max(1) >= 0.5, labels[np.argmax(sims, 1)], -1))
    # the planted coefficients are recovered: projections correlate with c
    assert np.corrcoef(r1.features[:, 0], c[:, 0])[0, 1] > 0.99


@pytest.mark.parametrize("metric", [ef.METRIC_COSINE_SK, ef.METRIC_COSINE_G1])
@pytest.mark.parametrize("k,n_gallery", [(10, 1024), (3, 1), (16, 255), (32, 700), (7, 2049)])
def test_tensor_core_filter_is_exact_on_adversarial_galleries(metric, k, n_gallery):
    """The tcgen05 float16 filter + float64 re-score of the cluster kernel returns the SAME arg-best as the full float64
    scan (paths 0 and 1) on galleries built to defeat an approximate matcher: exact duplicates (lowest index must win),
    near-duplicates 1e-9 .. 1e-5 apart, zero rows, rows of wildly different norms, and queries that are gallery rows."""
    require_gpu()
    rng = np.random.default_rng(1000 * k + n_gallery + metric)
    D = 32 * 32
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    mu = rng.uniform(60, 200, D)
    G = rng.normal(size=(n_gallery, k)) * rng.uniform(1, 300, (1, k))
    if n_gallery >= 64:
        G[40] = G[7]                                            # exact duplicates, far apart in index
        G[n_gallery - 1] = G[7]
        for i, eps in enumerate((1e-9, 1e-7, 1e-6, 1e-5)):       # near duplicates below / at the filter resolution
            G[50 + i] = G[11] * (1.0 + 0.0) + eps * rng.normal(size=k) * np.abs(G[11]).max()
        G[60] = 0.0                                             # zero row
        G[61] = G[12] * 1e-6                                    # same direction, tiny norm
        G[62] = G[12] * 1e6
    B = 300
    coef = np.concatenate([G[rng.integers(0, n_gallery, B - 44)] + rng.normal(size=(B - 44, k)) * 0.5,
                           G[rng.integers(0, n_gallery, 44)]])
    coef = coef / np.abs(coef).max() * 40.0                      # keep the crops inside [0, 255]
    Q = np.clip(np.rint(mu + coef @ E.T), 0, 255).astype(np.uint8)
    Q[5] = np.clip(np.rint(mu), 0, 255).astype(np.uint8)         # (almost) zero projection
    rec = ef.Recognizer(E, mu, G, metric=metric)
    outs = []
    for mode in (1, 2):
        rec.use_tensor_cores(mode)
        outs.append(rec.recognize(Q, 0.3))
        assert rec.pipeline_timeouts() == 0
    for f in ("features", "score", "index", "label", "resid2"):
        assert np.array_equal(getattr(outs[0], f), getattr(outs[1], f)), f
    # and against numpy on the device features (float64): arg-max with first-maximum tie rule
    P = outs[1].features
    if metric == ef.METRIC_COSINE_SK:
        sims = gen2.sk_cosine_similarity(P, G)
    else:
        sims = gen1.cosine_matrix(P, G)
    srt = np.sort(sims, axis=1)
    decided = (srt[:, -1] - srt[:, -2]) > 1e-12 if n_gallery > 1 else np.ones(B, bool)
    assert np.array_equal(outs[1].index[decided], np.argmax(sims, axis=1)[decided])
    np.testing.assert_allclose(outs[1].score, sims.max(1), atol=SCORE_ATOL)


@pytest.mark.parametrize("serving", [(0, 16), (0, -16), (0, -8), (0, -3), (0, 1), (1, 0)])
@pytest.mark.parametrize("metric", [ef.METRIC_COSINE_G1, ef.METRIC_COSINE_SK])
def test_pipelined_submission_is_bit_identical(metric, serving, light_model, golden):
    """ef_model_submit_device / ef_model_flush_device return exactly what ef_model_recognize_device returns, for batches
    of changing sizes, including ragged tiles, a batch larger and one smaller than its predecessor, a single crop, and
    an interleaved ordinary call -- through the persistent queue kernel (serving kernel 0: all queued batches in one
    launch: adaptive depth 16, fixed depths 16 / 8 / 3, depth 1) and through the pipelined kernel (1: stream of batch i + match of batch i-1)."""
    torch = require_gpu()
    X = golden("gen1_light.npz")["X_u8"]
    rng = np.random.default_rng(77 + metric)
    k = 10
    E = light_model["eigenfaces"][:, :k].copy(order="F")
    G = light_model["projected_data"][:, :k].copy()
    kw = {}
    if metric == ef.METRIC_COSINE_SK:
        kw = dict(scale=rng.uniform(20.0, 60.0, 10000), pca_mean=rng.normal(0, 1e-3, 10000))
    rec = ef.Recognizer(E, light_model["mean_face"], G, metric=metric, labels=np.arange(len(G)) % 3, **kw)
    rec.set_serving(*serving)
    sizes = [300, 1000, 129, 1, 640, 128, 4096]
    batches = []
    for n in sizes:
        xb = torch.zeros((n, 10112), dtype=torch.uint8, device="cuda")
        xb[:, :10000] = torch.from_numpy(face_like(rng, X, n)).cuda()
        batches.append(xb)
    want = [{f: v.clone() for f, v in rec.recognize_device(xb, 0.8).items()} for xb in batches]
    torch.cuda.synchronize()
    outs = [rec.submit_device(xb, 0.8) for xb in batches]
    rec.flush_device()
    torch.cuda.synchronize()
    assert rec.pipeline_timeouts() == 0
    assert rec.serving_path() == (4 if serving[0] == 0 else 3)
    for i, (o, w) in enumerate(zip(outs, want)):
        for f in ("features", "score", "index", "label", "resid2"):
            assert torch.equal(o[f], w[f]), (i, sizes[i], f)
    # the same queue twice more (buffers, barriers and TMEM of a fresh launch), then a long queue of equal batches
    for _ in range(2):
        outs = [rec.submit_device(xb, 0.8) for xb in batches]
        rec.flush_device()
    big = [rec.submit_device(batches[6], 0.8) for _ in range(19)]
    rec.flush_device()
    torch.cuda.synchronize()
    assert rec.pipeline_timeouts() == 0
    for i, (o, w) in enumerate(zip(outs, want)):
        for f in ("features", "score", "index", "label", "resid2"):
            assert torch.equal(o[f], w[f]), ("second round", i, sizes[i], f)
    for o in big:
        for f in ("features", "score", "index", "label", "resid2"):
            assert torch.equal(o[f], want[6][f]), ("long queue", f)
    # an ordinary call in the middle of a pipeline flushes the pending batch first
    o1 = rec.submit_device(batches[0], 0.8)
    mid = rec.recognize_device(batches[1], 0.8)
    torch.cuda.synchronize()
    for f in ("score", "index", "label"):
        assert torch.equal(o1[f], want[0][f]) and torch.equal(mid[f], want[1][f]), f
    rec.flush_device()                                            # nothing pending: no-op
    # a host-buffer call (the model's own stream) while a batch is pending on torch's stream: ordered and correct
    o2 = rec.submit_device(batches[2], 0.8)
    host = rec.recognize(batches[3][:, :10000].cpu().numpy(), 0.8)
    torch.cuda.synchronize()
    assert torch.equal(o2["index"], want[2]["index"]) and torch.equal(o2["score"], want[2]["score"])
    assert np.array_equal(host.index, want[3]["index"].cpu().numpy())
    # shapes outside the pipelined kernel fall back to the immediate path
    rec50 = ef.gen1.recognizer_for(light_model)                    # k = 50
    o50 = rec50.submit_device(batches[0], 0.8)
    w50 = rec50.recognize_device(batches[0], 0.8)
    torch.cuda.synchronize()
    for f in ("features", "score", "index", "label"):
        assert torch.equal(o50[f], w50[f]), f


def test_random_shapes_all_paths_agree():
    """Fuzz: random pixel counts (not multiples of 16 / 128), component counts, digit-plane counts, gallery and batch
    sizes, metrics and scalers -- the CUDA-core path (0), the stream-K tensor-core path (1), the single cluster kernel
    (2) and the pipelined submission must return identical features / scores / indices / labels / residuals."""
    torch = require_gpu()
    rng = np.random.default_rng(20261018)
    for trial in range(24):
        D = int(rng.choice([64, 100, 257, 1024, 1600, 4096, 5000]))
        k = int(rng.integers(1, 41))
        S = int(rng.choice([0, 0, 4, 6, 8]))
        n = int(rng.integers(1, 1500))
        B = int(rng.integers(1, 700))
        metric = int(rng.choice([ef.METRIC_COSINE_SK, ef.METRIC_COSINE_G1, ef.METRIC_L2]))
        scaled = bool(rng.integers(0, 2))
        E = np.linalg.qr(rng.normal(size=(D, min(k, D))))[0]
        k = E.shape[1]
        G = rng.normal(size=(n, k)) * rng.uniform(0.5, 200, (1, k))
        if n > 3:
            G[n - 1] = G[0]                                        # a duplicate: lowest index must win everywhere
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D)) if scaled else {}
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), G, metric=metric, n_slices=S, labels=rng.integers(0, 5, n), **kw)
        X = rng.integers(0, 256, (B, D), dtype=np.uint8)
        thr = 0.3 if metric != ef.METRIC_L2 else 1e12
        outs = []
        for mode in (0, 1, 2):
            rec.use_tensor_cores(mode)
            outs.append(rec.recognize(X, thr))
            assert rec.pipeline_timeouts() == 0
        ld = (D + 15) // 16 * 16
        xd = torch.zeros((B, ld), dtype=torch.uint8, device="cuda")
        xd[:, :D] = torch.from_numpy(X).cuda()
        po = rec.submit_device(xd, thr)
        rec.flush_device()
        torch.cuda.synchronize()
        tag = f"trial {trial}: D={D} k={k} S={S} n={n} B={B} metric={metric} scaled={scaled}"
        for f in ("features", "score", "index", "label", "resid2"):
            a0 = getattr(outs[0], f)
            for o in outs[1:]:
                assert np.array_equal(a0, getattr(o, f)), (tag, f)
            assert np.array_equal(a0, po[f].cpu().numpy()), (tag, f, "pipelined")
        rec.close()


def test_small_gallery_fused_match_equals_generic_chain():
    """k > 32: the one-launch residual + match + label kernel (ef_match_small.cu) against the generic
    finalize_resid -> match -> reduce -> label chain (EF_NO_MATCH_SMALL=1): every output bit for bit."""
    import os
    require_gpu()
    rng = np.random.default_rng(77)
    cases = [(1600, 50, 229, 700, ef.METRIC_COSINE_G1, False), (1024, 178, 178, 300, ef.METRIC_COSINE_SK, True),
             (1024, 50, 590, 513, ef.METRIC_COSINE_SK, True), (900, 33, 1, 40, ef.METRIC_L2, False),
             (700, 600, 129, 50, ef.METRIC_COSINE_G1, False), (640, 97, 1000, 33, ef.METRIC_L2, True),
             (4096, 64, 257, 31, ef.METRIC_COSINE_G1, True)]
    for D, k, n, B, metric, scaled in cases:
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        G = rng.normal(size=(n, k)) * rng.uniform(0.5, 50, (1, k))
        if n > 10:
            G[n - 2] = G[3]
            G[7] = 0.0                                             # a zero gallery row: score 0.0 by the reference's rule
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D)) if scaled else {}
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), G, metric=metric, labels=rng.integers(0, 9, n), **kw)
        X = rng.integers(0, 256, (B, D), dtype=np.uint8)
        thr = 0.1 if metric != ef.METRIC_L2 else 1e12
        os.environ.pop("EF_NO_MATCH_SMALL", None)
        l0 = ef.launch_count()
        a = rec.recognize(X, thr)
        fused_launches = ef.launch_count() - l0
        os.environ["EF_NO_MATCH_SMALL"] = "1"
        try:
            l0 = ef.launch_count()
            b = rec.recognize(X, thr)
            chain_launches = ef.launch_count() - l0
        finally:
            os.environ.pop("EF_NO_MATCH_SMALL", None)
        # (small batches against a long gallery take the split chain by choice: B = 1 x 590 x 590 was 218 us in ONE
        # match_small CTA)
        if -(-B // 32) * 8 >= 148 or n * k < 16384:
            assert fused_launches < chain_launches
        else:
            assert fused_launches <= chain_launches
        for f in ("features", "score", "index", "label", "resid2"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), (D, k, n, B, metric, f)
        rec.close()


@pytest.mark.parametrize("metric", [ef.METRIC_COSINE_G1, ef.METRIC_COSINE_SK, ef.METRIC_L2])
def test_tensor_core_small_matcher_equals_float64_kernels(metric):
    """k = 33 ... 1024, galleries of <= 4096 rows (the shipped model shapes): the one-launch matcher whose all-pairs scan
    runs on tensor cores (ef_match_small_tc.cu: float16 hi/lo filter, float64 re-score of the rows inside the error
    band) against the float64 one-launch kernel (EF_NO_MATCH_SMALL_TC=1) and the generic chain (EF_NO_MATCH_SMALL=1):
    features, score, index, label and residual bit for bit -- on galleries with exact duplicates, near-duplicates
    (1e-9 ... 1e-13 relative), zero rows, rows of very different norms, and a batch with zero / duplicated crops."""
    import os
    require_gpu()
    rng = np.random.default_rng(1234 + metric)
    cases = [(1600, 50, 229, 4096, False), (1024, 178, 178, 300, True), (1024, 50, 590, 513, True),
             (900, 33, 1, 40, False), (640, 97, 1000, 1, True), (2048, 64, 257, 129, True), (512, 191, 64, 128, False),
             (512, 40, 4096, 200, False), (700, 50, 65, 5, False),
             (1024, 300, 500, 300, True), (1200, 590, 590, 200, True), (1100, 1024, 70, 140, False)]   # streamed K slabs
    for D, k, n, B, scaled in cases:
        if metric == ef.METRIC_L2 and k == 191:
            k = 190                                                # one extra component: 3 (k + 1) <= 576
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        mean = rng.uniform(40, 210, D)
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D)) if scaled else {}
        X = rng.integers(0, 256, (B, D), dtype=np.uint8)
        if B > 4:
            X[1] = X[0]
            X[2] = np.round(mean).astype(np.uint8)                 # (almost) the mean face: features near zero
        # gallery = projections of crops like the queries (so that the matches are close), then the adversarial rows
        Xg = rng.integers(0, 256, (n, D), dtype=np.uint8)
        if n >= B:
            Xg[rng.permutation(n)[:B]] = X
        A = (Xg.astype(np.float64) - mean)
        if scaled:
            A = A / kw["scale"] - kw["pca_mean"]
        G = A @ E
        if n > 40:
            G[n - 2] = G[3]                                        # exact duplicate: the lower index wins
            G[7] = 0.0
            G[11] = G[12] * (1 + 1e-9); G[13] = G[14] * (1 + 1e-13); G[15] = G[16] + 1e-11 * rng.normal(size=k)
            if metric != ef.METRIC_L2:
                G[20] = G[21] * 1e6; G[22] = G[23] * 1e-6          # same direction, other norms (cosine ties)
            else:
                G[20] = G[21] * 3.0
        rec = ef.Recognizer(E, mean, G, metric=metric, labels=rng.integers(0, 9, n), **kw)
        thr = 0.1 if metric != ef.METRIC_L2 else 1e12
        outs = []
        variants = [{}, {"EF_NO_MATCH_SMALL_TC": "1"}, {"EF_NO_MATCH_SMALL": "1"}, {"EF_MST_NO_FUSED_FINALIZE": "1"},
                    {"EF_MST_NO_BULK": "1"}, {"EF_MST_BNP": "64"}, {"EF_MST_BNP": "128"}, {"EF_MST_BNP": "256"},
                    {"EF_NO_PDL": "1"}, {"EF_NO_SLAB_COMBINE": "1"}, {"EF_MST_NO_CLUSTER": "1"}]
        for env in variants:
            os.environ.update(env)
            try:
                l0 = ef.launch_count()
                outs.append((rec.recognize(X, thr), ef.launch_count() - l0))
            finally:
                for name in env:
                    os.environ.pop(name, None)
        a, la = outs[0]
        assert la <= outs[2][1] and outs[1][1] <= outs[2][1]     # (two launches: query operand + filter / re-score)
        what = ["", "float64 one-launch kernel", "generic chain", "features formed by the slab finalize kernel",
                "rows staged by cp.async", "64-row pieces", "128-row pieces", "256-row pieces", "no dependent launches",
                "int32 plane slabs", "winners through global memory"]
        for (o, _), w in zip(outs[1:], what[1:]):
            for f in ("features", "score", "index", "label", "resid2"):
                assert np.array_equal(getattr(a, f), getattr(o, f)), (D, k, n, B, metric, f, "tc vs " + w)
        # a shorter batch after a longer one through the same scratch buffers (layout by capacity, stale rows zeroed)
        for b in (1, 3, 130):
            if b < B:
                few = rec.recognize(X[:b], thr)
                for f in ("features", "score", "index", "label", "resid2"):
                    assert np.array_equal(getattr(few, f), getattr(a, f)[:b]), (D, k, n, B, metric, f, b)
        assert rec.pipeline_timeouts() == 0
        rec.close()


def test_back_to_back_batches_overlap_safely():
    """Device-resident batches enqueued back to back, nothing synchronised in between: with the programmatic dependent
    launches the projection of batch i + 1 runs while the matcher of batch i drains, and the row sums run beside the
    projection.  Every batch of a 60-call run (two interleaved models, outputs kept per call) must equal the result of
    the same call made alone without dependent launches (EF_NO_PDL=1)."""
    import os
    torch = require_gpu()
    rng = np.random.default_rng(2024)
    B = 2048
    recs, xs = [], []
    for D, k, n, metric, scaled in ((1024, 50, 229, ef.METRIC_COSINE_G1, False), (1024, 178, 178, ef.METRIC_COSINE_SK, True),
                                    (768, 300, 300, ef.METRIC_COSINE_SK, True)):
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D)) if scaled else {}
        recs.append(ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 20, metric=metric,
                                  labels=rng.integers(0, 9, n), **kw))
        xs.append([torch.randint(0, 256, (B, D), dtype=torch.uint8, device="cuda") for _ in range(3)])
    os.environ["EF_NO_PDL"] = "1"
    try:
        want = []
        for m, rec in enumerate(recs):
            row = []
            for x in xs[m]:
                o = rec.recognize_device(x, 0.1)
                torch.cuda.synchronize()
                row.append({f: v.clone() for f, v in o.items() if v is not None})
            want.append(row)
    finally:
        os.environ.pop("EF_NO_PDL", None)
    got = []
    for i in range(60):
        m = i % len(recs) if i % 7 else (i // 7) % len(recs)        # mostly alternating models, sometimes the same twice
        j = i % 3
        got.append((m, j, recs[m].recognize_device(xs[m][j], 0.1)))
    torch.cuda.synchronize()
    for m, j, o in got:
        for f, v in want[m][j].items():
            assert torch.equal(o[f], v), (m, j, f)
    for rec in recs:
        assert rec.pipeline_timeouts() == 0
        rec.close()


def test_projection_tail_split_is_bit_identical():
    """More (crop tile, column tile) pairs than SMs: the tiles of the projection's last, partial wave are split along K and
    their partial (hi, lo) tiles summed by the slab finalize.  Same features and answers as without the tail split
    (EF_TC_NO_TAIL_SPLIT=1), as the int32 plane slabs and as the generic chain -- also for k <= 191, where the matcher's
    query kernel otherwise reads the slabs itself."""
    import os
    require_gpu()
    rng = np.random.default_rng(777)
    for D, k, n, B in ((384, 300, 200, 4096), (512, 590, 300, 3000), (256, 178, 178, 8192)):
        E = np.linalg.qr(rng.normal(size=(max(D, k), k)))[0][:D]
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 20, metric=ef.METRIC_COSINE_SK,
                            labels=rng.integers(0, 9, n), scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D))
        X = rng.integers(0, 256, (B, D), dtype=np.uint8)
        a = rec.recognize(X, 0.1)
        for env in ("EF_TC_NO_TAIL_SPLIT", "EF_NO_SLAB_COMBINE", "EF_NO_MATCH_SMALL"):
            os.environ[env] = "1"
            try:
                o = rec.recognize(X, 0.1)
            finally:
                os.environ.pop(env, None)
            for f in ("features", "score", "index", "label", "resid2"):
                assert np.array_equal(getattr(a, f), getattr(o, f)), (D, k, n, B, env, f)
        few = rec.recognize(X[:130], 0.1)                          # no tail at this size, the same slab buffer
        assert np.array_equal(few.index, a.index[:130]) and np.array_equal(few.features, a.features[:130])
        rec.close()


def test_split_k_slabs_equal_stream_k_atomics():
    """k > 32 on tensor cores: the split-K schedule that STORES partial tiles into slabs (default) against the stream-K
    schedule that merges them with int32 RED atomics (EF_NO_SLABS=1) and against the CUDA-core path (mode 0)."""
    import os
    require_gpu()
    rng = np.random.default_rng(99)
    # (D, k, n, B): 1..3 column tiles, 1..8 K ranges per tile, more tiles than SMs, K blocks fewer than SMs / tiles
    cases = [(1024, 50, 229, 4096), (1024, 50, 229, 20000), (640, 178, 178, 300), (4096, 64, 100, 129),
             (256, 40, 64, 1), (10000, 50, 229, 1000), (130, 33, 10, 257)]
    for D, k, n, B in cases:
        E = np.linalg.qr(rng.normal(size=(D, min(k, D))))[0]
        k = E.shape[1]
        G = rng.normal(size=(n, k)) * 30
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), G, metric=ef.METRIC_COSINE_G1)
        X = rng.integers(0, 256, (B, D), dtype=np.uint8)
        os.environ.pop("EF_NO_SLABS", None)
        rec.use_tensor_cores(1)
        a = rec.recognize(X, 0.2)
        a2 = rec.recognize(X, 0.2)                                 # nothing left behind by the first call
        os.environ["EF_NO_SLABS"] = "1"
        try:
            b = rec.recognize(X, 0.2)
        finally:
            os.environ.pop("EF_NO_SLABS", None)
        os.environ["EF_TC_MULTICAST"] = "1"                        # basis tile TMA-multicast across crop-tile clusters
        try:
            e = rec.recognize(X, 0.2)
        finally:
            os.environ.pop("EF_TC_MULTICAST", None)
        rec.use_tensor_cores(0)
        c = rec.recognize(X, 0.2)
        for f in ("features", "score", "index", "label", "resid2"):
            for other in (a2, b, c, e):
                assert np.array_equal(getattr(a, f), getattr(other, f)), (D, k, n, B, f)
        assert rec.pipeline_timeouts() == 0
        rec.close()


def test_async_host_submit_wait_equals_recognize():
    """ef_model_submit_host / ef_model_wait_host: two batches in flight, results identical to the synchronous call;
    a third submit without a wait, a wait on an idle ticket and a synchronous call with batches in flight are refused."""
    torch = require_gpu()
    rng = np.random.default_rng(31)
    D, k, n = 10000, 10, 300
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    rec = ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 50, metric=ef.METRIC_COSINE_G1,
                        labels=rng.integers(0, 4, n))
    batches = [torch.from_numpy(rng.integers(0, 256, (B, D), dtype=np.uint8)).pin_memory().numpy()
               for B in (4096, 700, 2500, 1, 4096)]
    want = [rec.recognize(x, 0.3) for x in batches]
    tickets = [rec.submit(batches[0], 0.3), rec.submit(batches[1], 0.3)]
    with pytest.raises(ef.EigenfacesError):
        rec.submit(batches[2], 0.3)                               # both slots busy
    with pytest.raises(ef.EigenfacesError):
        rec.recognize(batches[2], 0.3)                            # synchronous call while batches are in flight
    got = []
    for i in range(len(batches)):
        got.append(rec.wait(tickets[i % 2]))
        if i + 2 < len(batches):
            tickets[i % 2] = rec.submit(batches[i + 2], 0.3)
    for a, b in zip(want, got):
        for f in ("features", "score", "index", "label", "resid2"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    with pytest.raises(ef.EigenfacesError):
        rec.wait(tickets[0])                                      # nothing in flight behind this ticket any more
    t = rec.submit(batches[1], 0.3, want_features=False, want_residual=False)
    r = rec.wait(t)
    assert r.features is None and r.resid2 is None and np.array_equal(r.label, want[1].label)
    rec.close()


def test_close_with_batches_in_flight():
    """Destroying a model while submitted batches are still running must drain them first (no use-after-free)."""
    torch = require_gpu()
    rng = np.random.default_rng(5)
    D, k, n = 4096, 12, 200
    E = np.linalg.qr(rng.normal(size=(D, k)))[0]
    for _ in range(3):
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 50, metric=ef.METRIC_COSINE_SK)
        x = torch.from_numpy(rng.integers(0, 256, (4096, D), dtype=np.uint8)).pin_memory().numpy()
        rec.submit(x, 0.3)
        rec.submit(x, 0.3)
        xd = torch.from_numpy(x).cuda()
        rec.submit_device(xd, 0.3)
        rec.close()                                               # never waited for
    torch.cuda.synchronize()
    rec = ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 50, metric=ef.METRIC_COSINE_SK)
    assert rec.recognize(x[:8], 0.3).index.shape == (8,)
    rec.close()


def test_few_query_match_is_bit_identical_to_the_batched_kernels():
    """The reference calls recognition with ONE face at a time: B <= 8 goes through match_few_kernel (one gallery row per
    thread).  Its rows must equal, bit for bit, the same crops recognised inside a large batch (match_small_kernel)."""
    require_gpu()
    rng = np.random.default_rng(91)
    cases = [(1024, 178, 178, ef.METRIC_COSINE_SK, True), (4096, 590, 590, ef.METRIC_COSINE_SK, True),
             (1600, 50, 229, ef.METRIC_COSINE_G1, False), (640, 97, 1000, ef.METRIC_L2, True)]
    for D, k, n, metric, scaled in cases:
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        G = rng.normal(size=(n, k)) * rng.uniform(0.5, 50, (1, k))
        G[n - 2] = G[3]                                              # an exact duplicate: the lower index must win
        G[7] = 0.0
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-2, D)) if scaled else {}
        rec = ef.Recognizer(E, rng.uniform(40, 210, D), G, metric=metric, labels=rng.integers(0, 9, n), **kw)
        X = rng.integers(0, 256, (700, D), dtype=np.uint8)
        thr = 0.1 if metric != ef.METRIC_L2 else 1e12
        big = rec.recognize(X, thr)
        for b in (1, 2, 5, 8):
            few = rec.recognize(X[:b], thr)
            for f in ("features", "score", "index", "label", "resid2"):
                assert np.array_equal(getattr(few, f), getattr(big, f)[:b]), (D, k, n, metric, b, f)
        rec.close()


def test_all_models_one_call_equals_per_model_calls():
    """ef_models_recognize_boxes_host (one upload, K1 once, every model's K2, one download) returns, model by model, exactly
    what ef_model_recognize_boxes_host returns for that model alone; a box outside its frame is refused, not recognised."""
    require_gpu()
    rng = np.random.default_rng(2024)
    D = 64 * 64
    recs = []
    for k, n, metric in ((40, 178, ef.METRIC_COSINE_SK), (60, 300, ef.METRIC_COSINE_SK), (10, 500, ef.METRIC_COSINE_G1)):
        E = np.linalg.qr(rng.normal(size=(D, k)))[0]
        kw = dict(scale=rng.uniform(5.0, 80.0, D), pca_mean=rng.normal(0, 1e-3, D)) if metric == ef.METRIC_COSINE_SK else {}
        recs.append(ef.Recognizer(E, rng.uniform(40, 210, D), rng.normal(size=(n, k)) * 30, metric=metric,
                                  labels=rng.integers(0, 5, n), **kw))
    frames = rng.integers(0, 256, (3, 240, 320, 3), dtype=np.uint8)
    for B in (1, 5, 70):
        boxes = np.stack([rng.integers(0, 3, B), rng.integers(0, 100, B), rng.integers(0, 60, B), rng.integers(40, 200, B),
                          rng.integers(40, 170, B)], axis=1).astype(np.int32)
        score, index, label = ef.engine.recognize_boxes_all_models(recs, frames, boxes, 64, 0.3)
        assert score.shape == (3, B)
        for i, rec in enumerate(recs):
            one = rec.recognize_boxes(frames, boxes, 64, 0.3, want_features=False, want_residual=False)
            assert np.array_equal(score[i], one.score) and np.array_equal(index[i], one.index) and np.array_equal(label[i], one.label)
    bad = np.array([[0, 300, 10, 64, 64]], dtype=np.int32)               # reaches past the right edge of the frame
    with pytest.raises(ef.EigenfacesError):
        ef.engine.recognize_boxes_all_models(recs, frames, bad, 64, 0.3)
    score, _, _ = ef.engine.recognize_boxes_all_models(recs, frames, boxes, 64, 0.3)     # and the models still work
    assert np.isfinite(score).all()
    for rec in recs:
        rec.close()

