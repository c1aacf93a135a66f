"""CPU-side checks of the round-2 rows: the F3 oracle pinned by the reference's own ManualPCA / ManualStandardScaler
outputs, the reader / writer of the manual generation's pickles, and the Gen-1 model store formats -- no GPU needed."""
import io
import json
import os
import pickle

import numpy as np

import eigenfaces_b200 as ef
from conftest import GOLDEN
from oracle import gen2


def _align(A, B):
    """Sign of every row of A made equal to B's (the reference leaves eigenvector signs to LAPACK)."""
    s = np.sign(np.sum(A * B, axis=1))
    s[s == 0] = 1.0
    return s


def test_oracle_manual_generation_matches_the_reference(golden):
    """oracle/gen2.py:manual_* (F3) against scripts/manual/train-v2.py run in the build container."""
    g = golden("manual_joseph.npz")
    X = golden("gen2_joseph.npz")["X_u8"]
    k = int(g["k"])
    mean, std = gen2.manual_scaler_fit(X)
    np.testing.assert_allclose(mean, g["ref_scaler_mean"], rtol=1e-14)
    np.testing.assert_allclose(std, g["ref_scaler_scale"], rtol=1e-13)
    fit = gen2.manual_pca_fit((X - mean) / std, k)
    s = _align(fit["components"], g["ref_components"])
    np.testing.assert_allclose(fit["components"] * s[:, None], g["ref_components"], atol=1e-9)
    np.testing.assert_allclose(fit["explained_variance_ratio"], g["ref_evr"], rtol=1e-9)


def test_reference_manual_pickle_loads_into_the_mirror_classes(golden):
    g = golden("manual_joseph.npz")
    model = ef.manual.load_manual_pickle(os.path.join(GOLDEN, "manual_model.pkl"))
    assert isinstance(model["pca"], ef.manual.ManualPCA) and isinstance(model["scaler"], ef.manual.ManualStandardScaler)
    assert np.array_equal(model["pca"].components_, g["ref_components"])
    assert np.array_equal(model["scaler"].scale_, g["ref_scaler_scale"])
    assert set(model) == {"pca", "scaler", "face_features", "face_labels", "face_info", "person_id_map", "n_components",
                          "mean_face", "eigenfaces", "face_shape", "training_date"}
    # written back: the estimators are named like the reference's script names them (run as __main__)
    buf = io.BytesIO()
    ef.manual.dump_manual_pickle(model, buf)
    raw = buf.getvalue()
    assert b"__main__" in raw and b"ManualPCA" in raw and b"ManualStandardScaler" in raw and b"eigenfaces_b200" not in raw
    again = ef.manual.load_manual_pickle(raw)
    assert np.array_equal(again["pca"].components_, model["pca"].components_)
    assert ef.manual.ManualPCA.__module__.endswith("manual")          # the class itself is left untouched


def test_gen1_model_store_schema_equals_the_reference(tmp_path, golden):
    """gen1.save_pca_model / load_pca_model against useless/train.py:130-192 and useless/scan.py:9-33."""
    ref_path = os.path.join(GOLDEN, "gen1_store", "toy_v1_pca_model.pkl")
    ref = ef.gen1.load_pca_model(ref_path)                            # a pickle the reference wrote
    assert ref is not None and ref["person_name"] == "toy" and ref["version"] == "v1"
    assert ref["eigenfaces"].flags["F_CONTIGUOUS"] == bool(golden("gen1_store.npz")["eigenfaces_is_fortran"])
    out = ef.gen1.save_pca_model(ref["eigenfaces"], ref["mean_face"], ref["projected_data"], ref["eigenvalues"],
                                 ref["training_filenames"], "toy", str(tmp_path), "v1")
    assert os.path.basename(out) == "toy_v1_pca_model.pkl"
    mine = pickle.load(open(out, "rb"))
    assert list(mine) == list(ref)                                    # same keys in the same order
    for key in ref:
        if key == "training_timestamp":
            continue
        a, b = mine[key], ref[key]
        assert type(a) is type(b), key
        if isinstance(b, np.ndarray):
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), key
            assert a.flags["F_CONTIGUOUS"] == b.flags["F_CONTIGUOUS"], key
        else:
            assert a == b, key
    info_ref = json.load(open(os.path.join(GOLDEN, "gen1_store", "toy_v1_model_info.json")))
    info = json.load(open(os.path.join(str(tmp_path), "toy_v1_model_info.json")))
    assert list(info) == list(info_ref)
    for key in info_ref:
        if key != "training_timestamp":
            assert info[key] == info_ref[key], key
    # unreadable file: message + None, like the reference
    assert ef.gen1.load_pca_model(str(tmp_path / "missing.pkl")) is None
    assert ef.gen1.load_dual_pca_models(out, str(tmp_path / "missing.pkl")) == (None, None)


def test_gen1_annotation_filter_rules():
    """useless/scan.py:283-287: unrecognised low-confidence detections and boxes below 200 x 200 are not drawn."""
    frame = np.zeros((600, 800, 3), np.uint8)
    dets = [(10, 300, 250, 250, "a", 0.9, True),       # drawn
            (300, 300, 150, 250, "a", 0.9, True),      # too small
            (500, 300, 250, 250, "a", 0.2, False)]     # not recognised and < 0.3
    out = ef.gen1.draw_face_annotations(frame, dets)
    assert out[:, :280].any() and not out[:, 290:].any()
    assert not frame.any()                                             # the input frame is not modified
