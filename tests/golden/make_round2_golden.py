#!/usr/bin/env python
"""Round-2 golden fixtures, generated from the reference itself (build container only: needs /root/reference).

    python tests/golden/make_round2_golden.py

  manual_joseph.npz   scripts/manual/train-v2.py: ManualStandardScaler + ManualPCA (k = 12) on the Joseph_Lai crops of
                      gen2_joseph.npz, and scripts/manual/scan-template-v2.py: FaceScanner.extract_face_features /
                      manual_cosine_similarity / recognize_face on the crops of gen2_recog.npz
  manual_model.pkl    the model dict of scripts/manual/train-v2.py:271-283 for that fit, pickled with the estimator classes
                      named __main__.ManualPCA / __main__.ManualStandardScaler -- byte for byte what the reference script
                      (run as __main__) writes
  gen1_store/         useless/train.py:save_pca_model + visualize_eigenfaces outputs for a 40 x 1024 synthetic training set:
                      toy_v1_pca_model.pkl, toy_v1_model_info.json, the JPEG renderings as arrays in gen1_store.npz
"""
import contextlib
import importlib.util
import io
import os
import pickle
import shutil
import sys
import tempfile

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    tr = load_module("manual_train_v2", "scripts/manual/train-v2.py")
    sc = load_module("manual_scan_v2", "scripts/manual/scan-template-v2.py")
    X = np.load(os.path.join(OUT, "gen2_joseph.npz"))["X_u8"]
    k = 12
    trainer = tr.FaceTrainer(n_components=k)
    trainer.face_images = X
    trainer.face_info = [{"face_id": i} for i in range(len(X))]
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.assign_labels_interactive("Joseph_Lai")
        assert trainer.train_pca_model()
    # the reference script runs as __main__: its pickles name the classes there
    main_mod = sys.modules["__main__"]
    for cls in (tr.ManualPCA, tr.ManualStandardScaler):
        cls.__module__ = "__main__"
        setattr(main_mod, cls.__name__, cls)
    tmp = tempfile.mkdtemp()
    model_path = os.path.join(tmp, "face_model.pkl")
    with contextlib.redirect_stdout(io.StringIO()):
        trainer.save_model(model_path)
    shutil.copy(model_path, os.path.join(OUT, "manual_model.pkl"))
    raw = open(model_path, "rb").read()
    assert b"__main__" in raw and b"ManualPCA" in raw
    # recognition through the reference's manual scanner (its own re-declared classes unpickle the model)
    for cls in (sc.ManualPCA, sc.ManualStandardScaler):
        cls.__module__ = "__main__"
        setattr(main_mod, cls.__name__, cls)
    scanner = sc.FaceScanner(model_path, "unused.json")
    with open(model_path, "rb") as f:
        scanner.model_data = pickle.load(f)
    scanner.is_loaded = True
    g = np.load(os.path.join(OUT, "gen2_recog.npz"))
    n_crops = int(g["n_crops"])
    feats, pids, names, confs = [], [], [], []
    for i in range(n_crops):
        f = scanner.extract_face_features(g[f"crop_{i:02d}"])
        pid, name, conf = scanner.recognize_face(f, threshold=0.7)
        feats.append(f); pids.append(pid); names.append(name); confs.append(conf)
    # the training crops themselves (64 x 64 gray): self recognition
    self_conf, self_idx = [], []
    for i in range(0, len(X), 7):
        f = scanner.extract_face_features(X[i].reshape(64, 64))
        sims = np.array([scanner.manual_cosine_similarity(f, kf) for kf in scanner.model_data["face_features"]])
        self_conf.append(sims.max()); self_idx.append(int(np.argmax(sims)))
    np.savez_compressed(
        os.path.join(OUT, "manual_joseph.npz"), k=k,
        ref_scaler_mean=trainer.scaler.mean_, ref_scaler_scale=trainer.scaler.scale_, ref_pca_mean=trainer.pca.mean_,
        ref_components=trainer.pca.components_, ref_evr=trainer.pca.explained_variance_ratio_,
        ref_features=trainer.face_features, ref_mean_face=trainer.mean_face,
        recog_features=np.array(feats), recog_pid=np.array(pids), recog_name=np.array(names), recog_conf=np.array(confs),
        self_rows=np.arange(0, len(X), 7), self_conf=np.array(self_conf), self_idx=np.array(self_idx))
    print("manual_joseph.npz", os.path.getsize(os.path.join(OUT, "manual_joseph.npz")) / 1e6, "MB;",
          "manual_model.pkl", os.path.getsize(os.path.join(OUT, "manual_model.pkl")) / 1e6, "MB")

    # ---- Gen-1 model store
    g1 = load_module("gen1_train", "useless/train.py")
    g1s = load_module("gen1_scan", "useless/scan.py")
    rng = np.random.default_rng(11)
    N, side, kk = 40, 32, 6
    D = side * side
    base = rng.normal(0, 1, (N, 5)) @ rng.normal(0, 1, (5, D))
    Xs = np.clip(np.rint(128 + 25 * base + rng.normal(0, 4, (N, D))), 0, 255).astype(np.uint8)
    store = os.path.join(OUT, "gen1_store")
    shutil.rmtree(store, ignore_errors=True)
    os.makedirs(store)
    with contextlib.redirect_stdout(io.StringIO()):
        E, mean, proj, ev = g1.manual_pca(Xs.astype(np.float64), kk)
        names_ = [f"face_{i:03d}.jpg" for i in range(N)]
        g1.save_pca_model(E, mean, proj, ev, names_, "toy", store, "v1")
        g1.visualize_eigenfaces(E, mean, store, "toy_v1")
        model = g1s.load_pca_model(os.path.join(store, "toy_v1_pca_model.pkl"))
        queries = np.clip(Xs[:8].astype(np.int32) + rng.integers(-3, 4, (8, D)), 0, 255).astype(np.uint8)
        sims = np.array([g1s.recognize_face(q.astype(np.float64), model, 0.7)[1] for q in queries])
    imgs = {}
    for f in sorted(os.listdir(store)):
        if f.endswith(".jpg"):
            imgs[f[:-4]] = cv2.imread(os.path.join(store, f), cv2.IMREAD_GRAYSCALE)
            os.remove(os.path.join(store, f))               # kept as arrays (the JPEG bytes depend on the encoder build)
    np.savez_compressed(os.path.join(OUT, "gen1_store.npz"), X_u8=Xs, k=kk, queries_u8=queries, ref_sims=sims,
                        eigenfaces_is_fortran=bool(model["eigenfaces"].flags["F_CONTIGUOUS"]),
                        **{"jpg_" + n: im for n, im in imgs.items()})
    print("gen1_store:", sorted(os.listdir(store)), "; jpgs:", sorted(imgs))


if __name__ == "__main__":
    main()
