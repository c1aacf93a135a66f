#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ from the reference itself.

Run ONLY in the build container (needs /root/reference, cv2, sklearn):
    python tests/golden/make_golden.py
It imports the reference's modules by path (they are import-side-effect free), runs the reference's
own functions on the reference's shipped data and on seeded synthetic inputs, and stores inputs +
outputs as small .npz files.  The GPU box has no /root/reference; tests read only the .npz files.

Fixtures:
  gen1_light.npz   faces/Light_version decoded (u8 [229,10000]) + shipped models/Joseph_Lai_light_pca_model.pkl
                   arrays + models/Joseph_Lai_light_model_info.json ratios + live useless/train.py:manual_pca run
  gen1_dark.npz    faces/Dark_version decoded (u8 [512,10000]) + models/Joseph_Lai_dark_model_info.json ratios
                   + live manual_pca outputs
  gen1_recog.npz   seeded queries + outputs of useless/scan.py:recognize_face / recognize_face_dual_model
  gen2_joseph.npz  faces/lock_version/Joseph_Lai through train-v5.py loader (u8 [178,4096]) + live
                   train-v5.py:train_pca_model outputs (solver full, k = N) + multi_person_model_info.json scalars
  gen2_recog.npz   a k=20 reference model + outputs of scan-template-v4.py:extract_face_features /
                   recognize_face_with_model / recognize_face_all_models on seeded BGR crops
  gen2_shipped.npz the shipped faces/lock_version/Joseph_Lai/face_model.pkl (77 faces, k=76, float32) as arrays
                   + its 77 training crops decoded, + reference-arithmetic recognition of those crops
  preprocess.npz   seeded random crops (shape + seed only) and cv2.cvtColor/cv2.resize outputs
"""
import contextlib
import importlib.util
import io
import json
import os
import pickle
import sys
import warnings

import cv2
import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_module(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    warnings.filterwarnings("ignore")
    os.chdir(REF)  # the reference uses relative paths
    train1 = load_module("ref_train1", "useless/train.py")
    scan1 = load_module("ref_scan1", "useless/scan.py")
    train5 = load_module("ref_train5", "train-v5.py")
    scan4 = load_module("ref_scan4", "scan-template-v4.py")

    # ------------------------------------------------------------------ Gen-1 fit
    gen1 = {}
    for version, folder in (("light", "faces/Light_version"), ("dark", "faces/Dark_version")):
        with quiet():
            X, names = train1.load_face_images(folder)
            ef, mean, proj, ev = train1.manual_pca(X, 50)
        assert np.array_equal(X, np.rint(X)) and X.min() >= 0 and X.max() <= 255
        info = json.load(open(f"models/Joseph_Lai_{version}_model_info.json"))
        gen1[version] = dict(X=X, ef=ef, mean=mean, proj=proj, ev=ev)
        arrays = dict(
            X_u8=X.astype(np.uint8), filenames=np.array(names),
            info_evr10=np.array(info["explained_variance_ratio"]),
            info_n_components=info["n_components"], info_face_dimensions=info["face_dimensions"],
            info_n_training_images=info["n_training_images"],
            ref_mean_face=mean, ref_eigenvalues=ev, ref_projected=proj,
            ref_eigenfaces_f64_first8=np.ascontiguousarray(ef[:, :8]),
        )
        if version == "light":
            pk = pickle.load(open("models/Joseph_Lai_light_pca_model.pkl", "rb"))
            arrays.update(
                pkl_mean_face=pk["mean_face"], pkl_eigenvalues=pk["eigenvalues"],
                pkl_projected=np.ascontiguousarray(pk["projected_data"]),
                pkl_eigenfaces_f64_first8=np.ascontiguousarray(pk["eigenfaces"][:, :8]),
                pkl_eigenfaces_f16=np.ascontiguousarray(pk["eigenfaces"]).astype(np.float16),
                pkl_filenames=np.array(pk["training_filenames"]),
                pkl_eigenfaces_is_fortran=pk["eigenfaces"].flags["F_CONTIGUOUS"],
            )
            gen1["light_pkl"] = pk
            # pin tightness of the live reference run against the shipped pickle
            print("light live-vs-pkl: eigenfaces %.1e eigenvalues(rel) %.1e projected %.1e" % (
                np.abs(ef - pk["eigenfaces"]).max(), np.abs(ev / pk["eigenvalues"] - 1).max(),
                np.abs(proj - pk["projected_data"]).max()))
        save(f"gen1_{version}.npz", **arrays)

    # ------------------------------------------------------------------ Gen-1 recognition
    rng = np.random.default_rng(20250820)
    Xl, Xd = gen1["light"]["X"], gen1["dark"]["X"]
    q = []
    q.append(Xl[rng.integers(0, len(Xl), 16)])                                   # exact training crops
    q.append(np.clip(np.rint(Xl[rng.integers(0, len(Xl), 16)] + rng.normal(0, 8, (16, 10000))), 0, 255))
    q.append(np.clip(np.rint(Xd[rng.integers(0, len(Xd), 16)] + rng.normal(0, 8, (16, 10000))), 0, 255))
    q.append(rng.integers(0, 256, (14, 10000)).astype(np.float64))               # uniform noise
    q.append(np.zeros((1, 10000)))                                               # all black
    q.append(np.full((1, 10000), 255.0))                                         # all white
    Q = np.concatenate(q).astype(np.uint8)
    light_model = gen1["light_pkl"]
    d = gen1["dark"]
    dark_model = dict(eigenfaces=d["ef"], mean_face=d["mean"], projected_data=d["proj"], eigenvalues=d["ev"],
                      person_name="Joseph_Lai_dark", n_components=50, face_dimensions=10000)
    sims_l, sims_d, dual = [], [], []
    projs_l = []
    for row in Q:
        v = row.flatten().astype(np.float64)                                     # useless/scan.py:252
        projs_l.append(scan1.project_face_to_eigenspace(v, light_model["eigenfaces"], light_model["mean_face"]))
        _, sl, _ = scan1.recognize_face(v, light_model, 0.8)
        _, sd, _ = scan1.recognize_face(v, dark_model, 0.8)
        name, best, rec, dsim, lsim = scan1.recognize_face_dual_model(v, dark_model, light_model, 0.8)
        sims_l.append(sl); sims_d.append(sd)
        dual.append((name == "Joseph_Lai_dark", best, rec, dsim, lsim))
    save("gen1_recog.npz", queries_u8=Q, ref_proj_light=np.array(projs_l), ref_sim_light=np.array(sims_l),
         ref_sim_dark=np.array(sims_d), ref_dual_name_is_dark=np.array([x[0] for x in dual]),
         ref_dual_best=np.array([x[1] for x in dual]), ref_dual_recognized=np.array([x[2] for x in dual]),
         threshold=0.8)

    # ------------------------------------------------------------------ Gen-2 fit (train-v5, solver full, k = N)
    person_dir = "faces/lock_version/Joseph_Lai"
    tr = train5.MultiFaceTrainer(n_components=178)
    with quiet():
        n = tr.load_face_images_from_json(os.path.join(person_dir, "Joseph_Lai_faces_detection.json"), person_dir)
        assert n == 178 and tr.train_pca_model()
    info = json.load(open(os.path.join(person_dir, "multi_person_model_info.json")))
    X2 = tr.face_images
    assert X2.dtype == np.uint8
    save("gen2_joseph.npz", X_u8=X2,
         info_total_faces=info["total_faces"], info_n_components=info["n_components"],
         info_evr_sum=info["explained_variance_ratio"],
         ref_solver=np.array(tr.pca._fit_svd_solver),
         ref_mean_face=tr.mean_face, ref_scaler_mean=tr.scaler.mean_, ref_scaler_var=tr.scaler.var_,
         ref_scaler_scale=tr.scaler.scale_, ref_pca_mean=tr.pca.mean_,
         ref_components_first10=tr.pca.components_[:10], ref_explained_variance=tr.pca.explained_variance_,
         ref_explained_variance_ratio=tr.pca.explained_variance_ratio_,
         ref_singular_values=tr.pca.singular_values_, ref_noise_variance=tr.pca.noise_variance_,
         ref_face_features_first20=np.ascontiguousarray(tr.face_features[:, :20]))

    # ------------------------------------------------------------------ Gen-2 recognition with a k=20 reference model
    models = {}
    scanner = scan4.MultiModelFaceScanner()
    for person, cap in (("Joseph_Lai", None), ("ruisheng", 120)):
        t = train5.MultiFaceTrainer(n_components=20)
        pdir = f"faces/lock_version/{person}"
        with quiet():
            t.load_face_images_from_json(os.path.join(pdir, f"{person}_faces_detection.json"), pdir)
            if cap:
                t.face_images = t.face_images[:cap]
                t.face_labels = t.face_labels[:cap]
                t.face_info = t.face_info[:cap]
            assert t.train_pca_model()
        md = dict(pca=t.pca, scaler=t.scaler, face_features=t.face_features, face_labels=t.face_labels,
                  person_id_map=t.person_id_map, n_components=20, mean_face=t.mean_face, eigenfaces=t.eigenfaces,
                  face_shape=(64, 64))
        models[person] = md
        scanner.models[person] = {"model_data": md}
    rng = np.random.default_rng(4096)
    specs, feats, single, multi = [], [], [], []
    train_imgs = sorted(f for f in os.listdir("faces/lock_version/Joseph_Lai") if "_face_" in f and f.endswith(".jpg"))
    crops = []
    for i in range(24):
        kind = i % 4
        if kind == 0:      # a real BGR training crop, stored via its resized form is impossible -> keep the raw bytes small
            img = cv2.imread(os.path.join("faces/lock_version/Joseph_Lai", train_imgs[i]))
            img = cv2.resize(img, (96 + 8 * (i % 5), 96 + 8 * (i % 5)), interpolation=cv2.INTER_AREA)
        elif kind == 1:    # gray crop of another person
            img = cv2.imread(os.path.join("faces/lock_version/ruisheng",
                                          sorted(os.listdir("faces/lock_version/ruisheng"))[20 + i]), cv2.IMREAD_GRAYSCALE)
            img = cv2.resize(img, (128, 128), interpolation=cv2.INTER_AREA)
        elif kind == 2:    # random BGR noise, odd size
            img = rng.integers(0, 256, (int(rng.integers(40, 160)), int(rng.integers(40, 160)), 3), dtype=np.uint8)
        else:              # smooth gradient + noise, gray
            h, w = int(rng.integers(64, 200)), int(rng.integers(64, 200))
            yy, xx = np.mgrid[0:h, 0:w]
            img = np.clip(128 + 60 * np.sin(xx / 17.0) + 40 * np.cos(yy / 11.0) + rng.normal(0, 6, (h, w)), 0, 255).astype(np.uint8)
        crops.append(img)
        per_model = []
        for person, md in models.items():
            f = scanner.extract_face_features(img, md)
            per_model.append(f)
            pid, name, sim = scanner.recognize_face_with_model(f, md, 0.7)
            single.append((int(pid), name, float(sim)))
        feats.append(np.stack(per_model))
        with quiet():
            pid, name, conf = scanner.recognize_face_all_models(img, 0.8)
        multi.append((int(pid), name, float(conf)))
    arrays = {}
    for i, c in enumerate(crops):
        arrays[f"crop_{i:02d}"] = c
    for person, md in models.items():
        arrays[f"{person}_scaler_mean"] = md["scaler"].mean_
        arrays[f"{person}_scaler_scale"] = md["scaler"].scale_
        arrays[f"{person}_scaler_var"] = md["scaler"].var_
        arrays[f"{person}_components"] = md["pca"].components_
        arrays[f"{person}_pca_mean"] = md["pca"].mean_
        arrays[f"{person}_explained_variance"] = md["pca"].explained_variance_
        arrays[f"{person}_face_features"] = md["face_features"]
        arrays[f"{person}_face_labels"] = md["face_labels"]
        arrays[f"{person}_mean_face"] = md["mean_face"]
    save("gen2_recog.npz", persons=np.array(list(models)), n_crops=len(crops),
         ref_features=np.stack(feats),
         ref_single_pid=np.array([s[0] for s in single]).reshape(len(crops), -1),
         ref_single_name=np.array([s[1] for s in single]).reshape(len(crops), -1),
         ref_single_sim=np.array([s[2] for s in single]).reshape(len(crops), -1),
         ref_multi_pid=np.array([m[0] for m in multi]), ref_multi_name=np.array([m[1] for m in multi]),
         ref_multi_conf=np.array([m[2] for m in multi]), **arrays)

    # ------------------------------------------------------------------ shipped Gen-2 pickle (data fixture)
    pk = pickle.load(open("faces/lock_version/Joseph_Lai/face_model.pkl", "rb"))
    pca, sc = pk["pca_model"], pk["scaler"]
    X77 = []
    for fi in pk["face_info"]:
        img = cv2.imread(os.path.join(person_dir, fi["image_filename"]))
        X77.append(cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (64, 64)).flatten())
    X77 = np.array(X77)
    md = dict(pca=pca, scaler=sc, face_features=pk["face_features"], face_labels=pk["face_labels"],
              person_id_map=pk["person_id_map"])
    f77, pid77, sim77, idx77 = [], [], [], []
    for row in X77:
        f = scanner.extract_face_features(row.reshape(64, 64), md)
        pid, name, sim = scanner.recognize_face_with_model(f, md, 0.7)
        f77.append(f); pid77.append(pid); sim77.append(sim)
        from sklearn.metrics.pairwise import cosine_similarity
        idx77.append(int(np.argmax(cosine_similarity([f], md["face_features"])[0])))
    save("gen2_shipped.npz", X_u8=X77, scaler_mean=sc.mean_, scaler_scale=sc.scale_, scaler_var=sc.var_,
         components=pca.components_, pca_mean=pca.mean_, explained_variance=pca.explained_variance_,
         face_features=pk["face_features"], face_labels=pk["face_labels"], mean_face=pk["mean_face"],
         person_names=np.array(list(pk["person_id_map"].keys())),
         person_ids=np.array(list(pk["person_id_map"].values())),
         ref_features=np.array(f77), ref_pid=np.array(pid77), ref_sim=np.array(sim77), ref_argmax=np.array(idx77))

    # ------------------------------------------------------------------ preprocess (cv2 itself)
    rng = np.random.default_rng(777)
    cases = []
    fixed = [(128, 128, 64, 64, 1), (200, 200, 100, 100, 3), (100, 100, 100, 100, 1), (64, 64, 64, 64, 3),
             (30, 30, 64, 64, 1), (31, 47, 100, 100, 3), (267, 267, 64, 64, 3), (300, 220, 100, 100, 1),
             (1, 1, 64, 64, 1), (2, 3, 64, 64, 3), (129, 127, 64, 64, 1), (640, 480, 100, 100, 3)]
    for i in range(48):
        if i < len(fixed):
            h, w, dh, dw, c = fixed[i]
        else:
            h, w = int(rng.integers(20, 320)), int(rng.integers(20, 320))
            if i % 2:
                w = h
            dh = dw = (64, 100)[i % 3 == 0]
            c = (1, 3)[i % 2]
        seed = 1000 + i
        img = np.random.default_rng(seed).integers(0, 256, (h, w, c) if c == 3 else (h, w), dtype=np.uint8)
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if c == 3 else img
        out = cv2.resize(gray, (dw, dh))
        cases.append((h, w, c, dw, dh, seed, out))
    arrays = {f"out_{i:02d}": c[6] for i, c in enumerate(cases)}
    # one frame + boxes case (ROI views of a bigger frame, as in scan-template-v4.py:360)
    frame = np.random.default_rng(4242).integers(0, 256, (270, 480, 3), dtype=np.uint8)
    boxes = np.array([[0, 0, 64, 64], [10, 20, 128, 128], [200, 100, 99, 77], [479 - 50, 269 - 40, 50, 40],
                      [5, 5, 200, 200], [300, 0, 100, 100], [17, 33, 61, 245 - 33]], dtype=np.int32)
    roi = np.stack([cv2.resize(cv2.cvtColor(frame[y:y + h, x:x + w], cv2.COLOR_BGR2GRAY), (100, 100)).flatten()
                    for x, y, w, h in boxes])
    save("preprocess.npz", specs=np.array([c[:6] for c in cases], dtype=np.int64), frame_seed=4242,
         frame_shape=np.array(frame.shape), boxes=boxes, roi_out_100=roi, cv2_version=np.array(cv2.__version__),
         **arrays)


if __name__ == "__main__":
    sys.exit(main())
