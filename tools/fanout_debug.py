import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef
from sklearn.decomposition import PCA
from sklearn.preprocessing import StandardScaler
rng = np.random.default_rng(44)
scanner = ef.gen2.MultiModelFaceScanner()
for name, n in (("a", 40), ("b", 60)):
    D = 4096
    comps = np.linalg.qr(rng.normal(size=(D, n)))[0].T.copy()
    pca = PCA(n_components=n); pca.components_, pca.mean_ = comps, rng.normal(0, 1e-15, D)
    sc = StandardScaler(); sc.mean_, sc.scale_ = rng.uniform(60, 200, D), rng.uniform(20, 60, D)
    md = {"pca": pca, "scaler": sc, "face_features": rng.normal(size=(n, n)) * 30, "face_labels": np.zeros(n, dtype=int),
          "person_id_map": {name: 0}, "n_components": n}
    scanner.models[name] = {"model_data": md}
crop = rng.integers(0, 256, (180, 160, 3), dtype=np.uint8)
print("all models:", scanner.recognize_face_all_models(crop, 0.0))
for name, info in scanner.models.items():
    f = scanner.extract_face_features(crop, info["model_data"])
    print(name, scanner.recognize_face_with_model(f, info["model_data"], 0.0))
    rec = ef.gen2.recognizer_for(info["model_data"])
    r = rec.recognize_boxes(crop, [[0, 0, 160, 180]], 64, 0.0)
    print("   host path:", r.score, r.index, r.label)
    x = ef.preprocess_device(torch.from_numpy(crop[None]).cuda(), torch.tensor([[0, 0, 0, 160, 180]], dtype=torch.int32, device="cuda"), 64)
    o = rec.recognize_device(x, 0.0, want_residual=False)
    print("   device path:", o["score"].cpu().numpy(), o["index"].cpu().numpy(), o["label"].cpu().numpy())
