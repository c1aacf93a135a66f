"""The scripts/manual generation of the reference ("manual PCA with sklearn-like classes"), backed by the B200 engine.

  ManualPCA / ManualStandardScaler   scripts/manual/train-v2.py:9-72 (re-declared in scripts/manual/scan-template-v2.py:9-72)
  FaceTrainer                        scripts/manual/train-v2.py:74-330  (train_pca_model :173-208, save_model :262-290)
  FaceScanner                        scripts/manual/scan-template-v2.py:74-300 (extract_face_features :206-231,
                                     manual_cosine_similarity :233-258, recognize_face :260-296)
The reference runs these scripts as __main__, so its pickles name the two estimator classes `__main__.ManualPCA` and
`__main__.ManualStandardScaler`.  save_model writes exactly that (the reference's scan script loads our pickles with
its own classes) and load_manual_pickle reads the reference's pickles into the classes below.  All arithmetic runs on
the GPU (ef_fit_manual_host, ef_pca_fit_f64_host, ef_scaler_fit_u8_host, K1 + K2); there is no numpy fallback.
"""
import io
import json
import os
import pickle
import sys
from datetime import datetime

import numpy as np

from . import engine
from ._lib import METRIC_COSINE_G1

_CACHE = {}


class ManualPCA:
    """scripts/manual/train-v2.py:9-50.  components_ [k, D] are the leading eigenvectors of np.cov (the reference leaves
    their signs to LAPACK; here the entry of largest magnitude of every component is positive)."""

    def __init__(self, n_components=50):
        self.n_components = n_components
        self.components_ = None
        self.mean_ = None
        self.explained_variance_ratio_ = None

    def fit(self, X):
        fit = engine.pca_fit_f64(X, self.n_components)
        self.mean_ = fit["pca_mean"]
        self.components_ = fit["components"]
        self.explained_variance_ratio_ = fit["explained_variance_ratio"]
        self._features = fit["features"]
        return self

    def transform(self, X):
        return engine.project_f64(X, self.components_, self.mean_)

    def fit_transform(self, X):
        self.fit(X)
        feats = self.__dict__.pop("_features")
        return feats


class ManualStandardScaler:
    """scripts/manual/train-v2.py:52-72: np.std (population), exact zeros -> 1.  Takes 8-bit pixel data (what the
    reference's loader produces, :127-131)."""

    def __init__(self):
        self.mean_ = None
        self.scale_ = None

    def fit(self, X):
        self.mean_, _, self.scale_ = engine.scaler_fit_u8(X, 1)
        return self

    def transform(self, X):
        return engine.standardize_u8(X, self.mean_, self.scale_)

    def fit_transform(self, X):
        return self.fit(X).transform(X)


class _ManualUnpickler(pickle.Unpickler):
    """Pickles written by the reference's scripts/manual/train-v2.py name their estimators __main__.Manual*."""

    def find_class(self, module, name):
        if name == "ManualPCA":
            return ManualPCA
        if name == "ManualStandardScaler":
            return ManualStandardScaler
        return super().find_class(module, name)


def load_manual_pickle(path_or_bytes):
    """Model dict of scripts/manual/train-v2.py:271-283 (or any Gen-2 pickle) with the Manual* estimators mapped here."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        return _ManualUnpickler(io.BytesIO(path_or_bytes)).load()
    with open(path_or_bytes, "rb") as f:
        return _ManualUnpickler(f).load()


def dump_manual_pickle(model_data, f):
    """pickle.dump with the two estimator classes recorded as __main__.ManualPCA / __main__.ManualStandardScaler, the
    names the reference's scan script (run as __main__, classes re-declared at scan-template-v2.py:9-72) resolves."""
    main = sys.modules["__main__"]
    saved = {}
    for cls in (ManualPCA, ManualStandardScaler):
        saved[cls] = (cls.__module__, cls.__qualname__, getattr(main, cls.__name__, None))
        cls.__module__ = "__main__"
        setattr(main, cls.__name__, cls)
    try:
        pickle.dump(model_data, f)
    finally:
        for cls, (mod, qual, prev) in saved.items():
            cls.__module__ = mod
            if prev is None:
                delattr(main, cls.__name__)
            else:
                setattr(main, cls.__name__, prev)


def recognizer_for(model_data):
    """Device model of a manual-generation model dict: ManualStandardScaler + ManualPCA projection, the manual cosine
    rule (dot / (|a| |b|), zero norm -> 0.0: scan-template-v2.py:244-258) with argmax + labels (:268-296)."""
    hit = _CACHE.get(id(model_data))
    if hit is not None and hit[0] is model_data:
        return hit[1]
    pca, scaler = model_data['pca'], model_data['scaler']
    rec = engine.Recognizer(pca.components_, scaler.mean_, model_data['face_features'], scale=scaler.scale_,
                            pca_mean=pca.mean_, labels=model_data['face_labels'], metric=METRIC_COSINE_G1,
                            basis_is_components=True, with_residual=True)
    if len(_CACHE) >= 16:                                # bounded: drop the oldest device model
        old = next(iter(_CACHE))
        _CACHE.pop(old)[1].close()
    _CACHE[id(model_data)] = (model_data, rec)
    return rec


class FaceTrainer:
    """scripts/manual/train-v2.py:74-330."""

    def __init__(self, n_components=50):
        self.n_components = n_components
        self.pca = ManualPCA(n_components=n_components)
        self.scaler = ManualStandardScaler()
        self.face_features = []
        self.face_labels = []
        self.face_info = []
        self.is_trained = False
        self.mean_face = None
        self.eigenfaces = None
        self.face_shape = (64, 64)
        self.face_images = np.zeros((0, 4096), np.uint8)
        self.person_id_map = {}

    def load_face_images(self, json_path, face_dir):
        """:92-138 -- crops named by image_path (falls back to image_filename inside face_dir on POSIX hosts); gray +
        resize of all crops in ONE K1 launch."""
        import cv2
        from .gen2 import preprocess_images
        print(f"Loading face data from {json_path}")
        with open(json_path, 'r', encoding='utf-8') as f:
            data = json.load(f)
        faces_data = data['faces']
        print(f"Found {len(faces_data)} faces in JSON")
        images, valid = [], []
        for face_info in faces_data:
            cands = [face_info.get('image_path', ''), face_info.get('image_path', '').replace('\\', '/'),
                     os.path.join(face_dir, face_info.get('image_filename', ''))]
            path = next((c for c in cands if c and os.path.exists(c)), None)
            if path is None:
                print(f"Warning: Image {face_info.get('image_path')} not found, skipping...")
                continue
            img = cv2.imread(path)
            if img is None:
                print(f"Warning: Could not read image {path}, skipping...")
                continue
            images.append(img)
            valid.append(face_info)
        print(f"Successfully loaded {len(images)} face images")
        self.face_images = preprocess_images(images, 64) if images else np.zeros((0, 4096), np.uint8)
        self.face_info = valid
        return len(images)

    def assign_labels_interactive(self, person_name):
        """:140-171 -- every face gets the given person's id 0."""
        print("\n=== Face Labeling ===")
        print(f"Labeling all {len(self.face_info)} faces as '{person_name}'...")
        for info in self.face_info:
            info['person_name'] = person_name
            info['person_id'] = 0
        self.face_labels = np.zeros(len(self.face_images), dtype=int)
        self.person_id_map = {person_name: 0}
        return self.face_labels.tolist()

    def train_pca_model(self):
        """:173-208 -- one device call: ManualStandardScaler.fit_transform + ManualPCA.fit_transform."""
        if len(self.face_images) == 0:
            print("Error: No face images loaded!")
            return False
        if len(self.face_labels) == 0:
            print("Error: No face labels assigned!")
            return False
        X = np.asarray(self.face_images)
        print(f"\nTraining PCA model with {len(X)} faces...")
        print(f"Original feature dimension: {X.shape[1]}")
        print(f"Reducing to {self.n_components} components")
        fit = engine.fit_manual(X, self.n_components)
        self.mean_face = fit["mean_face"]
        print(f"Mean face calculated with shape: {self.mean_face.shape}")
        self.scaler.mean_, self.scaler.scale_ = fit["scaler_mean"], fit["scaler_scale"]
        self.pca.mean_, self.pca.components_ = fit["pca_mean"], fit["components"]
        self.pca.explained_variance_ratio_ = fit["explained_variance_ratio"]
        self.eigenfaces = self.pca.components_
        print(f"Generated {len(self.eigenfaces)} eigenfaces")
        print(f"PCA explained variance ratio: {self.pca.explained_variance_ratio_.sum():.3f}")
        print(f"Reduced feature dimension: {fit['features'].shape[1]}")
        self.face_features = fit["features"]
        self.fit_info = fit["info"]
        self.is_trained = True
        return True

    def save_eigenfaces(self, output_dir, person_name):
        """:210-260 -- {person}_mean_face.jpg, {person}_eigenface_XX.jpg (min-max to u8), {person}_model_info.json."""
        if not self.is_trained:
            print("Error: Model not trained yet!")
            return False
        import cv2
        os.makedirs(output_dir, exist_ok=True)
        mean_img = cv2.normalize(self.mean_face.reshape(self.face_shape), None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)
        cv2.imwrite(os.path.join(output_dir, f"{person_name}_mean_face.jpg"), mean_img)
        n_save = min(10, len(self.eigenfaces))
        for i in range(n_save):
            ef = cv2.normalize(self.eigenfaces[i].reshape(self.face_shape), None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)
            cv2.imwrite(os.path.join(output_dir, f"{person_name}_eigenface_{i + 1:02d}.jpg"), ef)
        model_info = {
            'person_name': person_name,
            'training_date': datetime.now().isoformat(),
            'total_faces': len(self.face_images),
            'n_components': self.n_components,
            'explained_variance_ratio': float(self.pca.explained_variance_ratio_.sum()),
            'face_shape': self.face_shape,
            'eigenfaces_saved': n_save,
        }
        with open(os.path.join(output_dir, f"{person_name}_model_info.json"), 'w', encoding='utf-8') as f:
            json.dump(model_info, f, indent=2, ensure_ascii=False)
        return True

    def save_model(self, model_path):
        """:262-290 -- same keys; the estimators are pickled under the reference's class names."""
        if not self.is_trained:
            print("Error: Model not trained yet!")
            return False
        model_data = {
            'pca': self.pca,
            'scaler': self.scaler,
            'face_features': self.face_features,
            'face_labels': self.face_labels,
            'face_info': self.face_info,
            'person_id_map': self.person_id_map,
            'n_components': self.n_components,
            'mean_face': self.mean_face,
            'eigenfaces': self.eigenfaces,
            'face_shape': self.face_shape,
            'training_date': datetime.now().isoformat(),
        }
        with open(model_path, 'wb') as f:
            dump_manual_pickle(model_data, f)
        print(f"Model saved to {model_path}")
        return True

    def load_model(self, model_path):
        """:292-330."""
        if not os.path.exists(model_path):
            print(f"Error: Model file {model_path} not found!")
            return False
        model_data = load_manual_pickle(model_path)
        self.pca = model_data['pca']
        self.scaler = model_data['scaler']
        self.face_features = model_data['face_features']
        self.face_labels = model_data['face_labels']
        self.face_info = model_data['face_info']
        self.person_id_map = model_data['person_id_map']
        self.n_components = model_data['n_components']
        self.mean_face = model_data.get('mean_face', None)
        self.eigenfaces = model_data.get('eigenfaces', None)
        self.face_shape = model_data.get('face_shape', (64, 64))
        self.is_trained = True
        print(f"Model loaded from {model_path}")
        return True


class FaceScanner:
    """scripts/manual/scan-template-v2.py:74-300, the recognition half (the template-matching detector of the same
    class is template.TemplateMatcher / K6)."""

    def __init__(self, model_path, detection_json_path=None):
        self.model_path = model_path
        self.detection_json_path = detection_json_path
        self.model_data = None
        self.detection_data = None
        self.template_image = None
        self.is_loaded = False

    def load_model_and_data(self):
        if not os.path.exists(self.model_path):
            print(f"Error: Model file {self.model_path} not found!")
            return False
        self.model_data = load_manual_pickle(self.model_path)
        print(f"PCA model loaded: {len(self.model_data['face_features'])} faces, "
              f"{len(self.model_data['person_id_map'])} persons")
        if self.detection_json_path:
            if not os.path.exists(self.detection_json_path):
                print(f"Error: Detection JSON {self.detection_json_path} not found!")
                return False
            with open(self.detection_json_path, 'r', encoding='utf-8') as f:
                self.detection_data = json.load(f)
            print(f"Detection data loaded: {len(self.detection_data['faces'])} detected faces")
        self.is_loaded = True
        return True

    def extract_face_features(self, face_img):
        """:206-231 -- gray + resize 64x64 (K1) + scaler.transform + pca.transform (K2 projection)."""
        if not self.is_loaded:
            return None
        rec = recognizer_for(self.model_data)
        face_img = np.asarray(face_img)
        h, w = face_img.shape[:2]
        return rec.recognize_boxes(face_img, [[0, 0, w, h]], 64, 0.0, want_residual=False).features[0]

    def manual_cosine_similarity(self, vector1, vector2):
        """:233-258 -- dot / (|a| |b|), zero norm -> 0.0, on the device (the Gen-1 cosine rule of ef_match_device)."""
        from .gen2 import match_features
        g = np.asarray(vector2, dtype=np.float64)[None, :]
        score, _ = match_features(np.asarray(vector1, dtype=np.float64)[None, :], {'face_features': g},
                                  metric=METRIC_COSINE_G1, cache=False)
        return float(score[0])

    def recognize_face(self, face_features, threshold=0.7):
        """:260-296 -- best manual cosine over the stored features, (person_id, person_name, confidence)."""
        if not self.is_loaded or face_features is None:
            return -1, "unknown", 0.0
        from .gen2 import _label_tuple, match_features
        score, idx = match_features(np.asarray(face_features, dtype=np.float64)[None, :], self.model_data,
                                    metric=METRIC_COSINE_G1)
        return _label_tuple(score[0], idx[0], self.model_data, threshold)

    def recognize_crops(self, frames, boxes, threshold=0.7):
        """Batched extract_face_features + recognize_face for all boxes of a frame / clip: RecognitionResult."""
        rec = recognizer_for(self.model_data)
        return rec.recognize_boxes(frames, boxes, 64, threshold)
