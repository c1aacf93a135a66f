"""Gen-1 ("manual PCA") interface of the reference, backed by the B200 engine.

Same names, argument meaning, return values and error behaviour as the reference functions:
  manual_pca                  useless/train.py:56-128
  save_pca_model              useless/train.py:130-192   (pickle dict + *_model_info.json)
  load_pca_model              useless/scan.py:9-33       (prints and returns None on error)
  project_face_to_eigenspace  useless/scan.py:80-98
  recognize_face              useless/scan.py:100-132
  recognize_face_dual_model   useless/scan.py:134-166
plus batched forms (recognize_faces, recognize_faces_dual_model) that take all crops of a frame / clip at once.
The arithmetic runs on the GPU only; there is no numpy fallback.
"""
import json
import os
import pickle
from datetime import datetime

import numpy as np

from . import engine
from ._lib import METRIC_COSINE_G1

_CACHE = {}


def _as_u8(face_vectors, D):
    v = np.asarray(face_vectors)
    if v.ndim == 1:
        v = v[None, :]
    if v.ndim != 2 or v.shape[1] != D:
        raise ValueError(f"face vector(s) must have {D} pixels, got shape {np.asarray(face_vectors).shape}")
    if v.dtype != np.uint8:
        r = np.rint(v)
        if not (np.array_equal(r, v) and r.min() >= 0 and r.max() <= 255):
            raise ValueError("face vectors must hold 8-bit pixel values (the reference feeds flattened uint8 crops)")
        v = r.astype(np.uint8)
    return np.ascontiguousarray(v)


def recognizer_for(model_data, n_slices=0):
    """Device model for a Gen-1 model dict (cached per dict object)."""
    key = (id(model_data), n_slices)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is model_data:
        return hit[1]
    rec = engine.Recognizer(model_data['eigenfaces'], model_data['mean_face'], model_data['projected_data'],
                            metric=METRIC_COSINE_G1, n_slices=n_slices, with_residual=True)
    _CACHE[key] = (model_data, rec)
    return rec


def manual_pca(data_matrix, n_components=None):
    """PCA by the snapshot method on the GPU.  Returns (eigenfaces [D,k], mean_face [D], projected [N,k], eigenvalues [k])."""
    print("Starting manual PCA computation...")
    eigenfaces, mean_face, projected, eigenvalues, info = engine.fit_gen1(data_matrix, n_components)
    print(f"Mean face calculated, shape: {mean_face.shape}")
    print(f"Selected {eigenfaces.shape[1]} principal components")
    print(f"Explained variance ratio: {eigenvalues[:5] / np.sum(eigenvalues)}")
    print(f"PCA computation completed ({info['gpu_ms']:.2f} ms on device, {info['sweeps']} Jacobi sweeps)")
    print(f"Eigenfaces shape: {eigenfaces.shape}")
    print(f"Projected data shape: {projected.shape}")
    return eigenfaces, mean_face, projected, eigenvalues


def save_pca_model(eigenfaces, mean_face, projected_data, eigenvalues, filenames, person_name, model_dir, version=None):
    """Write {person}[_{version}]_pca_model.pkl and _model_info.json with the reference's schema."""
    os.makedirs(model_dir, exist_ok=True)
    stem = f"{person_name}_{version}" if version else person_name
    model_data = {
        'eigenfaces': eigenfaces,
        'mean_face': mean_face,
        'projected_data': projected_data,
        'eigenvalues': eigenvalues,
        'training_filenames': filenames,
        'person_name': person_name,
        'version': version,
        'training_timestamp': datetime.now().isoformat(),
        'n_components': eigenfaces.shape[1],
        'face_dimensions': eigenfaces.shape[0],
    }
    model_path = os.path.join(model_dir, f"{stem}_pca_model.pkl")
    with open(model_path, 'wb') as f:
        pickle.dump(model_data, f)
    info = {
        'person_name': person_name,
        'version': version,
        'training_timestamp': model_data['training_timestamp'],
        'n_components': int(eigenfaces.shape[1]),
        'face_dimensions': int(eigenfaces.shape[0]),
        'n_training_images': len(filenames),
        'explained_variance_ratio': (eigenvalues[:10] / np.sum(eigenvalues)).tolist(),
        'model_file': f"{stem}_pca_model.pkl",
    }
    info_path = os.path.join(model_dir, f"{stem}_model_info.json")
    with open(info_path, 'w') as f:
        json.dump(info, f, indent=2)
    print(f"PCA model saved to: {model_path}")
    print(f"Model info saved to: {info_path}")
    return model_path


def load_pca_model(model_path):
    """Load a pickled Gen-1 model; prints and returns None on any error, like the reference."""
    try:
        with open(model_path, 'rb') as f:
            model_data = pickle.load(f)
        print(f"PCA model loaded successfully from: {model_path}")
        print(f"Person: {model_data['person_name']}")
        print(f"Number of components: {model_data['n_components']}")
        print(f"Face dimensions: {model_data['face_dimensions']}")
        return model_data
    except Exception as e:
        print(f"Error loading PCA model: {str(e)}")
        return None


def project_face_to_eigenspace(face_vector, eigenfaces, mean_face):
    """(face - mean) . eigenfaces on the GPU; face_vector [D] -> [k] (or [B,D] -> [B,k])."""
    eigenfaces = np.asarray(eigenfaces)
    key = ("proj", id(eigenfaces), id(mean_face))
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is eigenfaces:
        rec = hit[1]
    else:
        rec = engine.Recognizer(eigenfaces, mean_face, np.zeros((1, eigenfaces.shape[1])), metric=METRIC_COSINE_G1,
                                with_residual=False)
        _CACHE[key] = (eigenfaces, rec)
    x = _as_u8(face_vector, eigenfaces.shape[0])
    feats = rec.recognize(x, 0.0, want_features=True, want_residual=False).features
    return feats[0] if np.asarray(face_vector).ndim == 1 else feats


def recognize_faces(face_vectors, model_data, similarity_threshold=0.7):
    """Batched recognize_face.  Returns (person_name, max_similarity [B], is_recognized [B], result)."""
    rec = recognizer_for(model_data)
    res = rec.recognize(_as_u8(face_vectors, rec.D), similarity_threshold)
    return model_data['person_name'], res.score, res.score >= similarity_threshold, res


def recognize_face(face_vector, model_data, similarity_threshold=0.7):
    """Returns (person_name, max_similarity, is_recognized) for one flattened crop."""
    name, sims, rec, _ = recognize_faces(face_vector, model_data, similarity_threshold)
    return name, sims[0], bool(rec[0])


def recognize_faces_dual_model(face_vectors, dark_model_data, light_model_data, similarity_threshold=0.7):
    """Batched OR-logic over the dark and light models.  Returns (names, best, is_recognized, dark_sim, light_sim)."""
    dark_name, dark_sim, dark_rec, _ = recognize_faces(face_vectors, dark_model_data, similarity_threshold)
    light_name, light_sim, light_rec, _ = recognize_faces(face_vectors, light_model_data, similarity_threshold)
    is_recognized = dark_rec | light_rec
    best = np.maximum(dark_sim, light_sim)
    names = np.where(dark_sim >= light_sim, dark_name, light_name)
    return names, best, is_recognized, dark_sim, light_sim


def recognize_face_dual_model(face_vector, dark_model_data, light_model_data, similarity_threshold=0.7):
    """Returns (person_name, best_confidence, is_recognized, dark_similarity, light_similarity)."""
    names, best, rec, ds, ls = recognize_faces_dual_model(face_vector, dark_model_data, light_model_data,
                                                          similarity_threshold)
    return str(names[0]), best[0], bool(rec[0]), ds[0], ls[0]
