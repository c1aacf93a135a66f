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


_CACHE_MAX = 16


def _cache_put(key, entry):
    """Bounded cache: the oldest entry goes first and its device model is released."""
    while len(_CACHE) >= _CACHE_MAX:
        _CACHE.pop(next(iter(_CACHE)))[1].close()
    _CACHE[key] = entry


def recognizer_for(model_data, n_slices=0):
    """Device model for a Gen-1 model dict (cached per dict object, bounded)."""
    key = (id(model_data), n_slices)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is model_data:
        return hit[1]
    rec = engine.Recognizer(model_data['eigenfaces'], model_data['mean_face'], model_data['projected_data'],
                            metric=METRIC_COSINE_G1, n_slices=n_slices, with_residual=True)
    _cache_put(key, (model_data, rec))
    return rec


def load_face_images(faces_dir):
    """useless/train.py:8-54: every .jpg/.jpeg/.png of the directory (sorted), gray, flattened to float64 rows."""
    import cv2
    print(f"Loading face images from: {faces_dir}")
    image_files = [f for f in os.listdir(faces_dir) if f.lower().endswith(('.jpg', '.jpeg', '.png'))]
    if not image_files:
        raise ValueError(f"No image files found in {faces_dir}")
    face_vectors, filenames = [], []
    for filename in sorted(image_files):
        img = cv2.imread(os.path.join(faces_dir, filename), cv2.IMREAD_GRAYSCALE)
        if img is None:
            print(f"Warning: Could not load image {filename}")
            continue
        face_vectors.append(img.flatten().astype(np.float64))
        filenames.append(filename)
    if not face_vectors:
        raise ValueError("No valid face images could be loaded")
    face_matrix = np.array(face_vectors)
    print(f"Loaded {len(face_vectors)} face images")
    print(f"Face vector dimension: {face_matrix.shape[1]}")
    return face_matrix, filenames


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
    key = ("proj", id(eigenfaces), id(mean_face))
    hit = _CACHE.get(key)
    # both arrays are kept alive by the entry and identity-checked: a recycled id() can never alias another model
    if hit is not None and hit[0] is eigenfaces and hit[2] is mean_face:
        rec = hit[1]
    else:
        ef_arr = np.asarray(eigenfaces)
        rec = engine.Recognizer(ef_arr, mean_face, np.zeros((1, ef_arr.shape[1])), metric=METRIC_COSINE_G1,
                                with_residual=False)
        _cache_put(key, (eigenfaces, rec, mean_face))
    eigenfaces = np.asarray(eigenfaces)
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


def visualize_eigenfaces(eigenfaces, mean_face, output_dir, person_name, n_display=10):
    """useless/train.py:194-223: {person}_mean_face.jpg and {person}_eigenface_01..NN.jpg, min-max normalised to uint8."""
    import cv2
    os.makedirs(output_dir, exist_ok=True)
    face_dim = int(np.sqrt(len(mean_face)))
    mean_img = cv2.normalize(np.asarray(mean_face).reshape(face_dim, face_dim), None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
    mean_face_path = os.path.join(output_dir, f"{person_name}_mean_face.jpg")
    cv2.imwrite(mean_face_path, mean_img)
    print(f"Mean face saved to: {mean_face_path}")
    n_display = min(n_display, eigenfaces.shape[1])
    for i in range(n_display):
        ef_img = cv2.normalize(np.ascontiguousarray(eigenfaces[:, i]).reshape(face_dim, face_dim), None, 0, 255,
                               cv2.NORM_MINMAX).astype(np.uint8)
        cv2.imwrite(os.path.join(output_dir, f"{person_name}_eigenface_{i + 1:02d}.jpg"), ef_img)
    print(f"Saved {n_display} eigenfaces to: {output_dir}")


def train_single_model(faces_dir, person_name, model_dir, version, n_components=50):
    """useless/train.py:225-278: load the crops of one version, fit on the device, write pickle + JSON + JPEGs.
    Errors are reported and turned into False like in the reference."""
    try:
        print(f"\n=== Training {version.upper()} version model ===")
        if not os.path.exists(faces_dir):
            raise FileNotFoundError(f"Faces directory not found: {faces_dir}")
        face_matrix, filenames = load_face_images(faces_dir)
        eigenfaces, mean_face, projected_data, eigenvalues = manual_pca(face_matrix, n_components=n_components)
        model_path = save_pca_model(eigenfaces, mean_face, projected_data, eigenvalues, filenames, person_name, model_dir,
                                    version)
        visualize_eigenfaces(eigenfaces, mean_face, model_dir, f"{person_name}_{version}")
        print(f"\n=== {version.upper()} PCA Training Summary ===")
        print(f"Person: {person_name}")
        print(f"Version: {version}")
        print(f"Training images: {len(filenames)}")
        print(f"Principal components: {eigenfaces.shape[1]}")
        print(f"Face dimensions: {eigenfaces.shape[0]}")
        print(f"Model saved to: {model_path}")
        print(f"Total explained variance: {np.sum(eigenvalues / np.sum(eigenvalues)) * 100:.2f}%")
        return True
    except Exception as e:
        print(f"Error during {version} PCA training: {str(e)}")
        return False


def train_dual_models(base_faces_dir="faces", model_dir="models", person_name="Joseph_Lai", n_components=50):
    """useless/train.py:280-328 (main): the dark and the light model of one person."""
    versions = [{"name": "dark", "dir": os.path.join(base_faces_dir, "Dark_version")},
                {"name": "light", "dir": os.path.join(base_faces_dir, "Light_version")}]
    print("=== Starting PCA Training for Multiple Versions ===")
    print(f"Person: {person_name}")
    success_count = 0
    for v in versions:
        print(f"\n--- Processing {v['name'].upper()} version ---")
        if train_single_model(v["dir"], person_name, model_dir, v["name"], n_components):
            success_count += 1
        else:
            print(f"Failed to train {v['name']} model")
    print("=== FINAL TRAINING SUMMARY ===")
    print(f"Total models trained successfully: {success_count}/{len(versions)}")
    return success_count == len(versions)


def load_dual_pca_models(dark_model_path, light_model_path):
    """useless/scan.py:35-56: both models or (None, None)."""
    dark = load_pca_model(dark_model_path)
    light = load_pca_model(light_model_path)
    if dark is None or light is None:
        print("Error: Could not load both models")
        return None, None
    return dark, light


def _haar_boxes(frame, face_cascade):
    import cv2
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if np.asarray(frame).ndim == 3 else np.asarray(frame)
    faces = face_cascade.detectMultiScale(gray, scaleFactor=1.1, minNeighbors=5, minSize=(30, 30))
    return gray, np.asarray(faces, dtype=np.int32).reshape(-1, 4)


def detect_and_recognize_faces(frame, face_cascade, model_data, similarity_threshold=0.7):
    """useless/scan.py:168-215: host Haar detection; gray ROI -> resize -> project -> cosine for ALL boxes of the frame in
    one K1 + K2 pass.  Returns [(x, y, w, h, name, confidence, recognized)]."""
    gray, boxes = _haar_boxes(frame, face_cascade)
    if len(boxes) == 0:
        return []
    side = int(np.sqrt(model_data['face_dimensions']))
    res = recognizer_for(model_data).recognize_boxes(gray, boxes, side, similarity_threshold, want_features=False)
    return [(int(x), int(y), int(w), int(h), model_data['person_name'], float(s), bool(s >= similarity_threshold))
            for (x, y, w, h), s in zip(boxes, res.score)]


def detect_and_recognize_faces_dual_model(frame, face_cascade, dark_model_data, light_model_data,
                                          similarity_threshold=0.7):
    """useless/scan.py:217-268: the same with the dark / light OR-logic of recognize_face_dual_model (:134-166)."""
    gray, boxes = _haar_boxes(frame, face_cascade)
    if len(boxes) == 0:
        return []
    side = int(np.sqrt(dark_model_data['face_dimensions']))
    d = recognizer_for(dark_model_data).recognize_boxes(gray, boxes, side, similarity_threshold, want_features=False)
    l = recognizer_for(light_model_data).recognize_boxes(gray, boxes, side, similarity_threshold, want_features=False)
    out = []
    for (x, y, w, h), ds, ls in zip(boxes, d.score, l.score):
        is_rec = bool(ds >= similarity_threshold or ls >= similarity_threshold)
        name = dark_model_data['person_name'] if ds >= ls else light_model_data['person_name']
        out.append((int(x), int(y), int(w), int(h), name, float(max(ds, ls)), is_rec))
    return out


def draw_face_annotations(frame, detection_results):
    """useless/scan.py:270-330: red square per detection, label colour by recognition status; detections that are neither
    recognised nor >= 0.3 confident, and boxes smaller than 200 x 200, are skipped."""
    import cv2
    annotated = frame.copy()
    for (x, y, w, h, person_name, confidence, is_recognized) in detection_results:
        if (confidence < 0.3 and not is_recognized) or (w < 200 or h < 200):
            continue
        size = max(w, h)
        sx, sy = x + (w - size) // 2, y + (h - size) // 2
        cv2.rectangle(annotated, (sx, sy), (sx + size, sy + size), (0, 0, 255), 2)
        if is_recognized:
            label_color, label = (255, 255, 0), f"{person_name} ({confidence:.2f})"
        else:
            label_color, label = (0, 0, 255), f"Unknown ({confidence:.2f})"
        ls = cv2.getTextSize(label, cv2.FONT_HERSHEY_SIMPLEX, 0.6, 2)[0]
        cv2.rectangle(annotated, (x, y - ls[1] - 10), (x + ls[0], y), label_color, -1)
        cv2.putText(annotated, label, (x, y - 5), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (255, 255, 255), 2)
    return annotated


def process_video(input_video_path, dark_model_path, light_model_path, output_video_path, similarity_threshold=0.7,
                  max_frames=None):
    """useless/scan.py:332-429: decode -> Haar -> dual-model recognition on the device -> annotated video + statistics.
    Returns the reference's bool; the statistics dict is kept in process_video.last_stats."""
    import cv2
    dark, light = load_dual_pca_models(dark_model_path, light_model_path)
    if dark is None or light is None:
        return False
    face_cascade = cv2.CascadeClassifier(cv2.data.haarcascades + 'haarcascade_frontalface_default.xml')
    if face_cascade.empty():
        print("Error: Could not load face cascade classifier")
        return False
    cap = cv2.VideoCapture(input_video_path)
    if not cap.isOpened():
        print(f"Error: Could not open video file {input_video_path}")
        return False
    fps = int(cap.get(cv2.CAP_PROP_FPS))
    width, height = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    print(f"Processing video: {input_video_path}")
    print(f"Video properties: {width}x{height}, {fps} FPS, {total_frames} frames")
    out = cv2.VideoWriter(output_video_path, cv2.VideoWriter_fourcc(*'mp4v'), fps, (width, height)) if output_video_path else None
    frame_count = 0
    stats = {'recognized': 0, 'unknown': 0, 'total_detections': 0}
    while True:
        ret, frame = cap.read()
        if not ret or (max_frames is not None and frame_count >= max_frames):
            break
        frame_count += 1
        results = detect_and_recognize_faces_dual_model(frame, face_cascade, dark, light, similarity_threshold)
        for r in results:
            stats['total_detections'] += 1
            stats['recognized' if r[6] else 'unknown'] += 1
        if out is not None:
            out.write(draw_face_annotations(frame, results))
        if frame_count % 30 == 0 and total_frames:
            print(f"Progress: {frame_count / total_frames * 100:.1f}% ({frame_count}/{total_frames})")
    cap.release()
    if out is not None:
        out.release()
    print("\n=== Processing Complete ===")
    print(f"Total frames processed: {frame_count}")
    print(f"Total face detections: {stats['total_detections']}")
    print(f"Recognized faces: {stats['recognized']}")
    print(f"Unknown faces: {stats['unknown']}")
    if stats['total_detections'] > 0:
        print(f"Recognition rate: {stats['recognized'] / stats['total_detections'] * 100:.1f}%")
    stats['frames'] = frame_count
    process_video.last_stats = stats
    return True
