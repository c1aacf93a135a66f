"""Gen-2 ("sklearn PCA") interface of the reference, backed by the B200 engine.

  MultiFaceTrainer        train-v5.py:13-505   (train_pca_model :349-385, save_model :438-467, load_model :469-505)
  MultiModelFaceScanner   scan-template-v4.py:10-319 (extract_face_features :253-268, recognize_face_with_model
                          :270-287, recognize_face_all_models :289-319)
The pickles written by save_model hold genuine fitted sklearn PCA / StandardScaler objects, so the reference's own
scan-template-v4.py loads them and calls .transform on them; the pickles the reference wrote load here.
"""
import glob
import json
import os
import pickle
from datetime import datetime

import numpy as np

from . import engine, template
from ._lib import METRIC_COSINE_SK

_CACHE = {}


def _sklearn_objects(fit, n_samples):
    """Populate real sklearn estimators with the device fit (what pickle consumers call .transform on)."""
    from sklearn.decomposition import PCA
    from sklearn.preprocessing import StandardScaler
    k, D = fit["components"].shape
    pca = PCA(n_components=k)
    pca.components_ = fit["components"]
    pca.mean_ = fit["pca_mean"]
    pca.explained_variance_ = fit["explained_variance"]
    pca.explained_variance_ratio_ = fit["explained_variance_ratio"]
    pca.singular_values_ = fit["singular_values"]
    pca.n_components_ = k
    pca.n_samples_ = n_samples
    pca.n_features_in_ = D
    pca.noise_variance_ = fit["noise_variance"]
    pca._fit_svd_solver = "full"
    scaler = StandardScaler()
    scaler.mean_ = fit["scaler_mean"]
    scaler.var_ = fit["scaler_var"]
    scaler.scale_ = fit["scaler_scale"]
    scaler.n_samples_seen_ = np.int64(n_samples)
    scaler.n_features_in_ = D
    return pca, scaler


def _model_arrays(model_data):
    pca = model_data.get('pca', model_data.get('pca_model'))   # the shipped pickle uses 'pca_model' 
    scaler = model_data['scaler']
    return pca, scaler


_CACHE_MAX = 32


def _cache_put(key, entry):
    """Bounded cache (a long-running scanner reloads its models: load_all_models unpickles new dicts every time): the
    oldest entry goes first and its device model is released."""
    while len(_CACHE) >= _CACHE_MAX:
        old = _CACHE.pop(next(iter(_CACHE)))
        if len(old) > 1 and hasattr(old[1], "close"):
            old[1].close()
    _CACHE[key] = entry


def recognizer_for(model_data, n_slices=0):
    """Device model for a Gen-2 model dict (cached per dict object, bounded)."""
    key = (id(model_data), n_slices)
    hit = _CACHE.get(key)
    if hit is not None and hit[0] is model_data:
        return hit[1]
    pca, scaler = _model_arrays(model_data)
    rec = engine.Recognizer(pca.components_, scaler.mean_, model_data['face_features'], scale=scaler.scale_,
                            pca_mean=pca.mean_, labels=model_data['face_labels'], metric=METRIC_COSINE_SK,
                            basis_is_components=True, n_slices=n_slices, with_residual=True)
    _cache_put(key, (model_data, rec))
    return rec


class MultiFaceTrainer:
    def __init__(self, n_components=50):
        self.n_components = n_components
        self.pca = None
        self.scaler = None
        self.face_images = []
        self.face_features = []
        self.face_labels = []
        self.face_info = []
        self.is_trained = False
        self.mean_face = None
        self.eigenfaces = None
        self.face_shape = (64, 64)
        self.person_id_map = {}
        self.fit_info = None

    def load_face_images_from_json(self, json_path, person_dir):
        """train-v5.py:276-347: host JPEG decode, then gray + resize of all crops in one K1 launch."""
        import cv2
        print(f"Loading face data from {json_path}")
        if not os.path.exists(json_path):
            print(f"Error: JSON file {json_path} not found!")
            return 0
        with open(json_path, 'r', encoding='utf-8') as f:
            data = json.load(f)
        faces_data = data['faces']
        print(f"Found {len(faces_data)} faces in JSON")
        images, valid = [], []
        for i, face_info in enumerate(faces_data):
            if 'image_filename' in face_info:
                image_path = os.path.join(person_dir, face_info['image_filename'])
            elif 'image_path' in face_info:
                image_path = face_info['image_path']
            else:
                image_path = os.path.join(
                    person_dir, f"face_{face_info.get('face_id', i)}_frame_{face_info.get('frame_number', i)}.jpg")
            if not os.path.exists(image_path):
                print(f"Warning: Image {image_path} not found, skipping...")
                continue
            img = cv2.imread(image_path)
            if img is None:
                print(f"Warning: Could not read image {image_path}, skipping...")
                continue
            images.append(img)
            valid.append(face_info)
        self.face_images = preprocess_images(images, self.face_shape[0]) if images else np.zeros((0, 4096), np.uint8)
        print(f"Successfully loaded {len(images)} face images")
        self.face_info = valid
        self.face_labels = np.zeros(len(images), dtype=int)
        self.person_id_map = {os.path.basename(person_dir): 0}
        return len(images)

    def train_pca_model(self):
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
        fit = engine.fit_gen2(X, self.n_components)
        self.mean_face = fit["mean_face"]
        print(f"Mean face calculated with shape: {self.mean_face.shape}")
        self.pca, self.scaler = _sklearn_objects(fit, len(X))
        self.eigenfaces = self.pca.components_
        print(f"Generated {len(self.eigenfaces)} eigenfaces")
        print(f"PCA explained variance ratio: {self.pca.explained_variance_ratio_.sum():.3f}")
        print(f"Reduced feature dimension: {fit['features'].shape[1]}")
        self.face_features = fit["features"]
        self.fit_info = fit["info"]
        self.is_trained = True
        return True

    def save_eigenfaces(self, output_dir):
        """train-v5.py:387-436: mean face + first 10 eigenfaces as min-max normalised JPEGs + model-info JSON."""
        if not self.is_trained:
            print("Error: Model not trained yet!")
            return False
        import cv2
        os.makedirs(output_dir, exist_ok=True)
        mean_img = cv2.normalize(self.mean_face.reshape(self.face_shape), None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)
        cv2.imwrite(os.path.join(output_dir, "multi_person_mean_face.jpg"), mean_img)
        n_save = min(10, len(self.eigenfaces))
        for i in range(n_save):
            ef = cv2.normalize(self.eigenfaces[i].reshape(self.face_shape), None, 0, 255, cv2.NORM_MINMAX, dtype=cv2.CV_8U)
            cv2.imwrite(os.path.join(output_dir, f"multi_person_eigenface_{i + 1:02d}.jpg"), ef)
        model_info = {
            'training_date': datetime.now().isoformat(),
            'total_faces': len(self.face_images),
            'total_persons': len(self.person_id_map),
            'person_id_map': self.person_id_map,
            'n_components': self.n_components,
            'explained_variance_ratio': float(self.pca.explained_variance_ratio_.sum()),
            'face_shape': self.face_shape,
            'eigenfaces_saved': n_save,
        }
        with open(os.path.join(output_dir, "multi_person_model_info.json"), 'w', encoding='utf-8') as f:
            json.dump(model_info, f, indent=2, ensure_ascii=False)
        return True

    def save_model(self, model_path):
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
            pickle.dump(model_data, f)
        print(f"Model saved to {model_path}")
        return True

    def load_model(self, model_path):
        if not os.path.exists(model_path):
            print(f"Error: Model file {model_path} not found!")
            return False
        with open(model_path, 'rb') as f:
            model_data = pickle.load(f)
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


def preprocess_images(images, side):
    """Gray + resize a list of variable-size uint8 images (BGR or gray) to [B, side*side] with ONE K1 launch:
    the images are packed into one padded canvas per image (host), boxes name the valid region."""
    import torch
    from . import engine as _e
    H = max(im.shape[0] for im in images)
    W = max(im.shape[1] for im in images)
    color = any(im.ndim == 3 for im in images)
    canvas = np.zeros((len(images), H, W, 3) if color else (len(images), H, W), dtype=np.uint8)
    boxes = np.zeros((len(images), 5), dtype=np.int32)
    for i, im in enumerate(images):
        if color and im.ndim == 2:
            raise ValueError("mixing gray and BGR images in one batch is not supported")
        canvas[i, :im.shape[0], :im.shape[1]] = im
        boxes[i] = (i, 0, 0, im.shape[1], im.shape[0])
    dev = torch.device("cuda", torch.cuda.current_device())
    out = _e.preprocess_device(torch.from_numpy(canvas).to(dev), torch.from_numpy(boxes).to(dev), side)
    return out[:, :side * side].cpu().numpy()


class MultiModelFaceScanner:
    def __init__(self):
        self.models = {}

    def load_all_models(self, model_pattern="faces/lock_version/*/face_model.pkl"):
        """scan-template-v4.py:17-74: models, detection JSON and the first five face crops of every person as
        template images.  The reference opens face_data['image_path'] as written by detection-v4.py (Windows
        separators); here the same path is tried first, then with normalised separators, then the file name inside
        the person's directory (train-v5.py:305-306's rule) so that the templates also load on POSIX hosts."""
        model_paths = glob.glob(model_pattern)
        if not model_paths:
            print(f"No models found matching pattern: {model_pattern}")
            return False
        print(f"Found {len(model_paths)} model(s):")
        for model_path in model_paths:
            person_name = os.path.basename(os.path.dirname(model_path))
            try:
                with open(model_path, 'rb') as f:
                    model_data = pickle.load(f)
                person_dir = os.path.dirname(model_path)
                detection_json_path = os.path.join(person_dir, f"{person_name}_faces_detection.json")
                detection_data = None
                if os.path.exists(detection_json_path):
                    with open(detection_json_path, 'r', encoding='utf-8') as f:
                        detection_data = json.load(f)
                template_images = []
                if detection_data and detection_data.get('faces'):
                    import cv2
                    for face_data in detection_data['faces'][:5]:
                        cands = [face_data.get('image_path', ''), face_data.get('image_path', '').replace('\\', '/'),
                                 os.path.join(person_dir, face_data.get('image_filename', ''))]
                        path = next((c for c in cands if c and os.path.exists(c)), None)
                        if path is None:
                            continue
                        template_img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
                        if template_img is not None:
                            template_images.append({'image': template_img, 'width': face_data['width'],
                                                    'height': face_data['height']})
                self.models[person_name] = {'model_data': model_data, 'detection_data': detection_data,
                                            'template_images': template_images, 'model_path': model_path}
                print(f"  - {person_name}: {len(model_data['face_features'])} faces")
            except Exception as e:
                print(f"  - Failed to load {person_name}: {e}")
        print(f"Successfully loaded {len(self.models)} model(s)")
        return len(self.models) > 0

    # ---- template-matching detector (scan-template-v4.py:75-251) on the device
    def is_detection_in_corner(self, detection, frame_width, frame_height, corner_threshold=0.15,
                               border_threshold=0.05):
        return template.is_detection_in_corner(detection, frame_width, frame_height, corner_threshold, border_threshold)

    def non_max_suppression(self, detections, overlap_threshold=0.3):
        return template.non_max_suppression(detections, overlap_threshold)

    def calculate_overlap(self, det1, det2):
        return template.calculate_overlap(det1, det2)

    def template_match_all_models(self, frame):
        """scan-template-v4.py:129-197: every template image of every person at the scales 0.8 / 1.0 / 1.2 against the
        gray frame -- ONE batch of device launches for all of them (K6) instead of a matchTemplate call each; the
        per-person selection (strict >, corner / border filter, 0.6 threshold) follows the reference's loop order."""
        frame = np.asarray(frame)
        frame_height, frame_width = frame.shape[:2]
        persons = [(name, info) for name, info in self.models.items()
                   if info.get('template_images') and info.get('detection_data')]
        if not persons:
            return []
        key = tuple((name, len(info['template_images'])) for name, info in persons)
        if getattr(self, '_tm_key', None) != key:
            images = [t['image'] for _, info in persons for t in info['template_images']]
            self._tm = template.TemplateMatcher(images)
            self._tm_key = key
        results = self._tm.match(frame)
        detected_faces = []
        first = 0
        for name, info in persons:
            n_t = len(info['template_images'])
            mine = [r for r, (ti, _, _, _) in zip(results, self._tm.jobs) if first <= ti < first + n_t]
            first += n_t
            best = template.select_best_match(name, mine, frame_width, frame_height)
            if best:
                detected_faces.append(best)
        return detected_faces

    def recognize_frame_template(self, frame, frame_count=0):
        """The per-frame body of process_live_camera (scan-template-v4.py:342-424) without the GUI: template-matching
        detection on the gray frame (K6), the size + PCA-confidence selection when several persons' templates fire
        (:352-377), PCA verification of the detection through every model (K1 + K2) and the final-name rules
        (:393-401).  Returns the list of result dicts the reference appends to recognition_results (:411-419)."""
        import cv2
        frame = np.asarray(frame)
        gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY) if frame.ndim == 3 else frame
        detected_faces = self.template_match_all_models(gray)
        if not detected_faces:
            return []
        # PCA verification of every candidate in one batch (the reference calls recognize_face_all_models per box)
        boxes = [[d['x'], d['y'], d['width'], d['height']] for d in detected_faces]
        _, names, confs = self.recognize_faces_all_models(frame, boxes)
        for d, n, c in zip(detected_faces, names, confs):
            d['pca_person_name'], d['pca_confidence'] = n, float(c)
        if len(detected_faces) > 1:
            best_detection, best_score = None, -1
            for d in detected_faces:
                normalized_size = min(d['width'] * d['height'] / (200 * 200), 1.0)
                combined_score = normalized_size * 0.5 + d['pca_confidence'] * 0.5
                if combined_score > best_score:
                    best_score, best_detection = combined_score, d
            detected_faces = [best_detection] if best_detection else []
        results = []
        for d in detected_faces:
            person_name, template_confidence = d['person_name'], d['confidence']
            pca_person_name, pca_confidence = d['pca_person_name'], d['pca_confidence']
            if pca_person_name == person_name or pca_confidence < 0.5:
                final_person_name, final_confidence = person_name, template_confidence
            else:
                final_person_name, final_confidence = pca_person_name, pca_confidence
            if pca_confidence < 0.8 or template_confidence < 0.7:
                final_person_name = "unknown"
            results.append({'frame_number': frame_count, 'person_name': final_person_name,
                            'template_confidence': template_confidence, 'pca_confidence': pca_confidence,
                            'final_confidence': final_confidence, 'x': d['x'], 'y': d['y'], 'width': d['width'],
                            'height': d['height']})
        return results

    def process_video_template(self, video_path, max_frames=None, rank=0, world=1):
        """process_live_camera's loop over a video FILE instead of camera 0 (no preview window); frames are dealt
        round-robin over `world` ranks.  Returns the recognition_results list (scan-template-v4.py:337,420,437)."""
        import cv2
        if not self.models:
            print("Error: No models loaded!")
            return None
        cap = cv2.VideoCapture(video_path)
        if not cap.isOpened():
            print(f"Error: Could not open video {video_path}")
            return None
        recognition_results = []
        frame_count = 0
        while True:
            ret, frame = cap.read()
            if not ret or (max_frames is not None and frame_count >= max_frames):
                break
            if frame_count % world == rank:
                for result in self.recognize_frame_template(frame, frame_count):
                    print(result)
                    recognition_results.append(result)
            frame_count += 1
        cap.release()
        return recognition_results

    # ---- single-crop interface (reference signatures)
    def extract_face_features(self, face_img, model_data):
        rec = recognizer_for(model_data)
        face_img = np.asarray(face_img)
        h, w = face_img.shape[:2]
        res = rec.recognize_boxes(face_img, [[0, 0, w, h]], 64, 0.0, want_residual=False)
        return res.features[0]

    def recognize_face_with_model(self, face_features, model_data, threshold=0.7):
        """Cosine argmax of already extracted features against the model's gallery (device match kernel)."""
        score, idx = match_features(np.asarray(face_features, dtype=np.float64)[None, :], model_data)
        return _label_tuple(score[0], idx[0], model_data, threshold)

    def recognize_face_all_models(self, face_img, threshold=0.8):
        ids, names, confs = self.recognize_faces_all_models(np.asarray(face_img), None, threshold)
        return ids[0], names[0], confs[0]

    # ---- batched interface
    def recognize_faces_all_models(self, frames, boxes, threshold=0.8):
        """All detections of a frame (or clip) against every loaded model; keeps the best confidence per detection
        with the reference's rules (strict >, first model wins ties, below-threshold name falls back to the model's
        person; scan-template-v4.py:297-319).  boxes None = the whole image is one crop."""
        frames = np.ascontiguousarray(frames)
        if boxes is None:
            h, w = frames.shape[:2]
            boxes = [[0, 0, w, h]]
        B = len(boxes)
        best_conf = np.zeros(B)
        best_id = np.full(B, -1, dtype=np.int64)
        best_name = np.array(["unknown"] * B, dtype=object)
        if B == 0:
            return [], [], []
        # K1 once for all models (every model of the reference uses the same 64 x 64 gray crop, scan-template-v4.py:262)
        # and every model's K2 in ONE C-ABI call: frames and boxes go up once, all scores and labels come back with one
        # copy and one synchronisation -- the reference's per-model loop (scan-template-v4.py:297-314) costs one round trip
        launched = []
        for person_name, info in self.models.items():
            model_data = info['model_data']
            if model_data is None:
                continue
            try:
                launched.append((person_name, model_data, recognizer_for(model_data), None))
            except Exception as e:
                print(f"Error recognizing with model {person_name}: {e}")
        if not launched:
            return best_id.tolist(), best_name.tolist(), best_conf.tolist()
        scores, _, labels_all = engine.recognize_boxes_all_models([r for _, _, r, _ in launched], frames, boxes, 64, threshold)
        for i, (person_name, model_data, rec, _) in enumerate(launched):
            score, label = scores[i], labels_all[i]
            names = _names_for(label, model_data)
            better = score > best_conf
            best_conf = np.where(better, score, best_conf)
            best_id = np.where(better, label, best_id)
            best_name = np.where(better, np.where(names == "unknown", person_name, names), best_name)
        return best_id.tolist(), best_name.tolist(), best_conf.tolist()


def match_features(features, model_data, metric=METRIC_COSINE_SK, cache=True):
    """features [B,k] float64 -> (score [B], index [B]) with the cosine argmax of `metric` on the device (sklearn rule by
    default; METRIC_COSINE_G1 = dot / (|a| |b|) with zero norm -> 0.0, the manual scripts' rule)."""
    import ctypes as C
    import torch
    from . import _lib
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    key = ("gal", id(model_data), metric)
    hit = _CACHE.get(key) if cache else None
    gal = np.asarray(model_data['face_features'], dtype=np.float64)
    n, k = gal.shape
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if hit is None or hit[0] is not model_data:
        g = torch.from_numpy(np.ascontiguousarray(gal)).to(dev)
        gp = torch.empty_like(g)
        gn = torch.zeros(n, dtype=torch.float64, device=dev)
        _lib.check(L.ef_gallery_prepare_device(g.data_ptr(), k, n, k, metric, gp.data_ptr(), k, gn.data_ptr(), stream),
                   "ef_gallery_prepare_device")
        hit = (model_data, gp, gn)
        if cache:
            _cache_put(key, hit)
    gp, gn = hit[1], hit[2]
    p = torch.from_numpy(np.ascontiguousarray(features, dtype=np.float64)).to(dev)
    B = p.shape[0]
    score = torch.empty(B, dtype=torch.float64, device=dev)
    index = torch.empty(B, dtype=torch.int64, device=dev)
    _lib.check(L.ef_match_device(p.data_ptr(), k, B, k, gp.data_ptr(), k, gn.data_ptr(), n, 0, metric,
                                 score.data_ptr(), index.data_ptr(), None, stream), "ef_match_device")
    return score.cpu().numpy(), index.cpu().numpy()


def _names_for(labels, model_data):
    inv = {pid: name for name, pid in reversed(list(model_data['person_id_map'].items()))}
    return np.array([inv.get(int(l), "unknown") if l >= 0 else "unknown" for l in labels], dtype=object)


def _label_tuple(score, idx, model_data, threshold):
    if score >= threshold:
        person_id = model_data['face_labels'][idx]
        person_name = "unknown"
        for name, pid in model_data['person_id_map'].items():
            if pid == person_id:
                person_name = name
                break
        return person_id, person_name, score
    return -1, "unknown", score
