"""The callers either side of the hot path, host side only: detection-JSON ingestion, the
per-person training driver of train-v5.py / train-v4.py, and the video recognition loop of scan-template-v*.py with the
device doing gray + resize + projection + match for ALL detections of a frame and ALL loaded models.

Face DETECTION (Haar cascade) and video decode stay on the host with OpenCV, as BASELINE.json's north star says; they
are not re-implemented here.  File names, JSON keys and console messages follow the reference so that its scripts and
ours can be mixed (detection by one, training by the other, recognition by either).

  face_image_files / generate_detection_json_for_person   train-v5.py:33-142, :171-191
  train_person_model / train_all_persons                  train-v5.py:507-610
  train_single_person                                     train-v4.py:258-311 (fixed 50 components, --person)
  detect_faces_and_save_data                              detection-v4.py:8-115 (host Haar, JSON schema :71-108)
  process_video                                           scripts/auto/scan-template-v2.py:329-510 (results JSON :442-502)
"""
import glob
import json
import os
import re
from datetime import datetime

import numpy as np

from . import gen2

_SKIP = ("eigenface", "mean_face", "model_info")


def face_image_files(person_dir):
    """The crop JPEGs of a person directory: everything but the eigenface / mean-face renderings (train-v5.py:185-189)."""
    files = glob.glob(os.path.join(person_dir, "*.jpg"))
    return sorted(f for f in files if not any(s in os.path.basename(f).lower() for s in _SKIP))


def count_face_images(directory):
    """train-v5.py:143-191: a base directory counts all its person sub-directories, a person directory itself."""
    if not os.path.exists(directory):
        return 0
    subdirs = [d for d in sorted(os.listdir(directory)) if os.path.isdir(os.path.join(directory, d))]
    if subdirs and any(face_image_files(os.path.join(directory, d)) for d in subdirs):
        total = 0
        for d in subdirs:
            n = len(face_image_files(os.path.join(directory, d)))
            print(f"{d}: {n} face images")
            total += n
        return total
    return len(face_image_files(directory))


def _frame_number(filename):
    m = re.search(r"face_\d+_frame_(\d+)", filename) or re.search(r"_face_(\d+)", filename)
    return int(m.group(1)) if m else 0


def generate_detection_json_for_person(person_name, face_dir, fps=30.0):
    """Rebuild {person}_faces_detection.json from the crop files (train-v5.py:33-142; same keys, positions unknown)."""
    import cv2
    print(f"Generating detection JSON for {person_name}...")
    files = face_image_files(face_dir)
    print(f"Found {len(files)} face images for {person_name}")
    if not files:
        print(f"No face images found for {person_name}")
        return None
    faces = []
    for face_id, path in enumerate(files):
        name = os.path.basename(path)
        img = cv2.imread(path)
        h, w = img.shape[:2] if img is not None else (64, 64)
        frame = _frame_number(name)
        faces.append({"face_id": face_id, "frame_number": frame, "timestamp": frame / fps, "x": 0, "y": 0,
                      "width": int(w), "height": int(h), "center_x": int(w) // 2, "center_y": int(h) // 2,
                      "area": int(w) * int(h), "image_path": path, "image_filename": name})
    info = {"video_path": f"videos/{person_name}.mp4",
            "total_frames": max(f["frame_number"] for f in faces) + 1, "fps": fps,
            "total_faces_detected": len(faces), "processing_date": datetime.now().isoformat(), "faces": faces}
    json_path = os.path.join(face_dir, f"{person_name}_faces_detection.json")
    with open(json_path, "w", encoding="utf-8") as f:
        json.dump(info, f, indent=2, ensure_ascii=False)
    print(f"Generated {json_path}")
    print(f"Total faces: {len(faces)}")
    return json_path


def train_person_model(person_name, base_dir, n_components=None):
    """train-v5.py:507-568 (n_components = number of crops) or, with n_components given, train-v4.py's fixed size.
    Writes face_model.pkl, multi_person_mean_face.jpg, multi_person_eigenface_XX.jpg, multi_person_model_info.json."""
    print(f"\n=== Training model for {person_name} ===")
    person_dir = os.path.join(base_dir, person_name)
    json_file = os.path.join(person_dir, f"{person_name}_faces_detection.json")
    model_path = os.path.join(person_dir, "face_model.pkl")
    if not os.path.exists(person_dir):
        print(f"Error: Person directory {person_dir} not found!")
        return False
    face_count = count_face_images(person_dir)
    if face_count == 0:
        print(f"No face images found for {person_name}!")
        return False
    if not os.path.exists(json_file):
        generate_detection_json_for_person(person_name, person_dir)
    k = (face_count if face_count > 1 else 1) if n_components is None else min(int(n_components), face_count)
    print(f"Face images found for {person_name}: {face_count}")
    print(f"Setting n_components to: {k}")
    trainer = gen2.MultiFaceTrainer(n_components=k)
    num_faces = trainer.load_face_images_from_json(json_file, person_dir)
    if num_faces == 0:
        print(f"No valid face images loaded for {person_name}!")
        return False
    if trainer.n_components > num_faces:          # unreadable files: sklearn would refuse k > N, so does the engine
        trainer.n_components = num_faces
    print(f"Loaded {num_faces} face images for {person_name}")
    if not trainer.train_pca_model():
        print(f"Training failed for {person_name}!")
        return False
    trainer.save_eigenfaces(person_dir)
    trainer.save_model(model_path)
    print(f"Training completed successfully for {person_name}!")
    print(f"Model saved to: {model_path}")
    print(f"Eigenfaces and mean face saved to: {person_dir}")
    return True


def train_all_persons(base_dir="faces/lock_version"):
    """train-v5.py:570-610: one model per person directory; returns (successes, failures)."""
    if not os.path.exists(base_dir):
        print(f"Error: Base directory {base_dir} not found!")
        return 0, 0
    persons = [d for d in sorted(os.listdir(base_dir)) if os.path.isdir(os.path.join(base_dir, d))]
    if not persons:
        print(f"No person directories found in {base_dir}!")
        return 0, 0
    print(f"Found {len(persons)} person directories: {persons}")
    ok = bad = 0
    for person in persons:
        try:
            if train_person_model(person, base_dir):
                ok += 1
            else:
                bad += 1
        except Exception as e:                       # the reference reports and carries on (train-v5.py:594-601)
            print(f"Error training model for {person}: {e}")
            bad += 1
    print("\n=== Training Summary ===")
    print(f"Total persons processed: {len(persons)}")
    print(f"Successful trainings: {ok}")
    print(f"Failed trainings: {bad}")
    return ok, bad


# ------------------------------------------------------------------------------------------ host detection
def haar_detector():
    import cv2
    cascade = cv2.CascadeClassifier(cv2.data.haarcascades + "haarcascade_frontalface_default.xml")
    if cascade.empty():
        raise RuntimeError("OpenCV Haar cascade haarcascade_frontalface_default.xml not found")
    return cascade


def detect_boxes(cascade, frame_bgr):
    """detection-v4.py:47-55: gray + detectMultiScale(1.1, 5, (30, 30)) on the host; returns int [n, 4] (x, y, w, h)."""
    import cv2
    gray = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2GRAY)
    faces = cascade.detectMultiScale(gray, scaleFactor=1.1, minNeighbors=5, minSize=(30, 30))
    return np.asarray(faces, dtype=np.int32).reshape(-1, 4)


def detect_faces_and_save_data(video_path, output_face_dir, output_json_path):
    """detection-v4.py:8-115: host Haar detection over a video; crops + detection JSON (the training input format)."""
    import cv2
    os.makedirs(output_face_dir, exist_ok=True)
    cascade = haar_detector()
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        print(f"Error: Could not open video {video_path}")
        return None
    fps = cap.get(cv2.CAP_PROP_FPS)
    total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    print(f"Video info: {total_frames} frames, {fps:.2f} FPS")
    faces, frame_count = [], 0
    while True:
        ret, frame = cap.read()
        if not ret:
            break
        for (x, y, w, h) in detect_boxes(cascade, frame):
            face_id = len(faces)
            name = f"face_{face_id:06d}_frame_{frame_count:06d}.jpg"
            path = os.path.join(output_face_dir, name)
            cv2.imwrite(path, frame[y:y + h, x:x + w])
            faces.append({"face_id": face_id, "frame_number": frame_count,
                          "timestamp": frame_count / fps if fps > 0 else 0, "x": int(x), "y": int(y), "width": int(w),
                          "height": int(h), "center_x": int(x + w // 2), "center_y": int(y + h // 2),
                          "area": int(w * h), "image_path": path, "image_filename": name})
        frame_count += 1
        if frame_count % 100 == 0 and total_frames:
            print(f"Progress: {100.0 * frame_count / total_frames:.1f}% ({frame_count}/{total_frames} frames)")
    cap.release()
    info = {"video_path": video_path, "total_frames": total_frames, "fps": fps, "total_faces_detected": len(faces),
            "processing_date": datetime.now().isoformat(), "faces": faces}
    with open(output_json_path, "w", encoding="utf-8") as f:
        json.dump(info, f, indent=2, ensure_ascii=False)
    print("\nDetection completed!")
    print(f"Total faces detected: {len(faces)}")
    print(f"Face images saved to: {output_face_dir}")
    print(f"JSON data saved to: {output_json_path}")
    return info


# ------------------------------------------------------------------------------------------ video recognition
def recognize_frame(scanner, frame_bgr, boxes, threshold=0.8):
    """All detections of one frame against every loaded model in one batch per model (K1 + K2 on the device).
    Returns a list of dicts with the reference's result keys."""
    if len(boxes) == 0:
        return []
    ids, names, confs = scanner.recognize_faces_all_models(frame_bgr, np.asarray(boxes, dtype=np.int32), threshold)
    out = []
    for (x, y, w, h), pid, name, conf in zip(boxes, ids, names, confs):
        out.append({"x": int(x), "y": int(y), "width": int(w), "height": int(h), "person_id": int(pid),
                    "person_name": str(name), "confidence": float(conf), "recognized": bool(conf >= threshold)})
    return out


def process_video(video_path, scanner, output_json=None, output_video=None, threshold=0.8, max_frames=None,
                  rank=0, world=1):
    """Decode (host) -> Haar boxes (host) -> device recognition of every box -> optional annotated video + results JSON
    (schema of scripts/auto/scan-template-v2.py:490-502).  Frames are dealt round-robin over `world` ranks (each rank
    decodes the whole clip but only detects / recognises its own frames): the data-parallel split of config 5."""
    import cv2
    cascade = haar_detector()
    cap = cv2.VideoCapture(video_path)
    if not cap.isOpened():
        print(f"Error: Could not open video {video_path}")
        return None
    fps = cap.get(cv2.CAP_PROP_FPS)
    width, height = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    total_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    writer = None
    if output_video and world == 1:
        writer = cv2.VideoWriter(output_video, cv2.VideoWriter_fourcc(*"mp4v"), fps or 30.0, (width, height))
    results, frame_number, recognised = [], 0, 0
    while True:
        ret, frame = cap.read()
        if not ret or (max_frames is not None and frame_number >= max_frames):
            break
        if frame_number % world == rank:
            dets = recognize_frame(scanner, frame, detect_boxes(cascade, frame), threshold)
            for d in dets:
                d.update({"frame_number": frame_number, "timestamp": frame_number / fps if fps > 0 else 0})
                recognised += d["recognized"]
                if writer is not None:
                    color = (0, 255, 0) if d["recognized"] else (0, 0, 255)
                    cv2.rectangle(frame, (d["x"], d["y"]), (d["x"] + d["width"], d["y"] + d["height"]), color, 2)
                    cv2.putText(frame, f"{d['person_name']} {d['confidence']:.2f}", (d["x"], max(d["y"] - 8, 12)),
                                cv2.FONT_HERSHEY_SIMPLEX, 0.5, color, 1)
            results.extend(dets)
        if writer is not None:
            writer.write(frame)
        frame_number += 1
        if frame_number % 100 == 0 and total_frames:
            print(f"Progress: {100.0 * frame_number / total_frames:.1f}% ({frame_number}/{total_frames} frames)")
    cap.release()
    if writer is not None:
        writer.release()
    # recognition_results.json of the reference (scripts/auto/scan-template-v2.py:442-502): the six top-level keys and, per
    # result, frame_number / timestamp / x / y / width / height / person_id / person_name / confidence /
    # template_match_confidence / ref_frame_diff.  This loop detects with the Haar cascade instead of the reference's
    # template tracker, so the two tracker fields carry their neutral values (1.0: the detector gives no score; 0: the
    # box comes from the frame itself); `recognized` and the run summary are additional keys readers may ignore.
    for d in results:
        d.setdefault("template_match_confidence", 1.0)
        d.setdefault("ref_frame_diff", 0)
    summary = {"video_path": video_path, "total_frames": total_frames if total_frames > 0 else frame_number, "fps": fps,
               "total_recognitions": len(results), "processing_date": datetime.now().isoformat(), "results": results,
               "summary": {"frames_processed": frame_number, "threshold": threshold,
                           "recognized_detections": int(recognised),
                           "recognition_rate": (recognised / len(results)) if results else 0.0,
                           "models": sorted(scanner.models), "rank": rank, "world_size": world}}
    if output_json:
        with open(output_json, "w", encoding="utf-8") as f:
            json.dump(summary, f, indent=2, ensure_ascii=False)
        print(f"Results saved to: {output_json}")
    print(f"Total detections: {len(results)}, recognized: {recognised}")
    return summary
