#!/usr/bin/env python
"""scripts/manual/train-v2.py of the reference (ManualStandardScaler + ManualPCA trainer) on the B200 engine: same CLI
(--person, default Joseph_Lai), same output files under faces/{person}/."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description="Train a manual-PCA face model for one person")
    ap.add_argument("--person", default="Joseph_Lai")
    ap.add_argument("--components", type=int, default=50)
    args = ap.parse_args()
    face_dir = os.path.join("faces", args.person)
    json_file = os.path.join(face_dir, f"{args.person}_faces_detection.json")
    if not os.path.exists(json_file):
        print(f"Error: JSON file {json_file} not found!")
        print("Please run detection first to generate face detection data.")
        return 1
    trainer = ef.manual.FaceTrainer(n_components=args.components)
    n = trainer.load_face_images(json_file, face_dir)
    if n == 0:
        print("No valid face images loaded!")
        return 1
    trainer.n_components = trainer.pca.n_components = min(args.components, n)
    trainer.assign_labels_interactive(args.person)
    if not trainer.train_pca_model():
        print("Training failed!")
        return 1
    trainer.save_eigenfaces(face_dir, args.person)
    trainer.save_model(os.path.join(face_dir, "face_model.pkl"))
    print("Training completed successfully!")
    return 0


if __name__ == "__main__":
    sys.exit(main())
