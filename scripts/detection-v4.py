#!/usr/bin/env python
"""Drop-in for the reference's detection-v4.py: host-side Haar detection over a video (out of the GPU scope by design),
writing the crop JPEGs and {person}_faces_detection.json that the trainers ingest."""
import argparse
import os

import _bootstrap  # noqa: F401
from eigenfaces_b200 import pipeline

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Detect faces in video and save face images and data")
    ap.add_argument("--video", required=True, help="Input video file path")
    ap.add_argument("--person", required=True, help="Person name for organizing output directory")
    args = ap.parse_args()
    if not os.path.exists(args.video):
        print(f"Error: Video file {args.video} not found!")
        raise SystemExit(1)
    out_dir = f"faces/lock_version/{args.person}"
    info = pipeline.detect_faces_and_save_data(args.video, out_dir, f"{out_dir}/{args.person}_faces_detection.json")
    raise SystemExit(0 if info is not None else 1)
