#!/usr/bin/env python
"""useless/scan.py of the reference (dual-model recognition of a video file) on the B200 engine: host decode + Haar,
device recognition of every detection against the dark and the light model, annotated output video."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--video", default="videos/test.mp4")
    ap.add_argument("--dark-model", default="models/Joseph_Lai_dark_pca_model.pkl")
    ap.add_argument("--light-model", default="models/Joseph_Lai_light_pca_model.pkl")
    ap.add_argument("--output", default="videos/output_recognition.mp4")
    ap.add_argument("--threshold", type=float, default=0.8)
    a = ap.parse_args()
    sys.exit(0 if ef.gen1.process_video(a.video, a.dark_model, a.light_model, a.output, a.threshold) else 1)
