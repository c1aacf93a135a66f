#!/usr/bin/env python
"""Drop-in for the reference's run_pipeline.py --video V --person P: detection -> training -> recognition, each stage a
child process of the script next to this file (the reference chains its scripts with subprocess.run the same way and
aborts with exit code 1 when a stage fails).  --live (camera recording + GUI) is host-only and not provided."""
import argparse
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def run_stage(script, args, title):
    print(f"\n{'=' * 60}\n{title}\n{'=' * 60}")
    res = subprocess.run([sys.executable, os.path.join(HERE, script)] + args)
    if res.returncode != 0:
        print(f"\nPipeline aborted: {title} failed")
        sys.exit(1)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Complete face recognition pipeline")
    ap.add_argument("--video", required=True, help="Input video file path")
    ap.add_argument("--person", required=True, help="Person name for organizing output")
    ap.add_argument("--max-frames", type=int, default=None)
    a = ap.parse_args()
    if not os.path.exists(a.video):
        print("\nPipeline aborted: Input video not found")
        sys.exit(1)
    os.makedirs(f"faces/lock_version/{a.person}", exist_ok=True)
    run_stage("detection-v4.py", ["--video", a.video, "--person", a.person], "Face Detection")
    run_stage("train-v4.py", ["--person", a.person], "Model Training")
    extra = ["--max-frames", str(a.max_frames)] if a.max_frames else []
    run_stage("scan-template-v4.py", ["--video", a.video, "--person", a.person] + extra, "Face Recognition")
    out = f"faces/lock_version/{a.person}"
    print("\nPIPELINE COMPLETED SUCCESSFULLY")
    for name in (f"{a.person}_faces_detection.json", "face_model.pkl", "recognition_output.mp4", "recognition_results.json"):
        path = os.path.join(out, name)
        print(f"   {'ok ' if os.path.exists(path) else 'missing'} {name}" + (f" ({os.path.getsize(path):,} bytes)" if os.path.exists(path) else ""))
