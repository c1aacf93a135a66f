#!/usr/bin/env python
"""Recognition entry point: loads every faces/lock_version/*/face_model.pkl (the reference's or ours) and recognises
every detection of every frame against all models on the GPU.  The reference's scan-template-v4.py is a live-camera GUI
loop; this shim runs its per-frame logic over a video FILE (--video):
  --detector template   the reference's own detector: template matching of every person's first five face crops at three
                        scales (K6 on the GPU), PCA verification, the reference's final-name rules; prints the result
                        dicts like process_live_camera and writes them to template_recognition_results.json
  --detector haar       (default) Haar detections + recognition, recognition_results.json / recognition_output.mp4 like
                        scan-template-v2.py."""
import argparse
import os

import _bootstrap  # noqa: F401
from eigenfaces_b200 import gen2, pipeline

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Multi-model face recognition over a video file")
    ap.add_argument("--video", required=True, help="Input video file path")
    ap.add_argument("--person", default=None, help="Write outputs into faces/lock_version/{person}/")
    ap.add_argument("--threshold", type=float, default=0.8)
    ap.add_argument("--max-frames", type=int, default=None)
    ap.add_argument("--no-video-output", action="store_true")
    ap.add_argument("--detector", choices=("haar", "template"), default="haar")
    args = ap.parse_args()
    scanner = gen2.MultiModelFaceScanner()
    if not scanner.load_all_models("faces/lock_version/*/face_model.pkl"):
        print("No models loaded. Please train models first using train-v5.py")
        raise SystemExit(1)
    out_dir = f"faces/lock_version/{args.person}" if args.person else "."
    os.makedirs(out_dir, exist_ok=True)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        import torch
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    suffix = f".rank{rank}" if world > 1 else ""
    if args.detector == "template":
        import json
        results = scanner.process_video_template(args.video, args.max_frames, rank, world)
        if results is None:
            raise SystemExit(1)
        with open(os.path.join(out_dir, f"template_recognition_results{suffix}.json"), "w", encoding="utf-8") as f:
            json.dump(results, f, indent=2)
        print(f"\nLive recognition stopped! {len(results)} result(s)")
        raise SystemExit(0)
    res = pipeline.process_video(args.video, scanner, os.path.join(out_dir, f"recognition_results{suffix}.json"),
                                 None if args.no_video_output else os.path.join(out_dir, "recognition_output.mp4"),
                                 args.threshold, args.max_frames, rank, world)
    raise SystemExit(0 if res is not None else 1)
