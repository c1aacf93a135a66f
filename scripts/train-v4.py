#!/usr/bin/env python
"""Drop-in for the reference's train-v4.py (--person P, 50 components; invoked by run_pipeline.py).  The reference lets
sklearn pick the randomized solver here (non-deterministic, random_state=None); the engine always computes the exact
decomposition, so repeated runs give identical models."""
import argparse

import _bootstrap  # noqa: F401
from eigenfaces_b200 import pipeline

if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Train face recognition model with eigenfaces")
    ap.add_argument("--person", required=True, help="Person name to train model for")
    ap.add_argument("--components", type=int, default=50, help="PCA components (reference: 50)")
    args = ap.parse_args()
    raise SystemExit(0 if pipeline.train_person_model(args.person, "faces/lock_version", args.components) else 1)
