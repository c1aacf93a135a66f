#!/usr/bin/env python
"""useless/train.py of the reference (dark + light manual-PCA models of one person) on the B200 engine: reads
faces/Dark_version and faces/Light_version, writes models/{person}_{version}_pca_model.pkl, *_model_info.json and the
mean-face / eigenface JPEGs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigenfaces_b200 as ef  # noqa: E402

if __name__ == "__main__":
    sys.exit(0 if ef.gen1.train_dual_models("faces", "models", "Joseph_Lai", 50) else 1)
