#!/usr/bin/env python
"""Drop-in for the reference's train-v5.py (no arguments; walks faces/lock_version/*): one Eigenfaces model per person
with as many components as crops, fitted on the B200 engine (StandardScaler + exact full-SVD-equivalent PCA), written
as face_model.pkl + multi_person_* JPEG/JSON that the reference's scan-template-v4.py loads unchanged."""
import _bootstrap  # noqa: F401
from eigenfaces_b200 import pipeline

if __name__ == "__main__":
    ok, _ = pipeline.train_all_persons("faces/lock_version")
    if ok:
        print("\nModels saved in individual person directories under faces/lock_version")
        print("Each model can be used with scan-template scripts for person-specific recognition")
