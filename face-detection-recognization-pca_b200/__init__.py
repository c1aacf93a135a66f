"""B200-native Eigenfaces engine: the recognition + PCA-fit hot path of
saladbkp/face-detection-recognization-PCA behind the reference's own Python entry points.

Importable as `eigenfaces_b200` (see eigenfaces_b200.py at the repository root; the directory name carries the
reference's hyphenated name and is not a valid Python identifier).
"""
from . import _lib, dist, engine, gen1, gen2, manual, pipeline, template  # noqa: F401
from ._lib import METRIC_COSINE_G1, METRIC_COSINE_SK, METRIC_L2, EigenfacesError, launch_count  # noqa: F401
from .engine import Recognizer, fit_gen1, fit_gen2, preprocess_device  # noqa: F401

__all__ = ["Recognizer", "fit_gen1", "fit_gen2", "preprocess_device", "gen1", "gen2", "manual", "engine", "dist", "EigenfacesError",
           "METRIC_COSINE_G1", "METRIC_COSINE_SK", "METRIC_L2", "launch_count"]
