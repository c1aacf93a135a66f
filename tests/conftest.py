import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = _load(name)
        return cache[name]
    return get


@pytest.fixture(scope="session")
def light_model(golden):
    """Gen-1 light model rebuilt by the oracle from the golden training matrix (== shipped pkl to 1e-14)."""
    from oracle import gen1
    g = golden("gen1_light.npz")
    ef, mean, proj, ev = gen1.manual_pca(g["X_u8"].astype(np.float64), 50)
    return gen1.model_dict(ef, mean, proj, ev, [str(s) for s in g["filenames"]], "Joseph_Lai", "light", "golden")


@pytest.fixture(scope="session")
def dark_model(golden):
    from oracle import gen1
    g = golden("gen1_dark.npz")
    ef, mean, proj, ev = gen1.manual_pca(g["X_u8"].astype(np.float64), 50)
    return gen1.model_dict(ef, mean, proj, ev, [str(s) for s in g["filenames"]], "Joseph_Lai_dark", "dark", "golden")
