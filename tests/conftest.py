import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _n_gpus():
    try:
        from s2s_ismr_unet_b200.runtime import device_count
        return device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _n_gpus() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (d if d > 0 else 1.0))


@pytest.fixture(scope="session")
def stream():
    from s2s_ismr_unet_b200.runtime import Stream
    return Stream()
