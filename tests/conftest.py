import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available() -> bool:
    try:
        import pseudo_3d_interpolation_b200 as m
        return m._lib.load().p3d_device_count() > 0
    except Exception:          # noqa: BLE001 - library missing or not loadable: no GPU tests
        return False


def pytest_collection_modifyitems(config, items):
    """Skip `gpu`-marked tests on hosts without a CUDA device or without the built library (plain `pytest tests` here)."""
    if not any("gpu" in it.keywords for it in items) or _gpu_available():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libp3d_b200.so")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_pocs.npz"))
