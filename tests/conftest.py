import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `-m gpu` under gpurun")


def load_pkg():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("abstracts-search_b200")


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def gpu_pkg():
    """Product package on a GPU box; fails (not skips) if the native library is unusable there."""
    import torch

    assert torch.cuda.is_available(), "GPU test selected but no CUDA device is visible"
    p = load_pkg()
    assert os.path.exists(p.LIB_PATH), "libabsb200.so missing: the CUDA path must be built, there is no fallback"
    p.lib()
    return p


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))
