import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "real-robot-nerf-actor_b200"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg(sub=None):
    return importlib.import_module(PKG_NAME + ("." + sub if sub else ""))


@pytest.fixture(scope="session")
def pkg():
    return load_pkg()


@pytest.fixture(scope="session")
def syn():
    return load_pkg("synthetic")


def golden(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name + ".npz"))
