import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def frames():
    return dict(np.load(os.path.join(GOLDEN, "frames.npz")))


@pytest.fixture(scope="session")
def masks():
    return dict(np.load(os.path.join(GOLDEN, "masks.npz")))


@pytest.fixture(scope="session")
def pictures():
    return dict(np.load(os.path.join(GOLDEN, "pictures.npz")))


@pytest.fixture(scope="session")
def fields():
    return dict(np.load(os.path.join(GOLDEN, "fields.npz")))


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


def iou(a, b):
    a = np.asarray(a, bool)
    b = np.asarray(b, bool)
    return (a & b).sum() / max((a | b).sum(), 1)
