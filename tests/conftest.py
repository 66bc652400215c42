"""pytest configuration: markers and shared fixtures."""

import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

from ms_test_helpers import GOLDEN  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def kernels_golden():
    return dict(np.load(os.path.join(GOLDEN, "kernels.npz")))


@pytest.fixture(scope="session")
def minimizer_golden():
    return dict(np.load(os.path.join(GOLDEN, "minimizer.npz")))
