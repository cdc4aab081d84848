import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_numpy.npz"), allow_pickle=False)


class FakeModel:
    """Same stand-in as tests/golden/make_golden.py (kept in sync by hand)."""

    def predict_single(self, image, mean, std):
        x = (image - mean) / (std + 1e-10)
        x = x.astype(np.float32)
        y = 0.6 * x + 0.3 * np.roll(x, 1, axis=0) - 0.2 * np.roll(x, 2, axis=1)
        return (1.0 / (1.0 + np.exp(-y))).astype(np.float32)


@pytest.fixture(scope="session")
def fake_model():
    return FakeModel()
