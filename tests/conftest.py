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


def rel_err(a, b):
    """max|a-b| / max|b| -- the tolerance form used for every floating-point parity check here
    (north_star: 1e-4 relative; the denominator is the tensor's inf-norm so that entries near zero,
    e.g. distance-field values on the surface, do not blow the ratio up)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elementwise_err(a, b):
    """max over the elements of |a - b| / max(|b|, 1e-3 * max|b|) -- the per-element form SURVEY.md 7 proposes
    (|a - b| <= tol * max(|b|, 1e-3 * ||b||_inf)); reported next to ``rel_err`` by the parity tests that state both."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.abs(b), 1e-3 * max(np.abs(b).max(), 1e-30))
    return float((np.abs(a - b) / den).max())


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load
