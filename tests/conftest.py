import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def rosenbrock_training(m0, seed=57, dim=2):
    """Fixture of the reference KATs (tests/test_GPUtil.py:30-41): seed, U(-5,5) sample, y = -rosen/100."""
    from scipy.optimize import rosen
    np.random.seed(seed)
    theta = np.random.uniform(low=-5, high=5, size=(m0, dim))
    y = np.array([-rosen(t) / 100.0 for t in theta])
    return theta, y


def synthetic_gp_problem(N, d, seed=0, logM=None, amp=None):
    """cfg5-style synthetic problem: X ~ U(-5,5)^d, y ~ N(0,1), log M = log d."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, size=(N, d))
    y = rng.standard_normal(N)
    logM = np.full(d, np.log(d)) if logM is None else np.asarray(logM, dtype=float)
    return X, y, logM, amp


@pytest.fixture
def kat_small():
    return rosenbrock_training(20)
