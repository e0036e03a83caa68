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


def extended_truth(X, y, logM, Xq, amp=None, mean=0.0, wn=-12.0, nvar=0):
    """Higher-precision arbiter for fp64 disagreements: the GP predictive mean (and, for the first ``nvar`` queries,
    variance) of the george model evaluated in x87 extended precision (64-bit mantissa) with iteratively refined
    solves -- error ~ cond(K) * 1e-19, far below either fp64 implementation's cond(K) * 1e-16.  Used where cond(K)
    makes a 1e-9 engine-vs-oracle comparison meaningless: the engine must then be as close to this as the oracle."""
    from scipy.linalg import cho_factor, cho_solve
    ld = np.longdouble
    if np.finfo(ld).eps > 1e-18:
        pytest.skip("no extended-precision long double on this platform")
    X = np.asarray(X, dtype=np.float64); Xq = np.asarray(Xq, dtype=np.float64)
    N, d = X.shape
    w = np.exp(-np.asarray(logM, dtype=ld))
    A = ld(1.0) if amp is None else ld(X.shape[1]) * np.exp(np.log(ld(amp) / ld(d)))   # george: ndim * exp(log(a/ndim))

    def kern(a, b):
        s = np.zeros((a.shape[0], b.shape[0]), dtype=ld)
        for i in range(d):
            df = a[:, i].astype(ld)[:, None] - b[:, i].astype(ld)[None, :]
            s += df * df * w[i]
        return A * np.exp(ld(-0.5) * s)

    K = kern(X, X)
    K[np.diag_indices(N)] += np.exp(ld(wn)) + ld(1.25e-12) ** 2
    cf = cho_factor(K.astype(np.float64), lower=True)

    def solve(B):                       # iterative refinement: fp64 factor, extended-precision residuals
        Xs = cho_solve(cf, B.astype(np.float64)).astype(ld)
        for _ in range(6):
            R = B - K @ Xs
            Xs = Xs + cho_solve(cf, R.astype(np.float64)).astype(ld)
        return Xs

    r = np.asarray(y, dtype=ld) - ld(mean)
    alpha = solve(r)
    Kq = kern(Xq, X)
    mu = (Kq @ alpha + ld(mean)).astype(np.float64)
    var = None
    if nvar:
        S = solve(Kq[:nvar].T.copy())
        var = (A - np.sum(Kq[:nvar].T * S, axis=0)).astype(np.float64)
    return mu, var


def utility_error_bound(kind, mu, var, A, scale, ybest=0.0, zeta=0.01):
    """Largest |u_engine - u_oracle| that north_star's 1e-9 parity on (mu, var) allows for the utilities of reference
    utility.py:136,183,229-244, by first-order propagation:  |du| <= |du/dmu| dmu + |du/dvar| dvar  with
    dmu = 1e-9 (|mu| + scale) and dvar = 1e-9 (A + |var|)  (the variance is a difference of O(A) terms).
      AGP    u = -(mu + 1/2 log(2 pi e var))                 du/dvar = -1/(2 var)
      BAPE   u = -(2 mu + 2 var + log(1 - exp(-var)))        du/dvar = -(2 + 1/expm1(var))
      Jones  u = -((mu - yb - zeta) Phi(z) + s phi(z))       du/dmu = -Phi(z), du/dvar = -phi(z)/(2 s)
    Returns (bound, resolved): `resolved` is False where var <= 4 dvar (the logarithm of an unresolved variance is
    not comparable; those entries are checked for sign/finite-ness only)."""
    mu, var = np.asarray(mu, dtype=np.float64), np.asarray(var, dtype=np.float64)
    dmu = 1e-9 * (np.abs(mu) + scale)
    dvar = 1e-9 * (A + np.abs(var))
    resolved = var > 4.0 * dvar
    v = np.where(resolved, var - dvar, 1.0)
    if kind == "agp":
        b = dmu + 0.5 * dvar / v
    elif kind == "bape":
        b = 2.0 * dmu + dvar * (2.0 + 1.0 / np.expm1(v)) + 4.0 * np.finfo(float).eps / (-np.expm1(-v))
    elif kind == "jones":
        s = np.sqrt(v)
        b = dmu + dvar * 0.2 / s
    else:
        raise ValueError(kind)
    return b, resolved
