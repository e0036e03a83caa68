"""Test problems used as benchmark/parity fixtures (reference ``approxposterior/likelihood.py``),
plus ``BoxPrior``: a uniform box log-prior object the engine can evaluate on the device."""
import numpy as np
from scipy.optimize import rosen

__all__ = ["BoxPrior", "rosenbrockLnlike", "rosenbrockLnprior", "rosenbrockSample", "rosenbrockLnprob",
           "testBOFn", "testBOFnSample", "testBOFnLnPrior", "sphereLnlike", "sphereSample", "sphereLnprior"]


class BoxPrior(object):
    """lnprior(theta) = ``value`` inside the closed box ``bounds``, -inf outside.

    Behaves like the reference's hand-written priors (e.g. likelihood.py:44-63) when called, and
    additionally exposes ``bounds`` / ``value`` so the sampler and predict kernels can apply the same
    gate on the device instead of calling back into Python once per walker."""

    def __init__(self, bounds, value=0.0):
        self.bounds = [(float(a), float(b)) for a, b in bounds]
        self.value = float(value)
        self._lo = np.array([b[0] for b in self.bounds])
        self._hi = np.array([b[1] for b in self.bounds])

    def __call__(self, theta):
        t = np.atleast_1d(np.asarray(theta, dtype=np.float64)).ravel()
        if t.size != self._lo.size or not np.all((t >= self._lo) & (t <= self._hi)):
            return -np.inf
        return self.value

    def sample(self, n=1):
        return np.random.uniform(low=self._lo, high=self._hi, size=(n, self._lo.size)).squeeze()


# -- Rosenbrock (Wang & Li 2017), reference likelihood.py:26-107
def rosenbrockLnlike(theta):
    return -rosen(theta) / 100.0


def rosenbrockLnprior(theta):
    return -np.inf if np.any(np.fabs(theta) > 5) else 0.0


def rosenbrockSample(n=1, dim=2):
    return np.random.uniform(low=-5, high=5, size=(n, dim)).squeeze()


def rosenbrockLnprob(theta):
    lp = rosenbrockLnprior(theta)
    return -np.inf if not np.isfinite(lp) else lp + rosenbrockLnlike(theta)


# -- 1-D Bayesian-optimisation test function, reference likelihood.py:116-170
def testBOFn(theta):
    theta = np.asarray(theta)
    return -np.sin(3.0 * theta) - theta ** 2 + 0.7 * theta


def testBOFnSample(n=1):
    return np.random.uniform(low=-1, high=2, size=(n, 1)).squeeze()


def testBOFnLnPrior(theta):
    return -np.inf if np.any(np.asarray(theta) < -1) or np.any(np.asarray(theta) > 2) else 0.0


# -- 2-D sphere, reference likelihood.py:179-240
def sphereLnlike(theta):
    return -np.sum(np.asarray(theta) ** 2)


def sphereSample(n=1):
    return np.random.uniform(low=-2, high=2, size=(n, 2)).squeeze()


def sphereLnprior(theta):
    return -np.inf if np.any(np.fabs(theta) > 2) else 0.0
