"""Restatement of the acquisition utilities (TEST INFRASTRUCTURE ONLY).

Formula-for-formula after reference ``approxposterior/utility.py``:
``logsubexp`` :69-89 (naive log(1-exp), kept on purpose), ``AGPUtility`` :99-142,
``BAPEUtility`` :145-189, ``JonesUtility`` :192-250.  Vectorised over queries;
``prior_ok`` is the boolean "lnprior finite" gate of :126/:173/:219.
"""
import numpy as np
from scipy.special import ndtr


def logsubexp(x1, x2):
    x1 = np.asarray(x1, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        out = x1 + np.log(1.0 - np.exp(x2 - x1))
    return np.where(x1 <= x2, -np.inf, out)


def agp_utility(mu, var, prior_ok=True):
    with np.errstate(divide="ignore", invalid="ignore"):
        u = -(mu + 0.5 * np.log(2.0 * np.pi * np.e * var))
    return np.where(prior_ok, u, np.inf)


def bape_utility(mu, var, prior_ok=True):
    u = -((2.0 * mu + var) + logsubexp(var, 0.0))
    return np.where(prior_ok, u, np.inf)


def jones_utility(mu, var, ybest, zeta=0.01, prior_ok=True):
    mu = np.asarray(mu, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        std = np.sqrt(var)
        z = (mu - ybest - zeta) / std
        cdf = ndtr(z)
        pdf = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)
        u = -((mu - ybest - zeta) * cdf + std * pdf)
    u = np.where(std > 0, u, 0.0)           # utility.py:235-238 (NaN std -> 0.0 too)
    return np.where(prior_ok, u, np.inf)


UTILITY_BY_NAME = {"agp": agp_utility, "bape": bape_utility, "jones": jones_utility}
