"""MCMC helpers: mirror of reference ``approxposterior/mcmcUtils.py``."""
import numpy as np

__all__ = ["validateMCMCKwargs", "batchMeansMCSE", "estimateBurnin"]


def validateMCMCKwargs(ap, samplerKwargs, mcmcKwargs, verbose=False):
    """Sanitise the sampler / run kwargs (reference mcmcUtils.py:15-100): ndim and log_prob_fn are
    forced, a user backend is dropped, defaults are nwalkers = 20*ndim, iterations = 10000 and
    initial_state = ap.priorSample(nwalkers).  (The reference's ``samplerKwargs=None`` branch reads
    a non-existent "dim" key, mcmcUtils.py:47; the intended 20*ndim default is used here.)"""
    if samplerKwargs is None:
        samplerKwargs = dict()
        samplerKwargs["ndim"] = ap.ndim
        samplerKwargs["nwalkers"] = 20 * samplerKwargs["ndim"]
        samplerKwargs["log_prob_fn"] = ap._gpll
    else:
        samplerKwargs.pop("ndim", None)
        samplerKwargs["ndim"] = ap.ndim
        if "nwalkers" not in samplerKwargs:
            print("WARNING: samplerKwargs provided but nwalkers not in samplerKwargs")
            print("Defaulting to nwalkers = 20 per dimension.")
            samplerKwargs["nwalkers"] = 20 * samplerKwargs["ndim"]
        if "backend" in samplerKwargs:
            print("WARNING: backend in samplerKwargs. approxposterior creates its own!")
            print("with filename = apRun.h5. Disregarding user-supplied backend.")
        samplerKwargs.pop("log_prob_fn", None)
        samplerKwargs.pop("backend", None)
        samplerKwargs["log_prob_fn"] = ap._gpll

    if mcmcKwargs is None:
        mcmcKwargs = dict()
        mcmcKwargs["iterations"] = 10000
        mcmcKwargs["initial_state"] = ap.priorSample(samplerKwargs["nwalkers"])
    else:
        if "iterations" not in mcmcKwargs:
            mcmcKwargs["iterations"] = 10000
            if verbose:
                print("WARNING: mcmcKwargs provided, but iterations not in mcmcKwargs.")
                print("Defaulting to iterations = 10000.")
        if "initial_state" not in mcmcKwargs:
            mcmcKwargs["initial_state"] = ap.priorSample(samplerKwargs["nwalkers"])
            if verbose:
                print("WARNING: mcmcKwargs provided, but initial_state not in mcmcKwargs.")
                print("Defaulting to nwalkers samples from priorSample.")
    return samplerKwargs, mcmcKwargs


def batchMeansMCSE(samples, bins=None, fn=None):
    """Non-overlapping batch-means Monte-Carlo standard error (reference mcmcUtils.py:103-161)."""
    samples = np.asarray(samples)
    vals = samples if fn is None else np.asarray(fn(samples))
    n = len(samples)
    if bins is None:
        bins = max(int(np.sqrt(n)), 2)
    assert isinstance(bins, int), "bins must be an integer"
    size = int(n / bins)
    total = np.mean(vals, axis=0)                       # statistic over the whole chain
    means = np.array([np.sum(vals[i * size:(i + 1) * size], axis=0) / size for i in range(bins)])
    mcse = size / (bins - 1) * np.sum((means - total) ** 2, axis=0)
    return np.sqrt(mcse / n)


def estimateBurnin(sampler, estBurnin=True, thinChains=True, verbose=False):
    """Burn-in = int(2 max tau), thin = max(int(0.5 min tau), 1) from the integrated autocorrelation
    time with tol=0 (reference mcmcUtils.py:164-227)."""
    tau = sampler.get_autocorr_time(tol=0)
    if np.any(~np.isfinite(tau)):
        tau = tau[np.isfinite(np.array(tau))]
        if len(tau) < 1:
            if verbose:
                print("Failed to compute integrated autocorrelation length, tau.")
                print("Setting tau = 1")
            tau = 1
    iburn = int(2.0 * np.max(tau)) if estBurnin else 0
    ithin = np.max((int(0.5 * np.min(tau)), 1)) if thinChains else 1
    if verbose:
        print("burn-in estimate: %d" % iburn)
        print("thin estimate: %d" % ithin)
    return iburn, ithin
