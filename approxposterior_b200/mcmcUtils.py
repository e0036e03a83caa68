"""MCMC helpers mirroring the public functions of reference ``approxposterior/mcmcUtils.py``
(``validateMCMCKwargs`` :15-100, ``batchMeansMCSE`` :103-161, ``estimateBurnin`` :164-227).
Same names, arguments, defaults and return values; written for the B200 engine's sampler objects
(anything exposing ``get_autocorr_time(tol=0)`` works)."""
import numpy as np

__all__ = ["validateMCMCKwargs", "batchMeansMCSE", "estimateBurnin"]

_DEFAULT_ITERATIONS = 10000
_WALKERS_PER_DIM = 20


def validateMCMCKwargs(ap, samplerKwargs, mcmcKwargs, verbose=False):
    """Return sanitised (samplerKwargs, mcmcKwargs) for an ``ApproxPosterior`` ``ap``.

    Sampler side: ``ndim`` is always ``ap.ndim`` and ``log_prob_fn`` always ``ap._gpll`` (user values are
    dropped), a user ``backend`` is dropped with a warning (the driver creates its own chain cache), and
    ``nwalkers`` defaults to 20 per dimension.  Run side: ``iterations`` defaults to 10000 and
    ``initial_state`` to ``ap.priorSample(nwalkers)``.  Passing ``samplerKwargs=None`` gives the same
    defaults -- the reference reads a non-existent ``"dim"`` key on that branch (mcmcUtils.py:47) and
    raises KeyError; the documented intent (20 * ndim walkers) is what is implemented here."""
    sk = {} if samplerKwargs is None else samplerKwargs
    had_sampler_kwargs = samplerKwargs is not None
    if had_sampler_kwargs and "backend" in sk:
        print("WARNING: backend in samplerKwargs. approxposterior creates its own!")
        print("with filename = apRun.h5. Disregarding user-supplied backend.")
    for forced in ("ndim", "log_prob_fn", "backend"):
        sk.pop(forced, None)
    sk["ndim"] = ap.ndim
    if "nwalkers" not in sk:
        if had_sampler_kwargs:
            print("WARNING: samplerKwargs provided but nwalkers not in samplerKwargs")
            print("Defaulting to nwalkers = 20 per dimension.")
        sk["nwalkers"] = _WALKERS_PER_DIM * ap.ndim
    sk["log_prob_fn"] = ap._gpll

    mk = {} if mcmcKwargs is None else mcmcKwargs
    had_mcmc_kwargs = mcmcKwargs is not None
    if "iterations" not in mk:
        mk["iterations"] = _DEFAULT_ITERATIONS
        if had_mcmc_kwargs and verbose:
            print("WARNING: mcmcKwargs provided, but iterations not in mcmcKwargs.")
            print("Defaulting to iterations = %d." % _DEFAULT_ITERATIONS)
    if "initial_state" not in mk:
        mk["initial_state"] = ap.priorSample(sk["nwalkers"])
        if had_mcmc_kwargs and verbose:
            print("WARNING: mcmcKwargs provided, but initial_state not in mcmcKwargs.")
            print("Defaulting to nwalkers samples from priorSample.")
    return sk, mk


def batchMeansMCSE(samples, bins=None, fn=None):
    """Monte-Carlo standard error of ``mean(fn(samples))`` by non-overlapping batch means:
    with b = len(samples) // bins,  MCSE^2 = b / (bins - 1) * sum_k (batchmean_k - overall)^2 / len(samples).
    ``bins`` defaults to max(int(sqrt(n)), 2); ``fn`` to the identity."""
    chain = np.asarray(samples)
    values = chain if fn is None else np.asarray(fn(chain))
    nsamp = len(chain)
    if bins is None:
        bins = max(int(np.sqrt(nsamp)), 2)
    if not isinstance(bins, int):
        raise AssertionError("bins must be an integer")
    width = int(nsamp / bins)
    overall = values.mean(axis=0)
    batch = np.stack([values[k * width:(k + 1) * width].sum(axis=0) / width for k in range(bins)])
    spread = ((batch - overall) ** 2).sum(axis=0)
    return np.sqrt(width / (bins - 1) * spread / nsamp)


def estimateBurnin(sampler, estBurnin=True, thinChains=True, verbose=False):
    """(iburn, ithin) from the integrated autocorrelation time tau of a finished chain:
    iburn = int(2 max tau) if ``estBurnin`` else 0;  ithin = max(int(0.5 min tau), 1) if ``thinChains`` else 1.
    Non-finite tau entries are ignored; if none is finite tau = 1 is used."""
    tau = np.atleast_1d(np.asarray(sampler.get_autocorr_time(tol=0), dtype=float))
    finite = np.isfinite(tau)
    if not finite.all():
        tau = tau[finite]
        if tau.size == 0:
            if verbose:
                print("Failed to compute integrated autocorrelation length, tau.")
                print("Setting tau = 1")
            tau = np.ones(1)
    iburn = int(2.0 * tau.max()) if estBurnin else 0
    ithin = max(int(0.5 * tau.min()), 1) if thinChains else 1
    if verbose:
        print("burn-in estimate: %d" % iburn)
        print("thin estimate: %d" % ithin)
    return iburn, ithin
