"""Drop-in shims: ``import george`` / ``import emcee`` resolve to the B200 engine.

``install()`` registers two synthetic modules in ``sys.modules`` so that the UNMODIFIED reference
package (dflemin3/approxposterior v0.4) imports and runs on ``libapgp.so``:

  george          ``GP`` (-> ``approxposterior_b200.GP``) and ``kernels`` (``ExpSquaredKernel``, constant
                  scaling, ``LinearKernel`` raising) -- everything the reference touches at
                  gpUtils.py:160-178, approx.py:712-717, utility.py:130-131,177-178,223-224.
  emcee           ``EnsembleSampler`` with emcee 3.0's constructor signature (scalar ``log_prob_fn(theta,
                  *args, **kwargs)`` returning ``lp`` or ``(lp, blob)``), ``run_mcmc`` / ``sample`` /
                  ``get_chain`` / ``get_log_prob`` / ``get_blobs`` / ``get_autocorr_time``,
                  ``backends.HDFBackend(name).reset(nwalkers, ndim)`` (the chain is written to ``name`` in emcee's
                  HDF5 layout by ``hdf5min``: h5py is not available offline), ``autocorr.integrated_time`` and ``__version__ = "3.0.2"`` --
                  approx.py:832-847, mcmcUtils.py:198, tests/test_Import.py.

The sampler follows emcee 3.0.x's NumPy RNG flow draw for draw (``sampler.EnsembleSampler``'s "host-rng"
engine), so seeded reference runs keep their random stream.  When ``log_prob_fn`` is the reference's own
``ApproxPosterior._gpll`` (approx.py:148-189) bound to an engine GP, each half-step's proposals are evaluated
with ONE batched mean-only predict on the device instead of one Python call per walker: the guards of
``_gpll`` (all-non-finite input, non-finite prior, non-finite mean -> ``(-inf, nan)``) are applied row by row
around that call.  ``install(box_prior_sampler=True)`` additionally lets such a sampler run the whole chain inside
the device kernel when the owner exposes ``bounds`` and its ``_lnprior`` is a box over them (the caller's
promise; Philox draws instead of NumPy's).

``accelerate()`` additionally points the reference's two multistart drivers (``gpUtils.optimizeGP``,
``utility.minimizeObjective``: module attributes looked up by approx.py) at the engine's batched drivers of the same
names, still without touching a reference source file.

``tests/test_reference_dropin.py`` runs the reference's own test modules through these shims.
"""
import sys
import types

import numpy as np

from . import kernels as _kernels
from .gp import GP as _GP
from . import sampler as _sampler

__all__ = ["install", "accelerate", "uninstall", "installed"]

_SAVED = {}
_OPTIONS = {"box_prior_sampler": False}


class HDFBackend(object):
    """``emcee.backends.HDFBackend(filename)``: the reference only constructs it, calls ``reset`` and hands it
    to the sampler (approx.py:832-839).  Chains are written to ``filename`` in emcee's HDF5 layout by
    ``approxposterior_b200.hdf5min`` (group ``mcmc``: chain, log_prob, blobs, accepted + attributes) and to an
    ``.npz`` twin with the same array names."""

    def __init__(self, filename, name="mcmc", read_only=False, **kwargs):
        self.filename = str(filename)
        self.name = name
        self.nwalkers = self.ndim = None

    def reset(self, nwalkers, ndim):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)

    def __str__(self):
        f = self.filename
        return f[:-3] if f.endswith(".h5") else f


def _is_reference_gpll(fn):
    """True for a bound ``_gpll`` whose owner carries an engine GP, its training targets and a prior."""
    owner = getattr(fn, "__self__", None)
    return (owner is not None and getattr(fn, "__name__", "") == "_gpll" and isinstance(getattr(owner, "gp", None), _GP)
            and hasattr(owner, "y") and callable(getattr(owner, "_lnprior", None)))


class EnsembleSampler(_sampler.EnsembleSampler):
    """emcee.EnsembleSampler signature on the engine's sampler."""

    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, moves=None, args=None, kwargs=None, backend=None,
                 vectorize=False, blobs_dtype=None, a=None, **unused):
        if moves is not None:
            raise NotImplementedError("only emcee's default StretchMove(a=2) is provided")
        if pool is not None:
            raise NotImplementedError("pool= is not supported: proposals are evaluated in batches on the device")
        fargs = () if args is None else tuple(args)
        fkw = {} if kwargs is None else dict(kwargs)
        engine, extra = "host-rng", {}
        if _is_reference_gpll(log_prob_fn):
            owner = log_prob_fn.__self__

            def batch(q):                                    # approx.py:148-189, row by row around ONE predict
                q = np.asarray(q, dtype=np.float64).reshape(-1, int(ndim))
                lp = np.full(q.shape[0], -np.inf)
                blob = np.full(q.shape[0], np.nan)
                pri = np.array([owner._lnprior(t) if np.any(np.isfinite(t)) else -np.inf for t in q], dtype=np.float64)
                ok = np.isfinite(pri) & np.all(np.isfinite(q), axis=1)
                if np.any(ok):
                    try:
                        mu = np.asarray(owner.gp.predict(owner.y, q[ok], return_cov=False, return_var=False))
                    except ValueError:
                        return lp, blob
                    fin = np.isfinite(mu)
                    idx = np.nonzero(ok)[0][fin]
                    lp[idx] = mu[fin]
                    blob[idx] = pri[idx]
                return lp, blob
            if _OPTIONS["box_prior_sampler"] and getattr(owner, "bounds", None) is not None:
                engine = "device"
                inside = np.array([0.5 * (float(b[0]) + float(b[1])) for b in owner.bounds])
                extra = dict(gp=owner.gp, y=owner.y, bounds=owner.bounds, lnprior_const=float(owner._lnprior(inside)))
        elif vectorize:
            def batch(q):
                out = log_prob_fn(q, *fargs, **fkw)
                if isinstance(out, tuple):
                    return np.asarray(out[0], dtype=np.float64), np.asarray(out[1], dtype=np.float64)
                return np.asarray(out, dtype=np.float64), np.full(len(q), np.nan)
        else:
            def batch(q):                                    # emcee's default: one Python call per walker
                lp = np.empty(len(q))
                blob = np.full(len(q), np.nan)
                for i, t in enumerate(q):
                    out = log_prob_fn(t, *fargs, **fkw)
                    if isinstance(out, (tuple, list)) or (isinstance(out, np.ndarray) and out.ndim > 0 and out.size > 1):
                        lp[i] = float(np.asarray(out[0]).ravel()[0])
                        try:
                            blob[i] = float(np.asarray(out[1]).ravel()[0])
                        except (TypeError, ValueError):
                            pass
                    else:
                        lp[i] = float(np.asarray(out).ravel()[0])
                if np.any(np.isnan(lp)):
                    raise ValueError("Probability function returned NaN")
                return lp, blob
        super(EnsembleSampler, self).__init__(nwalkers, ndim, log_prob_fn=batch, backend=backend, blobs_dtype=blobs_dtype,
                                              a=2.0 if a is None else a, engine=engine, **extra)


def _build_modules():
    george = types.ModuleType("george")
    george.__doc__ = "george surface used by approxposterior, served by approxposterior_b200 (libapgp.so)"
    george.__version__ = "0.3.1+b200"
    george.GP = _GP
    gk = types.ModuleType("george.kernels")
    for name in _kernels.__all__:
        setattr(gk, name, getattr(_kernels, name))
    george.kernels = gk

    emcee = types.ModuleType("emcee")
    emcee.__doc__ = "emcee surface used by approxposterior, served by approxposterior_b200 (libapgp.so)"
    emcee.__version__ = "3.0.2"
    emcee.EnsembleSampler = EnsembleSampler
    emcee.State = _sampler._State
    eb = types.ModuleType("emcee.backends")
    eb.HDFBackend = HDFBackend
    emcee.backends = eb
    ea = types.ModuleType("emcee.autocorr")
    ea.integrated_time = _sampler.integrated_time
    ea.AutocorrError = _sampler.AutocorrError
    emcee.autocorr = ea
    return {"george": george, "george.kernels": gk, "emcee": emcee, "emcee.backends": eb, "emcee.autocorr": ea}


def install(box_prior_sampler=False):
    """Register the george / emcee shims (idempotent).  Returns the dict of installed modules."""
    _OPTIONS["box_prior_sampler"] = bool(box_prior_sampler)
    mods = _build_modules()
    for name, mod in mods.items():
        if name not in _SAVED:
            _SAVED[name] = sys.modules.get(name)
        sys.modules[name] = mod
    return mods


_PATCHED = {}


def accelerate(box_prior=False):
    """Integration level between 0 and 2: keep the reference's source files untouched, but let its two multistart
    drivers -- ``approxposterior.gpUtils.optimizeGP`` (gpUtils.py:184-257) and ``approxposterior.utility.minimizeObjective``
    (utility.py:253-372), both looked up as module attributes by approx.py:222,664,918 -- resolve to the engine's batched
    drivers of the same names and signatures.  Every restart of a hyper-parameter fit then runs inside ONE device launch
    and the utility restarts advance in lock step (or on the device with ``box_prior=True``, the caller's promise that
    ``lnprior`` is the box over ``bounds``), instead of one GPU call per Python-level objective evaluation.
    Call after ``install()``; imports ``approxposterior`` (it must be importable)."""
    import importlib
    from . import gpUtils as _gu, utility as _ut
    rgu = importlib.import_module("approxposterior.gpUtils")
    rut = importlib.import_module("approxposterior.utility")
    for mod, name, new in ((rgu, "optimizeGP", _gu.optimizeGP), (rut, "minimizeObjective", _ut.minimizeObjective)):
        if (mod.__name__, name) not in _PATCHED:
            _PATCHED[(mod.__name__, name)] = (mod, getattr(mod, name))
        setattr(mod, name, new)
    _PATCHED.setdefault("box", _ut.ASSUME_BOX_PRIOR)
    _ut.ASSUME_BOX_PRIOR = bool(box_prior)


def uninstall():
    """Restore whatever ``sys.modules`` held before ``install()`` (and undo ``accelerate()``)."""
    from . import utility as _ut
    if "box" in _PATCHED:
        _ut.ASSUME_BOX_PRIOR = _PATCHED.pop("box")
    for key, (mod, orig) in list(_PATCHED.items()):
        setattr(mod, key[1], orig)
        del _PATCHED[key]
    for name, prev in list(_SAVED.items()):
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev
        del _SAVED[name]


def installed():
    return bool(_SAVED)
