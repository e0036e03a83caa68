"""george / emcee stand-ins backed by the CPU oracle (TEST INFRASTRUCTURE ONLY, see ``oracle/__init__.py``).

``install()`` puts two synthetic modules into ``sys.modules`` so that the UNMODIFIED reference package
(dflemin3/approxposterior v0.4, imported from ``/root/reference`` or from the ``baseline/_ref`` install) runs on
the NumPy/SciPy restatement of george 0.3.x and emcee 3.0.x.  Two uses:

* ``tests/test_reference_dropin.py`` (CPU half): the reference's own test modules execute against the oracle --
  a second, independent pin of the oracle (the reference's driver code instead of this repo's restated drivers);
* ``bench.py --impl reference``: the reference's own ``ApproxPosterior.run`` timed on the host cores
  (the "BAPE iteration time" half of the metric).

Surface provided = what the reference touches: ``george.GP(kernel=, fit_mean=, mean=, white_noise=,
fit_white_noise=)``, ``george.kernels.ExpSquaredKernel(metric=, ndim=)``, ``float * kernel``;
``emcee.EnsembleSampler(nwalkers, ndim, log_prob_fn, args=, kwargs=, backend=, blobs_dtype=)`` with
``run_mcmc`` / ``sample`` / ``get_chain`` / ``get_autocorr_time``, ``emcee.backends.HDFBackend``,
``emcee.__version__``.

``scipy_x0_compat(module)`` is not part of the shim: the reference passes a 2-D ``x0`` to
``scipy.optimize.minimize`` (utility.py:336,364), which the SciPy of its day flattened and SciPy >= 1.11 rejects;
the helper wraps the ``minimize`` name inside the reference's ``utility`` module so that x0 is flattened again.
"""
import sys
import types

import numpy as np

from .gp_oracle import GPOracle
from .sampler_oracle import stretch_move_oracle

_SAVED = {}


# ---------------------------------------------------------------------------------------------- george
class _ExpSquared(object):
    """``george.kernels.ExpSquaredKernel(metric, ndim)`` (+ optional constant factor): a parameter carrier."""
    __array_ufunc__ = None            # np.float64 * kernel -> kernel.__rmul__

    def __init__(self, metric, ndim=1, amp=None):
        self.ndim = int(ndim)
        m = np.atleast_1d(np.asarray(metric, dtype=np.float64))
        if m.size == 1 and self.ndim > 1:
            m = np.full(self.ndim, float(m[0]))
        self.metric = m
        self.amp = amp

    def __rmul__(self, a):
        return _ExpSquared(self.metric, self.ndim, amp=float(a))

    __mul__ = __rmul__

    def __add__(self, other):
        raise NotImplementedError("kernel sums (LinearKernel, gpUtils.py:169-173) are outside the restated path")


def _linear_kernel(*args, **kwargs):
    raise NotImplementedError("LinearKernel (defaultGP(order=...)) is outside the restated path")


class GP(GPOracle):
    """george.GP constructor signature on the oracle GP."""

    def __init__(self, kernel=None, fit_mean=False, mean=0.0, white_noise=None, fit_white_noise=False, **kwargs):
        if not isinstance(kernel, _ExpSquared) or fit_white_noise or not fit_mean:
            raise NotImplementedError("restated path: ExpSquared (x constant), fitted constant mean, frozen white noise")
        wn = np.log(1.25e-12 ** 2) if white_noise is None else float(white_noise)
        GPOracle.__init__(self, kernel.ndim, kernel.metric, amp=kernel.amp, mean=float(np.asarray(mean)), white_noise=wn)

    @property
    def kernel(self):                 # round-trips through approx.py:712-717
        return _ExpSquared(np.exp(self.log_M), self.ndim, amp=(self.amplitude if self.fit_amp else None))


# ---------------------------------------------------------------------------------------------- emcee
def _next_pow_two(n):
    i = 1
    while i < n:
        i = i << 1
    return i


def function_1d(x):
    """emcee.autocorr.function_1d: normalised autocorrelation of one series by zero-padded FFT."""
    x = np.atleast_1d(x)
    n = _next_pow_two(len(x))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[: len(x)].real
    acf /= acf[0]
    return acf


def auto_window(taus, c):
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return np.argmin(m)
    return len(taus) - 1


class AutocorrError(Exception):
    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super(AutocorrError, self).__init__(*args, **kwargs)


def integrated_time(x, c=5, tol=50, quiet=False):
    """emcee.autocorr.integrated_time (3.0.x): per dimension, the walker-averaged normalised autocorrelation,
    tau(M) = 2 cumsum - 1, Sokal's window (smallest M with M >= c tau(M))."""
    x = np.atleast_1d(x)
    if len(x.shape) == 1:
        x = x[:, np.newaxis, np.newaxis]
    if len(x.shape) == 2:
        x = x[:, :, np.newaxis]
    if len(x.shape) != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    windows = np.empty(n_d, dtype=int)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += function_1d(x[:, k, d])
        f /= n_w
        taus = 2.0 * np.cumsum(f) - 1.0
        windows[d] = auto_window(taus, c)
        tau_est[d] = taus[windows[d]]
    flag = tol * tau_est > n_t
    if np.any(flag) and not quiet:
        raise AutocorrError(tau_est, "The chain is shorter than {0} times the integrated autocorrelation time".format(tol))
    return tau_est


class HDFBackend(object):
    def __init__(self, filename, name="mcmc", read_only=False, **kwargs):
        self.filename = str(filename)

    def reset(self, nwalkers, ndim):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)


class _State(object):
    def __init__(self, coords, log_prob, blobs):
        self.coords, self.log_prob, self.blobs = coords, log_prob, blobs


class EnsembleSampler(object):
    """emcee.EnsembleSampler (3.0.x) on ``stretch_move_oracle``: one Python ``log_prob_fn`` call per walker."""

    def __init__(self, nwalkers, ndim, log_prob_fn, pool=None, moves=None, args=None, kwargs=None, backend=None,
                 vectorize=False, blobs_dtype=None, a=None, **unused):
        if moves is not None or pool is not None or vectorize:
            raise NotImplementedError("restated path: default StretchMove, no pool, scalar log_prob_fn")
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("emcee requires an even number of walkers, at least twice the dimension")
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self.a = 2.0 if a is None else float(a)
        self._fn, self._args, self._kwargs = log_prob_fn, (() if args is None else tuple(args)), ({} if kwargs is None else dict(kwargs))
        self.backend, self.blobs_dtype = backend, blobs_dtype
        self._random = np.random.RandomState()
        self._random.set_state(np.random.get_state())      # emcee copies the global state at construction
        self._chain = self._logp = self._blobs = None
        self.naccepted = np.zeros(self.nwalkers, dtype=np.int64)
        self.iteration = 0

    def _batch(self, q):
        lp = np.empty(len(q))
        blob = np.full(len(q), np.nan)
        for i, t in enumerate(q):
            out = self._fn(t, *self._args, **self._kwargs)
            if isinstance(out, (tuple, list)):
                lp[i] = float(np.asarray(out[0]).ravel()[0])
                blob[i] = float(np.asarray(out[1]).ravel()[0])
            else:
                lp[i] = float(np.asarray(out).ravel()[0])
        if np.any(np.isnan(lp)):
            raise ValueError("Probability function returned NaN")
        return lp, blob

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        p0 = np.asarray(getattr(initial_state, "coords", initial_state), dtype=np.float64).reshape(self.nwalkers, self.ndim)
        out = stretch_move_oracle(self._batch, p0, int(nsteps), a=self.a, rng=self._random)
        if self._chain is None:
            self._chain, self._logp, self._blobs = out["chain"], out["log_prob"], out["blobs"]
        else:
            self._chain = np.concatenate([self._chain, out["chain"]])
            self._logp = np.concatenate([self._logp, out["log_prob"]])
            self._blobs = np.concatenate([self._blobs, out["blobs"]])
        self.naccepted = self.naccepted + out["naccepted"]
        self.iteration += int(nsteps)
        return _State(self._chain[-1], self._logp[-1], self._blobs[-1])

    def sample(self, initial_state, iterations=1, **kwargs):
        state = self.run_mcmc(initial_state, iterations)
        for _ in range(int(iterations)):
            yield state

    def _get(self, arr, discard=0, flat=False, thin=1):
        if arr is None:
            raise AttributeError("you must run the sampler before accessing the results")
        v = arr[discard + thin - 1::thin]
        if flat:
            v = v.reshape((-1,) + v.shape[2:])
        return v

    def get_chain(self, **kw):
        return self._get(self._chain, **kw)

    def get_log_prob(self, **kw):
        return self._get(self._logp, **kw)

    def get_blobs(self, **kw):
        return self._get(self._blobs, **kw)

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(self.iteration)

    def get_autocorr_time(self, discard=0, thin=1, **kwargs):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), **kwargs)


# ---------------------------------------------------------------------------------------------- install
def install():
    george = types.ModuleType("george")
    george.__version__ = "0.3.1+oracle"
    george.GP = GP
    gk = types.ModuleType("george.kernels")
    gk.ExpSquaredKernel = _ExpSquared
    gk.LinearKernel = _linear_kernel
    george.kernels = gk
    emcee = types.ModuleType("emcee")
    emcee.__version__ = "3.0.2"
    emcee.EnsembleSampler = EnsembleSampler
    emcee.State = _State
    eb = types.ModuleType("emcee.backends")
    eb.HDFBackend = HDFBackend
    emcee.backends = eb
    ea = types.ModuleType("emcee.autocorr")
    ea.integrated_time, ea.AutocorrError, ea.function_1d = integrated_time, AutocorrError, function_1d
    emcee.autocorr = ea
    mods = {"george": george, "george.kernels": gk, "emcee": emcee, "emcee.backends": eb, "emcee.autocorr": ea}
    for name, mod in mods.items():
        if name not in _SAVED:
            _SAVED[name] = sys.modules.get(name)
        sys.modules[name] = mod
    return mods


def uninstall():
    for name, prev in list(_SAVED.items()):
        if prev is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = prev
        del _SAVED[name]


def scipy_x0_compat(utility_module):
    """Flatten x0 before ``scipy.optimize.minimize`` inside the reference's ``utility`` module, as the SciPy
    contemporary with the reference did itself (reference utility.py:336,364 pass shape (1, ndim))."""
    import scipy.optimize as so

    def minimize(fun, x0, *a, **kw):
        return so.minimize(fun, np.ravel(x0), *a, **kw)
    utility_module.minimize = minimize
