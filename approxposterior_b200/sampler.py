"""``EnsembleSampler``: the slice of emcee's surface that approxposterior drives
(reference approx.py:839-847, mcmcUtils.py:198, approx.py:479), backed by the B200 engine.

Two engines:
  "device"   the whole chain runs inside one CUDA kernel (GP.run_ensembles): stretch move, red/blue
             split, box prior, mean-only surrogate predict, accept/reject and chain storage all on the
             GPU with Philox draws.  Needs a box prior (``bounds``) and the surrogate GP.  ``nens``
             independent ensembles of ``nwalkers`` walkers run side by side.
  "host-rng" emcee 3.0.x's draw order is replayed on the host with NumPy's legacy RandomState
             (choice, shuffle, rand, randint, rand per walker) so a seeded run follows the reference's
             RNG flow; each half-step's proposals are evaluated in ONE batched call to
             ``log_prob_fn`` (vectorised: takes (Ns, ndim), returns (lp[Ns], blob[Ns])).
"""
import numpy as np

__all__ = ["EnsembleSampler", "integrated_time", "AutocorrError"]


class AutocorrError(Exception):
    def __init__(self, tau, *args, **kwargs):
        self.tau = tau
        super(AutocorrError, self).__init__(*args, **kwargs)


def _mean_norm_acf(x):
    """Mean over walkers of each walker's normalised autocorrelation function, x[n_t, n_w] -> [n_t]
    (what emcee's ``integrated_time`` averages).  Linear autocorrelation by zero-padded FFT; because the inverse
    transform is linear, the per-walker normalisation acf_w[0] = sum_t x_w(t)^2 is applied to the power spectra and
    ONE inverse transform serves all walkers.  Large chains (the device sampler produces tens of thousands of
    walkers) are transformed with torch.fft on the GPU when torch is already loaded."""
    import sys
    n_t = x.shape[0]
    if x.size >= (1 << 22) and "torch" in sys.modules:
        try:
            import torch
            if torch.cuda.is_available():
                t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
                t = t - t.mean(dim=0, keepdim=True)
                c0 = (t * t).sum(dim=0)
                f = torch.fft.rfft(t, n=2 * n_t, dim=0)
                p = (f.real ** 2 + f.imag ** 2) / c0
                return torch.fft.irfft(p.mean(dim=1), n=2 * n_t)[:n_t].cpu().numpy()
        except Exception:
            pass
    from scipy import fft as sfft
    n = sfft.next_fast_len(2 * n_t, real=True)
    xc = np.ascontiguousarray((x - np.mean(x, axis=0)).T)            # [n_w, n_t]: transforms along contiguous rows
    c0 = np.einsum("wt,wt->w", xc, xc)
    f = sfft.rfft(xc, n=n, axis=1, workers=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        p = (f.real ** 2 + f.imag ** 2) / c0[:, None]                # a walker that never moved gives nan, as emcee does
    return sfft.irfft(np.mean(p, axis=0), n=n)[:n_t]


def integrated_time(x, c=5, tol=50, quiet=False):
    """Integrated autocorrelation time per dimension, emcee's estimator (Sokal window with c=5;
    autocorrelation averaged over walkers).  x is (n_t, n_w, n_d)."""
    x = np.atleast_1d(x)
    if x.ndim == 1:
        x = x[:, np.newaxis, np.newaxis]
    if x.ndim == 2:
        x = x[:, :, np.newaxis]
    if x.ndim != 3:
        raise ValueError("invalid dimensions")
    n_t, n_w, n_d = x.shape
    tau_est = np.empty(n_d)
    for d in range(n_d):
        f = _mean_norm_acf(x[:, :, d])
        taus = 2.0 * np.cumsum(f) - 1.0
        m = np.arange(len(taus)) < c * taus
        window = np.argmin(m) if np.any(m) else len(taus) - 1
        tau_est[d] = taus[window]
    flag = tol * tau_est > n_t
    if np.any(flag) and not quiet:
        raise AutocorrError(tau_est, "The chain is shorter than %d times the integrated autocorrelation time" % tol)
    return tau_est


def _host(a):
    """NumPy view/copy of a NumPy array or torch tensor."""
    return a if isinstance(a, np.ndarray) else a.detach().cpu().numpy()


class _State(object):
    def __init__(self, coords, log_prob, blobs):
        self.coords, self.log_prob, self.blobs = coords, log_prob, blobs


class EnsembleSampler(object):
    def __init__(self, nwalkers, ndim, log_prob_fn=None, backend=None, args=None, kwargs=None, blobs_dtype=None,
                 a=2.0, engine="host-rng", gp=None, y=None, bounds=None, lnprior_const=0.0, nens=1, seed=None,
                 **unused):
        if nwalkers % 2 or nwalkers < 2 * ndim:
            raise ValueError("emcee requires an even number of walkers, at least twice the dimension")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self.log_prob_fn = log_prob_fn
        self.engine = engine
        self.gp, self.y, self.bounds, self.lnprior_const = gp, y, bounds, lnprior_const
        self.nens = int(nens)
        self.backend = backend
        self.blobs_dtype = blobs_dtype
        # emcee copies the *global* NumPy state at construction
        self._random = np.random.RandomState()
        self._random.set_state(np.random.get_state())
        self._seed = seed
        if engine == "device" and (gp is None or y is None or bounds is None):
            raise ValueError("engine='device' needs gp, y and bounds (box prior)")
        if engine == "host-rng" and log_prob_fn is None:
            raise ValueError("engine='host-rng' needs a vectorised log_prob_fn")
        self._chain = self._logp = self._blobs = None
        self.naccepted = None
        self.iteration = 0

    # ------------------------------------------------------------------ running
    def _run_device(self, p0, nsteps):
        """The chain, log-probabilities and blobs stay on the GPU (torch tensors); ``get_chain`` & co. copy what they
        are asked for, ``get_autocorr_time`` works on the device copy in place."""
        seed = self._seed if self._seed is not None else int(self._random.randint(0, 2 ** 31 - 1))
        out = self.gp.run_ensembles(self.y, p0, nsteps, self.bounds, nens=self.nens, a=self.a, seed=seed,
                                    lnprior_const=self.lnprior_const, device_out=True)
        return out["chain"], out["log_prob"], out["blobs"], out["naccepted"].cpu().numpy().astype(np.int64)

    def _run_host(self, p0, nsteps):
        rng, a = self._random, self.a
        coords = np.array(p0, dtype=np.float64, copy=True)
        nw, nd = coords.shape
        Ns = nw // 2
        lp, blob = self.log_prob_fn(coords)
        lp = np.array(lp, dtype=np.float64)
        blob = np.array(blob, dtype=np.float64)
        if np.any(np.isnan(lp)):
            raise ValueError("The initial log_prob was NaN")
        chain = np.empty((nsteps, nw, nd)); lps = np.empty((nsteps, nw)); blobs = np.empty((nsteps, nw))
        nacc = np.zeros(nw, dtype=np.int64)
        all_inds = np.arange(nw)
        for it in range(nsteps):
            rng.choice(1, p=[1.0])
            inds = all_inds % 2
            rng.shuffle(inds)
            for split in (0, 1):
                S1 = inds == split
                s, c = coords[S1], coords[~S1]
                zz = ((a - 1.0) * rng.rand(Ns) + 1.0) ** 2.0 / a
                factors = (nd - 1.0) * np.log(zz)
                rint = rng.randint(len(c), size=(Ns,))
                q = c[rint] - (c[rint] - s) * zz[:, None]
                nlp, nblob = self.log_prob_fn(q)
                nlp = np.asarray(nlp, dtype=np.float64)
                logu = np.log(rng.rand(Ns))        # same stream as Ns scalar rand() calls
                acc = (factors + nlp - lp[S1]) > logu
                jj = all_inds[S1][acc]
                coords[jj] = q[acc]; lp[jj] = nlp[acc]; blob[jj] = np.asarray(nblob)[acc]
                nacc[jj] += 1
            chain[it] = coords; lps[it] = lp; blobs[it] = blob
        return chain, lps, blobs, nacc

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        p0 = np.asarray(getattr(initial_state, "coords", initial_state), dtype=np.float64)
        p0 = p0.reshape(-1, self.ndim)
        if p0.shape[0] != self.nwalkers * (self.nens if self.engine == "device" else 1):
            raise ValueError("incompatible input dimensions")
        run = self._run_device if self.engine == "device" else self._run_host
        chain, lp, blobs, nacc = run(p0, int(nsteps))
        if self._chain is None:
            self._chain, self._logp, self._blobs, self.naccepted = chain, lp, blobs, nacc
        else:
            cat = np.concatenate
            if not isinstance(chain, np.ndarray):
                import torch
                cat = torch.cat
            self._chain = cat([self._chain, chain]); self._logp = cat([self._logp, lp])
            self._blobs = cat([self._blobs, blobs]); self.naccepted = self.naccepted + nacc
        self.iteration += int(nsteps)
        if self.backend is not None:
            # the reference's chain cache (approx.py:829-833: emcee.backends.HDFBackend(runName + ".h5")): the same
            # group / dataset / attribute layout written by hdf5min (h5py is not installable here), plus an .npz twin
            base = str(self.backend)
            for ext in (".h5", ".npz"):
                if base.endswith(ext):
                    base = base[:-len(ext)]
            chain_h, logp_h, blobs_h = _host(self._chain), _host(self._logp), _host(self._blobs)
            np.savez(base + ".npz", chain=chain_h, log_prob=logp_h, blobs=blobs_h, accepted=self.naccepted)
            from . import hdf5min
            hdf5min.write_emcee_backend(base + ".h5", chain_h, logp_h, blobs=blobs_h, accepted=self.naccepted)
        return _State(_host(self._chain[-1]), _host(self._logp[-1]), _host(self._blobs[-1]))

    def sample(self, initial_state, iterations=1, **kwargs):
        """Generator form used at approx.py:846 (``for _ in sampler.sample(**mcmcKwargs): pass``).

        Limitation (deliberate): the whole chain is produced on the first advance -- on the device engine it is ONE
        kernel launch -- and the same FINAL state is then yielded ``iterations`` times.  Breaking out of the loop
        early does not shorten the run and no intermediate states are exposed; callers that need either should call
        ``run_mcmc`` in chunks.  emcee options that would change what is stored (``thin_by``, ``store=False``,
        ``tune``, ``skip_initial_state_check``) are rejected rather than silently ignored."""
        if kwargs.get("thin_by", 1) not in (None, 1) or kwargs.get("store", True) is not True:
            raise NotImplementedError("sample(thin_by=..., store=False) are not supported; thin with get_chain(thin=)")
        unknown = set(kwargs) - {"thin_by", "store", "progress", "tune", "skip_initial_state_check", "log_prob0", "rstate0", "blobs0"}
        if unknown:
            raise TypeError("sample() got unexpected keyword arguments: %s" % sorted(unknown))
        state = self.run_mcmc(initial_state, iterations)
        for _ in range(int(iterations)):
            yield state

    # ------------------------------------------------------------------ results
    def _get(self, arr, discard=0, flat=False, thin=1):
        if arr is None:
            raise AttributeError("you must run the sampler before accessing the results")
        v = arr[discard + thin - 1::thin]
        if flat:
            v = v.reshape((-1,) + tuple(v.shape[2:]))
        return _host(v)                      # device-resident results: only the requested slice is copied

    def get_chain(self, discard=0, flat=False, thin=1):
        return self._get(self._chain, discard, flat, thin)

    def get_log_prob(self, discard=0, flat=False, thin=1):
        return self._get(self._logp, discard, flat, thin)

    def get_blobs(self, discard=0, flat=False, thin=1):
        b = self._get(self._blobs, discard, flat, thin)
        if self.blobs_dtype is not None:
            out = np.empty(b.shape, dtype=self.blobs_dtype)
            out[out.dtype.names[0]] = b
            return out
        return b

    @property
    def chain(self):            # emcee legacy layout (nwalkers, nsteps, ndim)
        return np.swapaxes(self.get_chain(), 0, 1)

    @property
    def acceptance_fraction(self):
        return self.naccepted / float(self.iteration)

    def get_autocorr_time(self, discard=0, thin=1, c=5, tol=50, quiet=False, **kwargs):
        """emcee's ``get_autocorr_time``.  A device-resident chain (engine="device") is analysed where it is by the
        engine's direct-sum kernel (GP.integrated_time); host chains, and series too long for that kernel, take the
        FFT estimator above."""
        tau = None
        if self._chain is not None and not isinstance(self._chain, np.ndarray) and self.gp is not None:
            res = self.gp.integrated_time(self._chain, discard=discard, thin=thin, c=c)
            if res is not None:
                tau = res[0]
                n_t = len(range(discard + thin - 1, self._chain.shape[0], thin))
                if np.any(tol * tau > n_t) and not quiet and tol > 0:
                    raise AutocorrError(tau, "The chain is shorter than %d times the integrated autocorrelation time" % tol)
        if tau is None:
            tau = integrated_time(self.get_chain(discard=discard, thin=thin), c=c, tol=tol, quiet=quiet)
        return thin * tau
