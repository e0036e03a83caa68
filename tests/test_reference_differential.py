"""Differential tests of this repo's host drivers against the REFERENCE'S OWN functions (CPU).

The reference's pure-Python modules import once `george` / `emcee` resolve to the oracle-backed shim
(oracle/refshim.py), so its drivers can be called side by side with the mirrors in approxposterior_b200 on the SAME GP
object and the same seeds.  Everything here must agree bit for bit -- the mirrors claim the reference's protocol,
RNG consumption included (gpUtils.py:22-110, 184-257; utility.py:69-250; mcmcUtils.py:103-227)."""
import importlib

import numpy as np
import pytest

from test_reference_dropin import _Reference, needs_ref


def _problem(n=40, seed=3):
    from approxposterior_b200 import likelihood as lh
    np.random.seed(seed)
    theta = lh.rosenbrockSample(n)
    y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
    return theta, y


@needs_ref
@pytest.mark.parametrize("fitAmp", [False, True])
@pytest.mark.parametrize("method", ["powell", "nelder-mead"])
def test_optimizeGP_sequential_is_the_references(fitAmp, method, tmp_path):
    from oracle import refshim
    from approxposterior_b200 import gpUtils as mine
    with _Reference(refshim, tmp_path):
        rgu = importlib.import_module("approxposterior.gpUtils")
        theta, y = _problem()
        np.random.seed(5); g_ref = rgu.defaultGP(theta, y, white_noise=-12, fitAmp=fitAmp)
        np.random.seed(5); g_me = rgu.defaultGP(theta, y, white_noise=-12, fitAmp=fitAmp)
        np.random.seed(9)
        with np.errstate(all="ignore"):
            a = rgu.optimizeGP(g_ref, theta, y, nGPRestarts=3, method=method)
        s_ref = np.random.get_state()[1].copy()
        np.random.seed(9)
        with np.errstate(all="ignore"):
            b = mine.optimizeGP(g_me, theta, y, nGPRestarts=3, method=method, batched=False)
        assert np.array_equal(a.get_parameter_vector(), b.get_parameter_vector())
        assert np.array_equal(np.random.get_state()[1], s_ref)
        assert a.log_likelihood(y, quiet=True) == b.log_likelihood(y, quiet=True)


@needs_ref
def test_nll_grad_and_hyper_prior_are_the_references(tmp_path):
    from oracle import refshim
    from approxposterior_b200 import gpUtils as mine
    with _Reference(refshim, tmp_path):
        rgu = importlib.import_module("approxposterior.gpUtils")
        theta, y = _problem()
        for fitAmp in (False, True):
            np.random.seed(5)
            gp = rgu.defaultGP(theta, y, white_noise=-12, fitAmp=fitAmp)
            rng = np.random.default_rng(1)
            p0 = gp.get_parameter_vector()
            for k in range(12):
                p = p0 + rng.normal(scale=[0.5, 3.0, 30.0][k % 3], size=p0.size)      # inside and far outside the hyper-prior
                assert rgu.defaultHyperPrior(p) == mine.defaultHyperPrior(p)
                for prior in (None, "default"):
                    pr, pm = (None, None) if prior is None else (rgu.defaultHyperPrior, mine.defaultHyperPrior)
                    with np.errstate(all="ignore"):
                        assert rgu._nll(p, gp, y, pr) == mine._nll(p, gp, y, pm)
                        assert np.array_equal(rgu._grad_nll(p, gp, y, pr), mine._grad_nll(p, gp, y, pm))


@needs_ref
def test_utilities_and_logsubexp_are_the_references(tmp_path):
    from oracle import refshim
    from approxposterior_b200 import utility as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rut = importlib.import_module("approxposterior.utility")
        rgu = importlib.import_module("approxposterior.gpUtils")
        theta, y = _problem()
        np.random.seed(5)
        gp = rgu.defaultGP(theta, y, white_noise=-12)
        rng = np.random.default_rng(2)
        pts = np.vstack([rng.uniform(-5, 5, size=(40, 2)), theta[:5], [[6.0, 0.0], [0.0, -5.5]]])   # incl. training points, outside the prior
        with np.errstate(all="ignore"):
            for t in pts:
                for name in ("AGPUtility", "BAPEUtility", "JonesUtility"):
                    a = getattr(rut, name)(t, y, gp, lh.rosenbrockLnprior)
                    b = getattr(mine, name)(t, y, gp, lh.rosenbrockLnprior)
                    assert np.array_equal(np.asarray(a, dtype=float).ravel(), np.asarray(b, dtype=float).ravel(), equal_nan=True), (name, t)
            for x1, x2 in rng.normal(scale=20.0, size=(50, 2)):
                a, b = rut.logsubexp(x1, x2), mine.logsubexp(x1, x2)
                assert (a == b) or (np.isnan(a) and np.isnan(b))
        x = np.linspace(0.1, 3.0, 50)
        p = lambda v: np.exp(-0.5 * v * v) / np.sqrt(2 * np.pi)
        q = lambda v: np.exp(-0.5 * (v - 0.3) ** 2 / 1.44) / np.sqrt(2 * np.pi * 1.44)
        assert rut.klNumerical(x, p, q) == mine.klNumerical(x, p, q)


@needs_ref
def test_mcse_and_burnin_are_the_references(tmp_path):
    """batchMeansMCSE on the same samples; estimateBurnin on the same emcee-flow sampler object (oracle shim)."""
    from oracle import refshim
    from approxposterior_b200 import mcmcUtils as mine
    with _Reference(refshim, tmp_path):
        rmc = importlib.import_module("approxposterior.mcmcUtils")
        import emcee                                           # the shim
        rng = np.random.default_rng(4)
        samples = rng.normal(size=(5000, 3)).cumsum(axis=0) * 0.01 + rng.normal(size=(5000, 3))
        assert np.array_equal(rmc.batchMeansMCSE(samples), mine.batchMeansMCSE(samples))
        assert np.array_equal(rmc.batchMeansMCSE(samples, bins=20, fn=np.square), mine.batchMeansMCSE(samples, bins=20, fn=np.square))
        np.random.seed(42)
        lnprob = lambda x: -0.5 * np.sum(x * x)
        sampler = emcee.EnsembleSampler(10, 2, lnprob)
        for _ in sampler.sample(np.random.randn(10, 2), iterations=3000):
            pass
        for est, thin in ((True, True), (True, False), (False, True), (False, False)):
            assert tuple(rmc.estimateBurnin(sampler, estBurnin=est, thinChains=thin)) == \
                tuple(mine.estimateBurnin(sampler, estBurnin=est, thinChains=thin))


def _build_ap(mod, lhmod, rgu, algorithm, n0=20):
    np.random.seed(3)
    theta = lhmod.rosenbrockSample(n0)
    y = np.array([lhmod.rosenbrockLnlike(t) + lhmod.rosenbrockLnprior(t) for t in theta])
    np.random.seed(5)
    gp = rgu.defaultGP(theta, y, white_noise=-12)
    return mod.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=lhmod.rosenbrockLnprior, lnlike=lhmod.rosenbrockLnlike,
                               priorSample=lhmod.rosenbrockSample, bounds=[(-5, 5), (-5, 5)], algorithm=algorithm)


@needs_ref
@pytest.mark.parametrize("algorithm", ["bape", "agp", "alternate"])
def test_run_is_the_references_run_on_the_same_gp(algorithm, tmp_path):
    """ApproxPosterior.run of the mirror against the reference's OWN ApproxPosterior.run (approx.py:229-524), both on the
    oracle-backed george/emcee shim with the same seeds: the same design points and targets, the same burn-in / thinning
    estimates, the same final chain, bit for bit -- the mirror keeps the reference's protocol and its np.random
    consumption through findNextPoint, optGP, runMCMC and estimateBurnin."""
    from oracle import refshim
    from approxposterior_b200 import approx as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rap = importlib.import_module("approxposterior.approx")
        rgu = importlib.import_module("approxposterior.gpUtils")
        rlh = importlib.import_module("approxposterior.likelihood")
        def kw():          # fresh dictionaries per run: the reference's validateMCMCKwargs writes into mcmcKwargs
            return dict(m=4, nmax=2, estBurnin=True, nGPRestarts=2, mcmcKwargs={"iterations": 300}, cache=False, verbose=False,
                        thinChains=True, onlyLastMCMC=False, seed=21)
        a = _build_ap(rap, rlh, rgu, algorithm)
        np.random.seed(21)
        with np.errstate(all="ignore"):
            a.run(samplerKwargs={"nwalkers": 10}, **kw())
        state_ref = np.random.get_state()[1].copy()
        b = _build_ap(mine, lh, rgu, algorithm)
        np.random.seed(21)
        with np.errstate(all="ignore"):
            b.run(samplerKwargs={"nwalkers": 10, "engine": "host-rng"}, **kw())
        assert a.theta.shape == (28, 2)
        assert np.array_equal(a.theta, b.theta) and np.array_equal(a.y, b.y)
        assert list(a.iburns) == list(b.iburns) and [int(v) for v in a.ithins] == [int(v) for v in b.ithins]
        assert np.array_equal(a.sampler.get_chain(), b.sampler.get_chain())
        assert np.array_equal(a.gp.get_parameter_vector(), b.gp.get_parameter_vector())
        assert np.array_equal(np.random.get_state()[1], state_ref)


@needs_ref
def test_run_with_non_default_options_is_the_references(tmp_path):
    """The same comparison on the branches the default call does not take: GP refits every second point only
    (optGPEveryN), Nelder-Mead for the hyper-parameters, Powell for the utility, no initial fit, no burn-in estimate,
    convergence check with a loose eps (the run may stop early -- on both sides at the same iteration)."""
    from oracle import refshim
    from approxposterior_b200 import approx as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rap = importlib.import_module("approxposterior.approx")
        rgu = importlib.import_module("approxposterior.gpUtils")
        rlh = importlib.import_module("approxposterior.likelihood")

        def kw():
            return dict(m=5, nmax=3, estBurnin=False, thinChains=False, nGPRestarts=1, mcmcKwargs={"iterations": 150}, cache=False,
                        verbose=False, seed=4, optGPEveryN=2, gpMethod="nelder-mead", gpOptions={"maxiter": 60}, initGPOpt=False,
                        nMinObjRestarts=2, minObjMethod="powell", minObjOptions={"maxiter": 3}, convergenceCheck=True, eps=5.0,
                        kmax=1)
        a = _build_ap(rap, rlh, rgu, "agp")
        np.random.seed(4)
        with np.errstate(all="ignore"):
            a.run(samplerKwargs={"nwalkers": 8}, **kw())
        state_ref = np.random.get_state()[1].copy()
        b = _build_ap(mine, lh, rgu, "agp")
        np.random.seed(4)
        with np.errstate(all="ignore"):
            b.run(samplerKwargs={"nwalkers": 8, "engine": "host-rng"}, **kw())
        assert a.theta.shape == b.theta.shape and np.array_equal(a.theta, b.theta) and np.array_equal(a.y, b.y)
        assert np.array_equal(a.sampler.get_chain(), b.sampler.get_chain())
        assert np.array_equal(a.gp.get_parameter_vector(), b.gp.get_parameter_vector())
        assert np.array_equal(np.random.get_state()[1], state_ref)


@needs_ref
def test_bayesopt_and_findmap_are_the_references(tmp_path):
    """bayesOpt (approx.py:929-1151, Jones utility) and findMAP (approx.py:857-926) against the reference's own."""
    from oracle import refshim
    from approxposterior_b200 import approx as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rap = importlib.import_module("approxposterior.approx")
        rgu = importlib.import_module("approxposterior.gpUtils")
        rlh = importlib.import_module("approxposterior.likelihood")
        out = []
        for mod, lhmod in ((rap, rlh), (mine, lh)):
            ap = _build_ap(mod, lhmod, rgu, "jones", n0=15)
            np.random.seed(8)
            with np.errstate(all="ignore"):
                soln = ap.bayesOpt(nmax=4, verbose=False, cache=False, nGPRestarts=2, nMinObjRestarts=3, findMAP=True, seed=8)
                m, v = ap.findMAP(nRestarts=3)
            out.append((ap.theta.copy(), ap.y.copy(), soln, np.asarray(m), np.asarray(v), np.random.get_state()[1].copy()))
        (ta, ya, sa, ma, va, ra), (tb, yb, sb, mb, vb, rb) = out
        assert np.array_equal(ta, tb) and np.array_equal(ya, yb)
        assert sorted(sa.keys()) == sorted(sb.keys())
        for k in sa:
            assert np.array_equal(np.asarray(sa[k], dtype=float), np.asarray(sb[k], dtype=float), equal_nan=True), k
        assert np.array_equal(ma, mb), (ma, mb)
        assert np.array_equal(np.ravel(va), np.ravel(vb)), (va, vb)
        assert np.array_equal(ra, rb)


@needs_ref
def test_cache_files_are_the_references(tmp_path):
    """N3: the four .npz caches ApproxPosterior.run writes (approx.py:363-364, 430-432, 469-473, 507-514) -- same file
    names, same keys, same contents as the reference's own run writes (timings aside); the mirror additionally writes the
    chain files runName{i}.h5 (emcee's HDFBackend layout, hdf5min.py), which the reference delegates to emcee + h5py."""
    import os
    from oracle import refshim
    from approxposterior_b200 import approx as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rap = importlib.import_module("approxposterior.approx")
        rgu = importlib.import_module("approxposterior.gpUtils")
        rlh = importlib.import_module("approxposterior.likelihood")
        dirs = []
        for mod, lhmod, extra in ((rap, rlh, {}), (mine, lh, {"engine": "host-rng"})):
            d = tmp_path / ("ref" if mod is rap else "mirror")
            d.mkdir()
            dirs.append(d)
            ap = _build_ap(mod, lhmod, rgu, "bape")
            np.random.seed(21)
            sk = {"nwalkers": 10}
            sk.update(extra)
            with np.errstate(all="ignore"):
                ap.run(m=3, nmax=2, estBurnin=True, nGPRestarts=1, mcmcKwargs={"iterations": 200}, samplerKwargs=sk, cache=True,
                       verbose=False, thinChains=True, convergenceCheck=True, timing=True, seed=21, runName=str(d / "apRun"))
        ref_files = sorted(f for f in os.listdir(dirs[0]) if f.endswith(".npz"))
        assert ref_files == ["apRunAPFModelCache.npz", "apRunAPGP.npz", "apRunAPTiming.npz", "apRunConvergenceCache.npz"]
        for f in ref_files:
            A, B = np.load(dirs[0] / f, allow_pickle=True), np.load(dirs[1] / f, allow_pickle=True)
            assert sorted(A.files) == sorted(B.files), f
            for k in A.files:
                if f == "apRunAPTiming.npz":
                    assert A[k].shape == B[k].shape, (f, k)
                else:
                    assert np.array_equal(A[k], B[k]), (f, k)
        assert (dirs[1] / "apRun0.h5").exists() and (dirs[1] / "apRun1.h5").exists()


@needs_ref
def test_fixture_functions_are_the_references(tmp_path):
    """likelihood.py's fixtures (the objectives, priors and prior samplers the reference's tests and README use): same
    values on the same points, same draws and np.random state from the same seed."""
    from oracle import refshim
    from approxposterior_b200 import likelihood as mine
    with _Reference(refshim, tmp_path):
        rlh = importlib.import_module("approxposterior.likelihood")
        rng = np.random.default_rng(6)
        for name, dim in (("rosenbrock", 2), ("sphere", 2), ("testBOFn", 1)):
            lnlike = {"rosenbrock": "rosenbrockLnlike", "sphere": "sphereLnlike", "testBOFn": "testBOFn"}[name]
            lnprior = {"rosenbrock": "rosenbrockLnprior", "sphere": "sphereLnprior", "testBOFn": "testBOFnLnPrior"}[name]
            sample = {"rosenbrock": "rosenbrockSample", "sphere": "sphereSample", "testBOFn": "testBOFnSample"}[name]
            pts = rng.uniform(-6, 6, size=(60, dim))                        # inside and outside the prior supports
            with np.errstate(all="ignore"):
                for t in pts:
                    t = t if dim > 1 else t[0]
                    for fn in (lnlike, lnprior):
                        a, b = getattr(rlh, fn)(t), getattr(mine, fn)(t)
                        assert np.array_equal(np.asarray(a, dtype=float), np.asarray(b, dtype=float), equal_nan=True), (fn, t)
            for n in (1, 7):
                np.random.seed(12); a = getattr(rlh, sample)(n); sa = np.random.get_state()[1].copy()
                np.random.seed(12); b = getattr(mine, sample)(n); sb = np.random.get_state()[1].copy()
                assert np.array_equal(np.asarray(a), np.asarray(b)) and np.shape(a) == np.shape(b) and np.array_equal(sa, sb), sample
        with np.errstate(all="ignore"):
            for t in rng.uniform(-6, 6, size=(20, 2)):
                assert rlh.rosenbrockLnprob(t) == mine.rosenbrockLnprob(t)


@needs_ref
def test_validate_mcmc_kwargs_is_the_references(tmp_path):
    """mcmcUtils.validateMCMCKwargs (mcmcUtils.py:15-100): the same sampler / mcmc dictionaries, the same prior draws for
    the initial state, on every branch the reference itself survives (its samplerKwargs=None branch reads a non-existent
    "dim" key -- SURVEY App. B -- and is the one place the mirror deliberately differs)."""
    from oracle import refshim
    from approxposterior_b200 import mcmcUtils as mine, likelihood as lh

    class AP(object):
        ndim = 2
        priorSample = staticmethod(lh.rosenbrockSample)

        def _gpll(self, theta, *a, **k):
            return 0.0, 0.0
    with _Reference(refshim, tmp_path):
        rmc = importlib.import_module("approxposterior.mcmcUtils")
        ap = AP()
        cases = [({"nwalkers": 12}, None), ({"nwalkers": 8, "ndim": 5, "backend": "x", "log_prob_fn": len}, {"iterations": 50}),
                 ({}, {"initial_state": np.zeros((40, 2))}), ({"nwalkers": 6}, {"iterations": 7, "initial_state": np.ones((6, 2))})]
        for sk, mk in cases:
            out = []
            for mod in (rmc, mine):
                np.random.seed(2)
                s, m = mod.validateMCMCKwargs(ap, dict(sk), None if mk is None else dict(mk), verbose=False)
                out.append((s, m, np.random.get_state()[1].copy()))
            (sa, ma, ra), (sb, mb, rb) = out
            assert sorted(sa) == sorted(sb) and sorted(ma) == sorted(mb)
            for k in sa:
                if k == "log_prob_fn":
                    assert sa[k] == ap._gpll and sb[k] == ap._gpll
                else:
                    assert sa[k] == sb[k], k
            for k in ma:
                assert np.array_equal(np.asarray(ma[k]), np.asarray(mb[k])), k
            assert np.array_equal(ra, rb)


@needs_ref
def test_argument_errors_are_the_references(tmp_path):
    """Constructor / run argument validation (approx.py:84-129, 390-394): the same exception types for the same bad
    arguments; the same objects accepted."""
    from oracle import refshim
    from approxposterior_b200 import approx as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rap = importlib.import_module("approxposterior.approx")
        rgu = importlib.import_module("approxposterior.gpUtils")
        rlh = importlib.import_module("approxposterior.likelihood")
        theta, y = _problem(12)
        np.random.seed(5)
        gp = rgu.defaultGP(theta, y, white_noise=-12)
        bad_theta = theta.copy(); bad_theta[0, 0] = np.inf
        bad_y = y.copy(); bad_y[3] = np.nan
        cases = [dict(theta=None, y=y), dict(theta=theta, y=None), dict(theta=bad_theta, y=y), dict(theta=theta, y=bad_y),
                 dict(theta=theta, y=y, bounds=[(-5, 5)]), dict(theta=theta, y=y, algorithm="nope"),
                 dict(theta=theta, y=y, algorithm="BAPE"), dict(theta=theta, y=y, algorithm="jones"), dict(theta=theta, y=y)]
        for c in cases:
            res = []
            for mod, lhmod in ((rap, rlh), (mine, lh)):
                kw = dict(gp=gp, lnprior=lhmod.rosenbrockLnprior, lnlike=lhmod.rosenbrockLnlike, priorSample=lhmod.rosenbrockSample,
                          bounds=[(-5, 5), (-5, 5)], algorithm="bape")
                kw.update(c)
                try:
                    ap = mod.ApproxPosterior(**kw)
                    res.append(("ok", ap.algorithm, ap.theta.shape, ap.y.shape))
                except Exception as e:                       # noqa: BLE001 -- the comparison is the point
                    res.append((type(e).__name__,))           # (the reference's "unknown algorithm" text lists "naive", which it
            assert res[0] == res[1], (c.keys(), res)             # does not accept, instead of "jones": messages are not compared)
        for mod, lhmod in ((rap, rlh), (mine, lh)):          # convergenceCheck needs an MCMC per iteration
            ap = mod.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=lhmod.rosenbrockLnprior, lnlike=lhmod.rosenbrockLnlike,
                                     priorSample=lhmod.rosenbrockSample, bounds=[(-5, 5), (-5, 5)], algorithm="bape")
            with pytest.raises(RuntimeError):
                ap.run(m=1, nmax=1, convergenceCheck=True, onlyLastMCMC=True, verbose=False, cache=False)
