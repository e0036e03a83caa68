"""The reference's OWN test modules, unmodified, executed through george/emcee shims.

* ``-m gpu`` half: ``approxposterior_b200.compat.install()`` -- ``import george`` / ``import emcee`` inside the
  reference resolve to the engine, so ``approxposterior/tests/test_*.py`` (reference files, not restatements) run on
  ``libapgp.so``.  This is what "drops in behind gpUtils.defaultGP" means (gpUtils.py:160-178, approx.py:712-717,
  839-847).
* CPU half: the same modules on the oracle-backed shim (``oracle/refshim.py``) -- a second pin of the oracle, this
  time driven by the reference's own code rather than by this repo's restated drivers.

The reference package is imported from ``/root/reference`` (authoring container) or from ``baseline/_ref`` (the
``pip install --no-deps --target baseline/_ref`` copy that travels to the GPU box); both are the unmodified v0.4
sources.  If neither exists the tests skip.

One environment patch, applied to both halves and to nothing else: ``oracle.refshim.scipy_x0_compat`` flattens the 2-D
``x0`` the reference hands to ``scipy.optimize.minimize`` (utility.py:336,364), which SciPy >= 1.11 rejects and the
SciPy of the reference's day flattened itself.  It touches only the reference tests that reach ``minimizeObjective``.
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_root():
    forced = os.environ.get("APGP_REFERENCE_ROOT")
    for cand in ([forced] if forced else ["/root/reference", os.path.join(ROOT, "baseline", "_ref")]):
        if os.path.isfile(os.path.join(cand, "approxposterior", "tests", "test_GPUtil.py")):
            return cand
    return None


REF = _reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="reference package not available (/root/reference or baseline/_ref)")

# reference test module -> functions; (module, function, needs the SciPy x0 patch)
CASES = [
    ("test_GPUtil", "testUtilsGPAmp"), ("test_GPUtil", "testUtilsGPNoAmp"),
    ("test_InitGP", "testInitGPAmp"), ("test_InitGP", "testInitGPNoAmp"),
    ("test_OptimizeGP", "testGPOptAmp"), ("test_OptimizeGP", "testGPOptNoAmp"),
    ("test_Burnin", "testBurnin"), ("test_MCSE", "testMCSE"), ("test_TestFns", "testTestFns"),
    ("test_Import", "test_import"),
    ("test_findNewPoint", "testFindNoAmp"), ("test_MAP", "testMAPAmp"),
    ("test_1DBayesOpt", "test_1DBO"), ("test_2DBayesOpt", "test_2DBO"), ("test_APRun", "testRun"),
]
# tests/test_findNewPoint.py:60 (amplitude case) depends on the Nelder-Mead path over a multi-modal surface and is not
# reproduced by SciPy 1.18 on ANY backend: run, but do not require
XFAIL = [("test_findNewPoint", "testFindAmp")]


def _purge():
    for name in [n for n in sys.modules if n == "approxposterior" or n.startswith("approxposterior.")]:
        del sys.modules[name]


class _Reference(object):
    """Context: shims installed, reference importable, module cache isolated, cwd = a scratch directory
    (the reference's drivers write their .npz caches into the working directory)."""

    def __init__(self, shim, tmpdir):
        self.shim, self.tmpdir = shim, str(tmpdir)

    def __enter__(self):
        _purge()
        self.shim.install()
        sys.path.insert(0, REF)
        self.cwd = os.getcwd()
        os.chdir(self.tmpdir)
        from oracle.refshim import scipy_x0_compat
        ut = importlib.import_module("approxposterior.utility")
        scipy_x0_compat(ut)
        return self

    def run(self, module, function):
        mod = importlib.import_module("approxposterior.tests." + module)
        assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF)), mod.__file__
        with np.errstate(all="ignore"):
            getattr(mod, function)()

    def __exit__(self, *exc):
        os.chdir(self.cwd)
        sys.path.remove(REF)
        self.shim.uninstall()
        _purge()
        return False


# the reference's two remaining test modules never reach george / emcee (gmmUtils, klNumerical: out of scope for the
# engine); with them the CPU half executes EVERY test module the reference ships
OFF_PATH = [("test_KL", "testKLApproximation"), ("test_GMM", "testGMMFit")]


# ------------------------------------------------------------------------------------------ CPU: oracle-backed shim
@needs_ref
@pytest.mark.parametrize("module,function", CASES + OFF_PATH)
def test_reference_tests_on_oracle_shim(module, function, tmp_path):
    from oracle import refshim
    with _Reference(refshim, tmp_path) as ref:
        ref.run(module, function)


@needs_ref
@pytest.mark.parametrize("module,function", XFAIL)
def test_reference_optimizer_path_dependent_on_oracle_shim(module, function, tmp_path):
    from oracle import refshim
    with _Reference(refshim, tmp_path) as ref:
        try:
            ref.run(module, function)
        except AssertionError:
            pytest.xfail("optimiser-path dependent golden (SURVEY 4.3)")


@needs_ref
@pytest.mark.parametrize("theta0", [None, [0.4, 0.9]])
def test_sequential_minimize_objective_is_the_references_rng_order(theta0, tmp_path):
    """the side-by-side engines draw all restart starts up front; the reference draws start r only after
    restart r - 1 and its retries (utility.py:336,364).  The sequential path (batched=False) follows the reference
    exactly: with a prior that rejects part of the domain (so that restarts ARE retried) the same seed gives the
    same optimum, the same objective value and leaves np.random in the same state as the reference's own function."""
    from oracle import refshim
    from approxposterior_b200 import utility as mine, likelihood as lh
    with _Reference(refshim, tmp_path):
        rut = importlib.import_module("approxposterior.utility")
        rgu = importlib.import_module("approxposterior.gpUtils")
        np.random.seed(3)
        theta = lh.rosenbrockSample(40)
        y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
        gp = rgu.defaultGP(theta, y, white_noise=-12)
        calls = {"rejected": 0, "seen": []}

        def box(x):                                    # what the utility itself sees
            return 0.0 if np.all(np.abs(np.asarray(x).ravel()) <= 5.0) else -np.inf

        def prior(x):                                  # what validates an optimum: the box minus a band optima end up in
            x = np.asarray(x).ravel()
            calls["seen"].append(x.copy())
            ok = np.all(np.abs(x) <= 5.0) and not (-1.6 < x[0] < 2.5)
            calls["rejected"] += 0 if ok else 1
            return 0.0 if ok else -np.inf
        kw = dict(nRestarts=4, method="nelder-mead", options={"adaptive": True, "maxiter": 40}, theta0=theta0,
                  args=(y, gp, box))
        np.random.seed(11)
        with np.errstate(all="ignore"):
            x_ref, f_ref = rut.minimizeObjective(rut.BAPEUtility, y, gp, lh.rosenbrockSample, prior, **kw)
        state_ref = np.random.get_state()[1].copy()
        rejected_ref, seen_ref = calls["rejected"], calls["seen"]
        calls.update(rejected=0, seen=[])
        np.random.seed(11)
        with np.errstate(all="ignore"):
            x_me, f_me = mine.minimizeObjective(rut.BAPEUtility, y, gp, lh.rosenbrockSample, prior, batched=False, **kw)
        assert rejected_ref > 0, "the prior never rejected an optimum: the test would not see the order"
        assert calls["rejected"] == rejected_ref
        assert len(calls["seen"]) == len(seen_ref) and all(np.array_equal(a, b) for a, b in zip(calls["seen"], seen_ref)), \
            "every optimum handed to the prior, in order: restarts and retries started from the same points"
        assert np.array_equal(np.asarray(x_me).ravel(), np.asarray(x_ref).ravel()) and f_me == f_ref
        assert np.array_equal(np.random.get_state()[1], state_ref)


# ------------------------------------------------------------------------------------------ GPU: engine-backed shim
@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("module,function", CASES)
def test_reference_tests_on_engine(module, function, tmp_path):
    from approxposterior_b200 import compat
    import approxposterior_b200._lib as _lib
    with _Reference(compat, tmp_path) as ref:
        george = sys.modules["george"]
        from approxposterior_b200 import GP
        assert george.GP is GP
        ref.run(module, function)
    assert _lib._lib is not None, "libapgp.so was not loaded"


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("module,function", XFAIL)
def test_reference_optimizer_path_dependent_on_engine(module, function, tmp_path):
    from approxposterior_b200 import compat
    with _Reference(compat, tmp_path) as ref:
        try:
            ref.run(module, function)
        except AssertionError:
            pytest.xfail("optimiser-path dependent golden (SURVEY 4.3)")


@pytest.mark.gpu
@needs_ref
@pytest.mark.parametrize("module,function", [("test_OptimizeGP", "testGPOptAmp"), ("test_OptimizeGP", "testGPOptNoAmp"),
                                             ("test_findNewPoint", "testFindNoAmp"), ("test_MAP", "testMAPAmp"),
                                             ("test_1DBayesOpt", "test_1DBO"), ("test_2DBayesOpt", "test_2DBO"),
                                             ("test_APRun", "testRun")])
def test_reference_tests_on_engine_accelerated(module, function, tmp_path):
    """compat.accelerate(): the reference's source files are still untouched, but its two multistart drivers resolve to
    the engine's batched ones (every optGP restart in one device launch, utility restarts in lock step).  The
    reference's own optimiser-level known answers must still come out."""
    from approxposterior_b200 import compat, gpUtils
    with _Reference(compat, tmp_path) as ref:
        compat.accelerate()
        gpUtils.optimizeGP.last_stats = None
        ref.run(module, function)
        import approxposterior
        assert approxposterior.gpUtils.optimizeGP is gpUtils.optimizeGP
        if module != "test_findNewPoint":                      # (findNextPoint(computeLnLike=False) never refits)
            assert gpUtils.optimizeGP.last_stats is not None and gpUtils.optimizeGP.last_stats["scheduler"] == "device"


@pytest.mark.gpu
@needs_ref
def test_reference_run_uses_the_engine_for_every_hot_call(tmp_path):
    """The reference's ApproxPosterior.run (README configuration, shortened) on the engine: the GP object the
    reference holds is the engine's, the kernels were launched, and the batched _gpll seam of the emcee shim
    returns what the reference's own scalar _gpll returns."""
    from approxposterior_b200 import compat, GP
    with _Reference(compat, tmp_path):
        import approxposterior as apx
        from approxposterior import approx, gpUtils, likelihood as lh
        np.random.seed(57)
        theta = lh.rosenbrockSample(50)
        y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
        gp = gpUtils.defaultGP(theta, y, white_noise=-12)
        assert isinstance(gp, GP)
        ap = approx.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=lh.rosenbrockLnprior, lnlike=lh.rosenbrockLnlike,
                                    priorSample=lh.rosenbrockSample, bounds=[(-5, 5), (-5, 5)], algorithm="bape")
        before = gp.launch_count
        ap.run(m=3, nmax=1, estBurnin=True, nGPRestarts=2, mcmcKwargs={"iterations": 400}, cache=False,
               samplerKwargs={"nwalkers": 20}, verbose=False, thinChains=False, onlyLastMCMC=True)
        assert isinstance(ap.gp, GP) and ap.gp.launch_count > 0 and before > 0
        assert ap.sampler.get_chain().shape == (400, 20, 2)
        assert ap.theta.shape == (53, 2) and np.all(np.isfinite(ap.y))
        q = np.vstack([lh.rosenbrockSample(6), [[7.0, 0.0]], [[np.nan, np.nan]]])
        lp_b, blob_b = ap.sampler.log_prob_fn(q)
        for i, t in enumerate(q):
            lp_s, blob_s = ap._gpll(t)
            if np.isfinite(lp_s):
                assert abs(lp_b[i] - float(lp_s)) <= 1e-12 * max(1.0, abs(float(lp_s))) and blob_b[i] == blob_s
            else:
                assert lp_b[i] == -np.inf and np.isnan(blob_b[i])
        assert apx.__version__ == "0.4"


# ------------------------------------------------------------------------------------------ CPU: engine shim, host parts
@needs_ref
def test_engine_shim_installs_and_serves_the_emcee_surface_on_cpu(tmp_path):
    """compat.install() without a GPU: the reference imports, george.GP is the engine's class (creating one would need
    the device), and the emcee shim -- emcee 3.0's signature on the engine's NumPy-RNG sampler -- reproduces the
    reference's own burn-in known answer (tests/test_Burnin.py:17-91, scalar log_prob_fn with args) and MCSE test."""
    from approxposterior_b200 import compat
    with _Reference(compat, tmp_path) as ref:
        import approxposterior_b200
        assert sys.modules["george"].GP is approxposterior_b200.GP
        assert sys.modules["george"].kernels.ExpSquaredKernel is approxposterior_b200.kernels.ExpSquaredKernel
        assert int(sys.modules["emcee"].__version__.split(".")[0]) > 2
        ref.run("test_Import", "test_import")
        ref.run("test_Burnin", "testBurnin")
        ref.run("test_MCSE", "testMCSE")
        ref.run("test_TestFns", "testTestFns")
        be = sys.modules["emcee"].backends.HDFBackend(str(tmp_path / "chain.h5"))
        be.reset(8, 2)
        s = sys.modules["emcee"].EnsembleSampler(8, 2, lambda t: -0.5 * float(np.sum(np.asarray(t) ** 2)), backend=be)
        np.random.seed(1)
        for _ in s.sample(initial_state=np.random.randn(8, 2), iterations=30):
            pass
        assert s.get_chain().shape == (30, 8, 2) and (tmp_path / "chain.npz").exists()
        from approxposterior_b200 import hdf5min                      # the file the reference asked for, emcee's layout
        h5 = hdf5min.read_emcee_backend(str(tmp_path / "chain.h5"))
        assert np.array_equal(h5["chain"], s.get_chain()) and int(h5["attrs"]["iteration"]) == 30
        with pytest.raises(NotImplementedError):
            for _ in s.sample(initial_state=np.random.randn(8, 2), iterations=3, thin_by=2):
                pass
