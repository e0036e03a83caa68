"""Fused cluster-per-restart log-likelihood (csrc/chol_group.cuh: apgp_loglik_batch beyond N ~ 224 and the device
optimisers' objective there) against the CPU oracle, the multi-launch tiled path, itself at every cluster size, and the
SciPy restatements of the optimisers.  Reference: gpUtils._nll (gpUtils.py:46-80) inside optimizeGP (gpUtils.py:223-247)."""
import numpy as np
import pytest

from conftest import rosenbrock_training, synthetic_gp_problem

pytestmark = pytest.mark.gpu


def _pair(N, d, amp=None, seed=0):
    from approxposterior_b200 import GP, kernels
    from oracle import GPOracle
    if d == 2:
        X, y = rosenbrock_training(N)
        logM = np.zeros(2)
    else:
        X, y, logM, _ = synthetic_gp_problem(N, d, seed=seed)
    k = kernels.ExpSquaredKernel(np.exp(logM), ndim=d)
    if amp is not None:
        k = amp * k
    gp = GP(kernel=k, fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    orc = GPOracle(d, np.exp(logM), amp=amp, mean=float(np.median(y)), white_noise=-12.0)
    orc.compute(X)
    return gp, orc, X, y


@pytest.mark.parametrize("N,d,amp,R", [(230, 2, None, 5), (256, 2, 3.0, 64), (300, 10, 2.0, 9), (512, 10, 1.5, 64),
                                       (640, 5, None, 3), (1000, 2, None, 2), (1024, 10, 2.5, 8), (1100, 3, None, 1)])
def test_loglik_group_vs_oracle_and_tiled(N, d, amp, R, monkeypatch):
    gp, orc, X, y = _pair(N, d, amp=amp, seed=N)
    rng = np.random.default_rng(N + R)
    P = np.column_stack([np.full(R, np.median(y))] + [0.6 * rng.standard_normal(R) for _ in range(len(gp) - 1)])
    P += gp.get_parameter_vector()[None, :] * np.r_[0.0, np.ones(len(gp) - 1)][None, :]
    if R > 2:
        P[1, -1] = np.nan                 # non-finite vector -> -inf
        P[2, 1:] = -30.0                  # tiny length scales / amplitude: still factorises or reports -inf, never NaN
    monkeypatch.setenv("APGP_LOGLIK_PATH", "group")
    ll = gp.log_likelihood_batch(P, y)
    assert not np.any(np.isnan(ll))
    monkeypatch.setenv("APGP_LOGLIK_PATH", "tiled")
    ll_t = gp.log_likelihood_batch(P, y)
    fin = np.isfinite(ll_t)
    assert np.array_equal(fin, np.isfinite(ll))
    np.testing.assert_allclose(ll[fin], ll_t[fin], rtol=1e-10)
    for r in range(0, R, max(1, R // 6)):
        if not np.all(np.isfinite(P[r])):
            assert ll[r] == -np.inf
            continue
        orc.set_parameter_vector(P[r])
        ref = orc.log_likelihood(y, quiet=True)
        if np.isfinite(ref):
            assert abs(ll[r] - ref) <= 1e-9 * abs(ref), (r, ll[r], ref)
        else:
            assert ll[r] == -np.inf


@pytest.mark.parametrize("N,d", [(256, 2), (700, 5), (1024, 3)])
def test_loglik_group_is_independent_of_the_cluster_size(N, d, monkeypatch):
    """Tile ownership changes with the cluster size, the arithmetic per tile does not: identical bits for
    C = 1, 2, 4, 8, 16 (and for repeated calls: the flags are re-armed per launch)."""
    gp, orc, X, y = _pair(N, d, seed=3)
    rng = np.random.default_rng(1)
    P = np.column_stack([np.full(6, np.median(y))] + [0.5 * rng.standard_normal(6) for _ in range(len(gp) - 1)])
    monkeypatch.setenv("APGP_LOGLIK_PATH", "group")
    outs = []
    for C in (1, 2, 4, 8, 16, 2):
        monkeypatch.setenv("APGP_CHOL_CLUSTER", str(C))
        outs.append(gp.log_likelihood_batch(P, y))
    for o in outs[1:]:
        assert np.array_equal(o, outs[0]), (o, outs[0])


def test_not_positive_definite_in_the_group_path(monkeypatch):
    from approxposterior_b200 import GP, kernels
    rng = np.random.default_rng(0)
    X = rng.uniform(-1, 1, size=(300, 2))
    X[150] = X[7]                                               # duplicate point
    y = rng.standard_normal(300)
    gp = GP(kernel=kernels.ExpSquaredKernel([1.0, 1.0], ndim=2), fit_mean=True, mean=0.0, white_noise=-80.0)
    gp._x, gp._y = X, y
    gp._upload_training()
    monkeypatch.setenv("APGP_LOGLIK_PATH", "group")
    ll = gp.log_likelihood_batch(np.array([[0.0, 0.0, 0.0], [0.0, 3.0, 3.0]]), y)
    assert np.all(ll == -np.inf)


@pytest.mark.parametrize("N,method,options", [(256, "powell", {"maxiter": 2}), (300, "nelder-mead", {"maxfev": 80}),
                                              (520, "powell", {"maxfev": 40})])
def test_device_minimize_nll_group_is_the_scipy_iteration(N, method, options):
    """apgp_minimize_nll beyond one CTA's shared memory: one cluster per restart, optimiser state replicated in every
    CTA.  Same iterates as the host restatement of SciPy driven by the same objective => same optimum, bit for bit."""
    from approxposterior_b200 import _optimizers as opt
    gp, orc, X, y = _pair(N, 2, amp=3.0)
    assert gp.can_minimize_nll()
    np.random.seed(5)
    P = len(gp)
    x0s = np.array([[np.median(y)] + [np.random.randn() for _ in range(P - 1)] for _ in range(3)])
    _, f0, _ = gp.minimize_nll(x0s, y, evaluate_only=True)
    for r in range(3):
        orc.set_parameter_vector(x0s[r])
        assert abs(f0[r] + orc.log_likelihood(y, quiet=True)) <= 1e-9 * abs(f0[r])
    pd_, fd, nd = gp.minimize_nll(x0s, y, method=method, options=options)
    o = dict(options or {})
    make = (lambda t0: opt.powell_gen(t0, **o)) if method == "powell" else (lambda t0: opt.nelder_mead_gen(t0, _stable=True, **o))
    out, rounds, evals = opt.run_generators([make(t0) for t0 in x0s],
                                            lambda Q: gp.minimize_nll(np.array(Q), y, evaluate_only=True)[1])
    ph, fh = np.array([v[0] for v in out]), np.array([v[1] for v in out])
    assert np.array_equal(pd_, ph), (pd_, ph)
    assert np.array_equal(fd, fh)
    assert int(np.sum(nd)) == evals
    assert np.all(fd <= f0)


def test_optimize_gp_uses_the_device_engine_beyond_220_points():
    """gpUtils.optimizeGP (reference gpUtils.py:184-257) at N = 256: one launch for all restarts, same optimum as the
    host lock-step engine to optimiser tolerance."""
    from approxposterior_b200 import gpUtils
    res = {}
    for engine in ("device", "lockstep"):
        gp, orc, X, y = _pair(256, 2)
        np.random.seed(3)
        gp = gpUtils.optimizeGP(gp, X, y, nGPRestarts=4, method="powell", options={"maxiter": 3}, engine=engine)
        res[engine] = (gp.get_parameter_vector(), gp.log_likelihood(y), dict(gpUtils.optimizeGP.last_stats))
    assert res["device"][2]["scheduler"] == "device" and res["device"][2]["batches"] == 1
    assert abs(res["device"][1] - res["lockstep"][1]) <= 1e-5 * abs(res["lockstep"][1])
