"""CUDA path vs the committed golden vectors (tests/golden/*.npz, made by make_golden.py from the
KAT-pinned oracle), plus size-independent properties at BASELINE's full size (N=2048, d=5, 2^20 queries)."""
import glob
import os

import numpy as np
import pytest

from conftest import utility_error_bound

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _gp_from(g):
    from approxposterior_b200 import GP, kernels
    d = g["X"].shape[1]
    amp = None if np.isnan(g["amp"]) else float(g["amp"])
    k = kernels.ExpSquaredKernel(np.exp(g["logM"]), ndim=d)
    if amp is not None:
        k = amp * k
    gp = GP(kernel=k, fit_mean=True, mean=float(g["mean"]), white_noise=-12.0)
    gp.compute(g["X"], y=g["y"])
    return gp, (1.0 if amp is None else amp)


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_against_golden(path):
    g = np.load(path)
    gp, A = _gp_from(g)
    y, Xq = g["y"], g["Xq"]
    d = Xq.shape[1]
    bounds = [(-5.0, 5.0)] * d
    scale = max(1.0, np.max(np.abs(y)))
    for kind in ("agp", "bape", "jones"):
        mu, var, u = gp.predict_utility(y, Xq, kind, bounds=bounds)
        np.testing.assert_allclose(mu, g["mu"], rtol=1e-9, atol=1e-9 * scale)
        assert np.all(np.abs(var - g["var"]) <= 1e-9 * A + 1e-9 * np.abs(g["var"]))
        ref = g[kind]
        assert np.array_equal(np.isposinf(u), np.isposinf(ref))
        # end-to-end utility parity at the bound the 1e-9 (mu, var) parity implies (conftest.utility_error_bound)
        bound, resolved = utility_error_bound(kind, g["mu"], g["var"], A, scale, ybest=float(np.max(y)))
        good = np.isfinite(ref) & resolved
        assert np.sum(good) > 0.5 * np.sum(np.isfinite(ref))
        assert np.all(np.abs(u[good] - ref[good]) <= bound[good] + 1e-9 * np.abs(ref[good])), \
            (kind, np.max(np.abs(u[good] - ref[good]) / (bound[good] + 1e-9 * np.abs(ref[good]))))
    ll = gp.log_likelihood_batch(g["P"], y)
    np.testing.assert_allclose(ll, g["ll"], rtol=1e-9)
    gp.set_parameter_vector(g["P"][2])
    np.testing.assert_allclose(gp.grad_log_likelihood(y), g["grad"], rtol=1e-7,
                               atol=1e-8 * np.max(np.abs(g["grad"])))


def test_empty_and_ragged_queries():
    g = np.load(GOLDEN[0])
    gp, A = _gp_from(g)
    y, Xq = g["y"], g["Xq"]
    mu, var = gp.predict(y, np.empty((0, Xq.shape[1])), return_cov=False, return_var=True)
    assert mu.shape == (0,) and var.shape == (0,)
    for Q in (1, 2, 63, 64, 65, 127, 129, 257):
        mu, var = gp.predict(y, Xq[:Q], return_cov=False, return_var=True)
        np.testing.assert_allclose(mu, g["mu"][:Q], rtol=1e-9, atol=1e-9 * max(1.0, np.max(np.abs(y))))
        assert np.all(np.abs(var - g["var"][:Q]) <= 2e-9 * A)
    with pytest.raises(ValueError):
        gp.predict(y, np.zeros((3, Xq.shape[1] + 1)), return_cov=False)
    q = Xq[:8].copy()
    q[3, 0] = np.nan
    mu = gp.predict(y, q, return_cov=False, return_var=False)
    assert np.isnan(mu[3]) and np.all(np.isfinite(np.delete(mu, 3)))


def test_full_size_properties():
    """N=2048, d=5, 2^20 queries (BASELINE configs[2]): properties that do not need the oracle at scale,
    plus an oracle check on a 1500-query subsample."""
    import torch
    from bench import BOUNDS, make_problem
    from approxposterior_b200 import GP, kernels
    from oracle import GPOracle, bape_utility
    X, y, logM, mean = make_problem()
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=5), fit_mean=True, mean=mean, white_noise=-12.0)
    gp.compute(X, y=y)
    dev = torch.device("cuda", 0)
    Q = 1 << 20
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    cand = -5.0 + 10.0 * torch.rand((Q, 5), dtype=torch.float64, device=dev, generator=gen)
    mu, var, u = gp._predict_raw(cand, True, utility="bape", bounds=BOUNDS, ybest=float(y.max()))
    torch.cuda.synchronize()
    assert bool(torch.all(torch.isfinite(mu))) and bool(torch.all(var > -1e-9)) and bool(torch.all(var < 1 + 1e-9))
    # (1) permutation invariance: every query's result is independent of its position in the batch
    perm = torch.randperm(Q, device=dev, generator=gen)
    mu2, var2, u2 = gp._predict_raw(cand[perm].contiguous(), True, utility="bape", bounds=BOUNDS, ybest=float(y.max()))
    assert torch.equal(mu2, mu[perm]) and torch.equal(var2, var[perm])
    assert torch.equal(torch.nan_to_num(u2, nan=7.0), torch.nan_to_num(u[perm], nan=7.0))
    # (2) interpolation: at the training inputs the posterior mean returns y and the variance collapses
    mu_t, var_t = gp.predict(y, X, return_cov=False, return_var=True)
    assert np.max(np.abs(mu_t - y)) < 1e-3 and np.max(np.abs(var_t)) < 1e-4
    # (3) mean-only kernel agrees with the mean of the fused kernel
    mu_only = gp._predict_raw(cand[:100000].contiguous(), False)[0]
    assert float(torch.max(torch.abs(mu_only - mu[:100000]))) < 1e-9 * max(1.0, float(np.max(np.abs(y))))
    # (4) oracle on a subsample
    idx = torch.randint(0, Q, (1500,), device=dev, generator=gen)
    orc = GPOracle(5, np.exp(logM), mean=mean, white_noise=-12.0)
    orc.compute(X)
    mu_o, var_o = orc.predict(y, cand[idx].cpu().numpy(), return_var=True)
    np.testing.assert_allclose(mu[idx].cpu().numpy(), mu_o, rtol=1e-9, atol=1e-9 * np.max(np.abs(y)))
    assert np.all(np.abs(var[idx].cpu().numpy() - var_o) <= 2e-9)
    u_o = bape_utility(mu[idx].cpu().numpy(), var[idx].cpu().numpy(), True)
    fin = np.isfinite(u_o)
    np.testing.assert_allclose(u[idx].cpu().numpy()[fin], u_o[fin], rtol=1e-9, atol=1e-9)


def test_sampler_many_ensembles_deterministic_and_in_bounds():
    g = np.load(os.path.join(HERE, "golden", "golden_n96_d2.npz"))
    gp, _ = _gp_from(g)
    y = g["y"]
    bounds = [(-5.0, 5.0)] * 2
    rng = np.random.default_rng(0)
    nens, nw = 64, 16
    p0 = rng.uniform(-5, 5, size=(nens * nw, 2))
    a = gp.run_ensembles(y, p0, 300, bounds, nens=nens, seed=9)
    b = gp.run_ensembles(y, p0, 300, bounds, nens=nens, seed=9)
    c = gp.run_ensembles(y, p0, 300, bounds, nens=nens, seed=10)
    assert np.array_equal(a["chain"], b["chain"]) and not np.array_equal(a["chain"], c["chain"])
    assert a["chain"].shape == (300, nens * nw, 2)
    assert np.all(np.abs(a["chain"]) <= 5.0)
    # log_prob stored with the chain equals the surrogate mean at the stored positions
    last = a["chain"][-1]
    mu = gp.predict(y, last, return_cov=False, return_var=False)
    np.testing.assert_allclose(a["log_prob"][-1], mu, rtol=1e-9, atol=1e-9 * np.max(np.abs(y)))
    thin = gp.run_ensembles(y, p0, 300, bounds, nens=nens, seed=9, thin=10)
    assert np.array_equal(thin["chain"], a["chain"][9::10])
