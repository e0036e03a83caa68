"""N2 on the device: integrated autocorrelation time / burn-in of a device-resident chain (csrc/autocorr.cu,
apgp_integrated_time) against emcee's estimator as restated in oracle/refshim.py (per-walker FFT), and the device-resident
sampler output it works on.  Reference: mcmcUtils.estimateBurnin (mcmcUtils.py:164-227), approx.py:479-481, 853."""
import numpy as np
import pytest

from conftest import rosenbrock_training

pytestmark = pytest.mark.gpu


def _gp():
    from approxposterior_b200 import GP, kernels
    theta, _ = rosenbrock_training(50)
    y = -0.5 * ((theta[:, 0] - 1.0) ** 2 / 1.5 ** 2 + (theta[:, 1] + 0.5) ** 2 / 0.8 ** 2)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp([0.5, 1.2]), ndim=2), fit_mean=True, mean=float(np.median(y)),
            white_noise=-12.0)
    gp.compute(theta, y=y)
    return gp, y


def _ar1(n, W, d, rho, seed):
    rng = np.random.default_rng(seed)
    x = np.empty((n, W, d))
    x[0] = rng.standard_normal((W, d))
    e = rng.standard_normal((n, W, d)) * np.sqrt(1 - np.asarray(rho) ** 2)
    for t in range(1, n):
        x[t] = np.asarray(rho) * x[t - 1] + e[t]
    return x


@pytest.mark.parametrize("n,W,d,rho,discard,thin", [
    (5000, 32, 2, (0.9, 0.97), 0, 1),          # few walkers: series split into s-chunks
    (3000, 20, 1, (0.95,), 100, 1),
    (3001, 24, 3, (0.5, 0.9, 0.99), 7, 3),     # window beyond the first 256 lags for the slow dimension; thinning
    (600, 6000, 2, (0.8, 0.9), 0, 1),          # many walkers: grouped CTAs
    (20000, 20, 2, (0.97, 0.98), 0, 1),        # README-sized series
    (300, 10, 2, (0.9999, 0.5), 0, 1),         # no window inside the chain for dimension 0: emcee takes the last lag
])
def test_integrated_time_matches_emcee_estimator(n, W, d, rho, discard, thin):
    import torch
    from oracle import refshim
    gp, _ = _gp()
    x = _ar1(n, W, d, rho, seed=n + W)
    ref = refshim.integrated_time(x[discard + thin - 1::thin], tol=0)
    for chain in (x, torch.from_numpy(x).cuda()):                      # host chain (uploaded) and device-resident chain
        tau, win = gp.integrated_time(chain, discard=discard, thin=thin)
        np.testing.assert_allclose(tau, ref, rtol=1e-9)
    # and the product's own host estimator agrees (FFT with the normalisation folded into the spectra)
    from approxposterior_b200.sampler import integrated_time
    np.testing.assert_allclose(integrated_time(x[discard + thin - 1::thin], tol=0), ref, rtol=1e-9)


def test_device_sampler_keeps_the_chain_on_the_gpu_and_estimates_burnin_there():
    """engine="device": chain / log_prob / blobs are torch CUDA tensors, get_autocorr_time and estimateBurnin run on
    them in place, get_chain copies only the requested slice; results equal the host path on the same chain."""
    import torch
    from approxposterior_b200 import mcmcUtils
    from approxposterior_b200.sampler import EnsembleSampler, integrated_time
    gp, y = _gp()
    bounds = [(-5.0, 5.0)] * 2
    rng = np.random.default_rng(0)
    nens, nw = 64, 20
    s = EnsembleSampler(nw, 2, engine="device", gp=gp, y=y, bounds=bounds, nens=nens, seed=11)
    s.run_mcmc(rng.uniform(-5, 5, size=(nens * nw, 2)), 1500)
    assert isinstance(s._chain, torch.Tensor) and s._chain.is_cuda and tuple(s._chain.shape) == (1500, nens * nw, 2)
    tau_dev = s.get_autocorr_time(tol=0)
    host = s.get_chain()
    assert isinstance(host, np.ndarray) and host.shape == (1500, nens * nw, 2)
    tau_host = integrated_time(host, tol=0)
    np.testing.assert_allclose(tau_dev, tau_host, rtol=1e-9)
    iburn, ithin = mcmcUtils.estimateBurnin(s, estBurnin=True, thinChains=True)
    assert iburn == int(2.0 * np.max(tau_host)) and ithin == max(int(0.5 * np.min(tau_host)), 1)
    assert s.get_chain(discard=iburn, thin=ithin, flat=True).shape == (len(range(iburn + ithin - 1, 1500, ithin)) * nens * nw, 2)
    # same seed through the host-output path: identical chain
    ref = gp.run_ensembles(y, np.random.default_rng(0).uniform(-5, 5, size=(nens * nw, 2)), 1500, bounds, nens=nens, seed=11)
    assert np.array_equal(ref["chain"], host)
    assert np.array_equal(ref["naccepted"], s.naccepted)


def test_burnin_kat_through_the_device_estimator():
    """reference tests/test_Burnin.py:17-91 ([iburn, ithin] = [67, 15]): emcee-flow chain produced on the host, the
    autocorrelation estimate taken by the device kernel."""
    from approxposterior_b200.sampler import EnsembleSampler
    gp, _ = _gp()
    np.random.seed(42)
    N = 50
    x = np.sort(10 * np.random.rand(N))
    obs = -0.9594 * x + 4.294
    obs += 0.5 * np.random.randn(N)

    def log_prob(T):
        T = np.atleast_2d(T)
        m, b = T[:, 0], T[:, 1]
        ok = (m > -5.0) & (m < 0.5) & (b > 0.0) & (b < 10.0)
        ll = -0.5 * np.sum((obs[None, :] - (m[:, None] * x[None, :] + b[:, None])) ** 2 / 0.5 ** 2, axis=1)
        return np.where(ok, ll, -np.inf), np.zeros(len(m))

    p0 = np.random.randn(32, 2)
    sampler = EnsembleSampler(32, 2, log_prob, engine="host-rng")
    with np.errstate(invalid="ignore"):
        sampler.run_mcmc(p0, 5000)
    tau, win = gp.integrated_time(sampler.get_chain())
    iburn, ithin = int(2.0 * np.max(tau)), max(int(0.5 * np.min(tau)), 1)
    assert np.allclose([67, 15], [iburn, ithin], rtol=1.0e-1), (iburn, ithin)
    np.testing.assert_allclose(tau, sampler.get_autocorr_time(tol=0), rtol=1e-9)
