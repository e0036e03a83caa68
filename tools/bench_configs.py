#!/usr/bin/env python
"""Secondary measurements for BASELINE.json's other configs (not the driver's bench line):
cfg1 README BAPE iteration time, cfg2 device sampler throughput, cfg4 batched optGP log-likelihood,
cfg5 predict sweep, plus the mean-only predict.  One JSON object per line; results are committed under
profiles/.  CPU numbers come from the oracle (restated george/emcee) on this box's host cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from approxposterior_b200 import GP, approx, gpUtils, kernels, likelihood as lh  # noqa: E402
from oracle import GPOracle  # noqa: E402

DEV = torch.device("cuda", 0)
which = set(sys.argv[1:]) or {"cfg1", "cfg2", "cfg4", "cfg5", "mean"}


def emit(**kw):
    print(json.dumps(kw), flush=True)


def dgemm_peak():
    a = torch.randn(8192, 8192, dtype=torch.float64, device=DEV); b = torch.randn_like(a)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * 8192 ** 3 / best * 1e-9


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def make_gp(N, d, seed=0, amp=None):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    k = kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d)
    if amp is not None:
        k = amp * k
    gp = GP(kernel=k, fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    return gp, X, y


if "cfg5" in which or "mean" in which:
    PEAK = dgemm_peak()

if "cfg5" in which:
    Q = 1 << 20
    for N in (256, 512, 1024, 2048, 4096, 8192):
        for d in (2, 5, 10, 20):
            gp, X, y = make_gp(N, d, seed=N + d)
            q = -5 + 10 * torch.rand((Q, d), dtype=torch.float64, device=DEV)
            ms = timed(lambda: gp._predict_raw(q, True, utility="bape"), reps=2 if N >= 4096 else 3)
            fl = (float(N) * N + (3 * d + 6) * N) * Q
            emit(config="cfg5", kernel="predict_var", N=N, d=d, Q=Q, ms=ms, evals_per_s=Q / ms * 1e3,
                 tflops_algorithmic=fl / ms * 1e-9, frac_of_dgemm=fl / ms * 1e-9 / PEAK, dgemm_tflops=PEAK)
            del gp, q
            torch.cuda.empty_cache()

if "mean" in which:
    Q = 1 << 22
    for N, d in ((1024, 2), (2048, 5), (8192, 20)):
        gp, X, y = make_gp(N, d, seed=1)
        q = -5 + 10 * torch.rand((Q, d), dtype=torch.float64, device=DEV)
        ms = timed(lambda: gp._predict_raw(q, False))
        ops = (2 * d + 12) * N * Q        # FP64 instructions actually needed per (query, point): 2d + table exp 10 + 1 + 1
        emit(config="mean-only", kernel="predict_mean", N=N, d=d, Q=Q, ms=ms, evals_per_s=Q / ms * 1e3,
             fp64_pipe_frac=ops / ms * 1e-9 / (PEAK / 2), note="exp-bound: fraction of the DFMA instruction peak (DGEMM/2 lane-ops)")

if "cfg2" in which:
    # 65536 walkers as 2048 independent ensembles x 32 walkers, N=1024 d=2, 1000 steps (thin 20 on output)
    np.random.seed(57)
    theta = lh.rosenbrockSample(1024)
    y = np.array([lh.rosenbrockLnlike(t) for t in theta])
    gp = gpUtils.defaultGP(theta, y, fitAmp=False)
    gp.set_parameter_vector([float(np.median(y)), 0.5, 1.2]); gp.recompute()
    bounds = [(-5, 5), (-5, 5)]
    for nens, nw, nsteps in ((2048, 32, 1000), (1, 20, 20000)):
        p0 = np.random.uniform(-5, 5, size=(nens * nw, 2))
        gp.run_ensembles(y, p0, 10, bounds, nens=nens, seed=1)
        t0 = time.perf_counter()
        out = gp.run_ensembles(y, p0, nsteps, bounds, nens=nens, seed=2, thin=1000 if nens > 1 else 1)
        dt = time.perf_counter() - t0
        emit(config="cfg2" if nens > 1 else "cfg1-mcmc", kernel="sampler", N=1024, d=2, nens=nens, nwalkers=nw,
             nsteps=nsteps, seconds=dt, lnprob_evals_per_s=nens * nw * nsteps / dt,
             acceptance=float(out["naccepted"].mean() / nsteps), note="wall time incl. D2H; the 65536-walker run keeps only the final state (thin=nsteps), the README-shaped "
                  "single ensemble returns its full chain")
    # CPU oracle: reference-shaped per-call loop (approx.py:178) and batched
    orc = GPOracle(2, np.exp([0.5, 1.2]), mean=float(np.median(y)), white_noise=-12.0); orc.compute(theta)
    q = np.random.uniform(-5, 5, size=(2000, 2))
    t0 = time.perf_counter()
    for i in range(300):
        orc.predict(y, q[i:i + 1], return_var=False)
    per_call = 300 / (time.perf_counter() - t0)
    t0 = time.perf_counter(); orc.predict(y, np.random.uniform(-5, 5, size=(200000, 2)), return_var=False)
    batched = 200000 / (time.perf_counter() - t0)
    emit(config="cfg2-cpu", N=1024, d=2, per_call_evals_per_s=per_call, batched_evals_per_s=batched, cores=os.cpu_count())

if "cfg4" in which:
    # d=10 Branin-style synthetic, fitAmp -> P=12, 64 restarts batched; N = 64..256
    def branin(u, v):
        return (v - 5.1 / (4 * np.pi ** 2) * u ** 2 + 5 / np.pi * u - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(u) + 10
    rng = np.random.default_rng(64)
    for N in (64, 128, 200, 256, 512):
        X = rng.uniform(-5, 5, size=(N, 10))
        U = (X + 5) / 10
        y = -sum(branin(15 * U[:, 2 * i] - 5, 15 * U[:, 2 * i + 1]) for i in range(5)) / 100.0
        gp = GP(kernel=float(np.var(y)) * kernels.ExpSquaredKernel(np.ones(10), ndim=10), fit_mean=True,
                mean=float(np.median(y)), white_noise=-12.0)
        gp.compute(X, y=y)
        P = np.column_stack([np.full(64, np.median(y)), rng.standard_normal((64, 11))])
        gp.log_likelihood_batch(P, y)
        t0 = time.perf_counter()
        for _ in range(20):
            ll = gp.log_likelihood_batch(P, y)
        dt = (time.perf_counter() - t0) / 20
        orc = GPOracle(10, np.ones(10), amp=float(np.var(y)), mean=float(np.median(y)), white_noise=-12.0); orc.compute(X)
        t0 = time.perf_counter()
        for p in P[:16]:
            orc.set_parameter_vector(p); orc.log_likelihood(y, quiet=True)
        cpu = (time.perf_counter() - t0) / 16
        emit(config="cfg4", kernel="loglik_batch", N=N, d=10, restarts=64, ms_per_batch=dt * 1e3,
             nll_evals_per_s=64 / dt, cpu_oracle_ms_per_eval=cpu * 1e3, cpu_nll_evals_per_s=1 / cpu,
             finite=int(np.isfinite(ll).sum()))
        for engine in ("device", "lockstep"):
            if engine == "device" and not gp.can_minimize_nll():
                continue
            if engine == "lockstep" and N not in (64, 200, 256):
                continue
            np.random.seed(64)
            keep = gp.get_parameter_vector()
            t0 = time.perf_counter()
            gpUtils.optimizeGP(gp, X, y, nGPRestarts=64, method="powell", options={"maxiter": 3}, engine=engine)
            emit(config="cfg4-optGP", N=N, restarts=64, engine=engine, seconds=time.perf_counter() - t0,
                 stats=gpUtils.optimizeGP.last_stats, best_ll=float(gp.log_likelihood(y)),
                 note="Powell capped at 3 outer iterations per restart")
            gp.set_parameter_vector(keep); gp.recompute()

if "cfg1" in which:
    # README configuration: m0=50, m=20, nmax=2, 20 walkers x 2e4 steps, nGPRestarts=3
    from oracle import UTILITY_BY_NAME, default_gp_oracle

    from oracle import refshim

    class OracleGP(refshim.GP):
        """CPU oracle behind the same drivers (george constructor signature + lock-step entry points)."""
        def predict_utility(self, y, t, kind, bounds=None, zeta=0.01):
            mu, var = self.predict(y, t, return_var=True)
            fn = UTILITY_BY_NAME[kind]
            return mu, var, (fn(mu, var, y.max(), zeta, True) if kind == "jones" else fn(mu, var, True))

        def log_likelihood_batch(self, P, y):
            p0 = self.get_parameter_vector(); out = []
            for p in np.atleast_2d(P):
                self.set_parameter_vector(p); out.append(self.log_likelihood(y, quiet=True))
            self.set_parameter_vector(p0); self.recompute(quiet=True)
            return np.array(out)

    def readme_problem():
        np.random.seed(57)
        theta = lh.rosenbrockSample(50)
        y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
        return theta, y

    bounds = [(-5, 5), (-5, 5)]
    from approxposterior_b200 import utility as ut
    for label, engine, scan in (("device optimisers (one CTA per restart)", "device", None),
                                ("host lock-step restarts", "device", None),
                                ("host lock-step restarts", "host-rng", None),
                                ("device scan 65536 + device polish", "device", 65536)):
        ut.DEVICE_OPTIMIZER = gpUtils.DEVICE_OPTIMIZER = label.startswith("device optimisers") or scan is not None
        theta, y = readme_problem()
        gp = gpUtils.defaultGP(theta, y, white_noise=-12)
        prior = lh.BoxPrior(bounds) if engine == "device" else lh.rosenbrockLnprior
        ap = approx.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=prior, lnlike=lh.rosenbrockLnlike,
                                    priorSample=lh.rosenbrockSample, bounds=bounds, algorithm="bape")
        t0 = time.perf_counter()
        ap.run(m=20, nmax=2, estBurnin=True, nGPRestarts=3, mcmcKwargs={"iterations": int(2.0e4)},
               samplerKwargs={"nwalkers": 20}, cache=False, verbose=False, thinChains=False, onlyLastMCMC=True,
               timing=True, seed=57, scanCandidates=scan)
        tot = time.perf_counter() - t0
        s = ap.sampler.get_chain(discard=ap.iburns[-1], flat=True, thin=ap.ithins[-1])
        emit(config="cfg1", acquisition=label, engine=engine, total_s=tot, trainingTime=ap.trainingTime,
             mcmcTime=ap.mcmcTime, bape_iteration_s=float(np.mean(ap.trainingTime)),
             posterior_mean=s.mean(axis=0).tolist(), posterior_std=s.std(axis=0).tolist(), iburn=int(ap.iburns[-1]),
             note="README.md:84-121 configuration")
    ut.DEVICE_OPTIMIZER = gpUtils.DEVICE_OPTIMIZER = True
    # the same drivers on the CPU oracle (one BAPE iteration + a 2000-step per-half-step-batched MCMC)
    theta, y = readme_problem()
    g0 = default_gp_oracle(theta, y)
    gp = OracleGP(kernel=refshim._ExpSquared(np.exp(g0.log_M), 2), fit_mean=True, mean=g0.mean, white_noise=-12); gp.compute(theta)
    ap = approx.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=lh.rosenbrockLnprior, lnlike=lh.rosenbrockLnlike,
                                priorSample=lh.rosenbrockSample, bounds=bounds, algorithm="bape")
    ap.run(m=20, nmax=1, estBurnin=True, nGPRestarts=3, mcmcKwargs={"iterations": 2000},
           samplerKwargs={"nwalkers": 20, "engine": "host-rng"}, cache=False, verbose=False, thinChains=False,
           onlyLastMCMC=True, timing=True, seed=57)
    # reference-shaped MCMC cost: one predict per walker per step (approx.py:178), extrapolated to 4e5 calls
    t0 = time.perf_counter()
    for i in range(2000):
        ap.gp.predict(ap.y, ap.theta[i % 50:i % 50 + 1], return_var=False)
    per_call = (time.perf_counter() - t0) / 2000
    emit(config="cfg1-cpu-oracle", bape_iteration_s=ap.trainingTime[0], mcmc_2000_steps_batched_s=ap.mcmcTime[0],
         mcmc_per_call_us=per_call * 1e6, mcmc_4e5_calls_extrapolated_s=per_call * 4e5, cores=os.cpu_count(),
         note="restated george/emcee oracle through the same host drivers; george itself is not installable")
