#!/usr/bin/env python
"""cfg4-shaped batched log-likelihood (R restarts, d = 10, P = 12): fused cluster-per-restart kernel vs the multi-launch
tiled path, as evals/s and as a fraction of the FP64 peak with SURVEY 8(d)'s algorithmic flop count
F_ll = N^3/3 + 2 N^2 + (3 d + 1) N^2 / 2; plus optimizeGP wall time on both engines.  One JSON object per line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from approxposterior_b200 import GP, gpUtils, kernels  # noqa: E402

DEV = torch.device("cuda", 0)


def dgemm_peak():
    a = torch.randn(8192, 8192, dtype=torch.float64, device=DEV); b = torch.randn_like(a)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * 8192 ** 3 / best * 1e-9


def branin(u, v):
    return (v - 5.1 / (4 * np.pi ** 2) * u ** 2 + 5 / np.pi * u - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(u) + 10


PEAK = dgemm_peak()
rng = np.random.default_rng(64)
d = 10
which = sys.argv[1:] or ["256", "512", "1024", "2048"]
for N in [int(v) for v in which]:
    X = rng.uniform(-5, 5, size=(N, d))
    U = (X + 5) / 10
    y = -sum(branin(15 * U[:, 2 * i] - 5, 15 * U[:, 2 * i + 1]) for i in range(5)) / 100.0
    gp = GP(kernel=float(np.var(y)) * kernels.ExpSquaredKernel(np.ones(d), ndim=d), fit_mean=True,
            mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    F = N ** 3 / 3.0 + 2.0 * N * N + (3 * d + 1) * N * N / 2.0
    for R in (1, 8, 64, 148, 296):
        if N >= 2048 and R > 64:
            continue
        P = np.column_stack([np.full(R, np.median(y)), rng.standard_normal((R, 11))])
        for path, clusters in (("tiled", [None]), ("group", [None, 1, 2, 4, 8, 16])):
            os.environ["APGP_LOGLIK_PATH"] = path
            for C in clusters:
                if C is None:
                    os.environ.pop("APGP_CHOL_CLUSTER", None)
                else:
                    if R * C > 148 * 4 or (R >= 64 and C > 4):
                        continue
                    os.environ["APGP_CHOL_CLUSTER"] = str(C)
                gp.log_likelihood_batch(P, y)
                reps = 10 if N <= 512 else 4
                t0 = time.perf_counter()
                for _ in range(reps):
                    ll = gp.log_likelihood_batch(P, y)
                dt = (time.perf_counter() - t0) / reps
                print(json.dumps(dict(config="cfg4-loglik", N=N, d=d, R=R, path=path, cluster=C, ms_per_batch=dt * 1e3,
                                      nll_evals_per_s=R / dt, tflops_algorithmic=R * F / dt * 1e-12,
                                      frac_of_dgemm=R * F / dt * 1e-12 / PEAK, dgemm_tflops=PEAK,
                                      finite=int(np.isfinite(ll).sum()))), flush=True)
    os.environ.pop("APGP_CHOL_CLUSTER", None)
    os.environ.pop("APGP_LOGLIK_PATH", None)
    if N <= 1024:
        for engine in ("device", "lockstep"):
            np.random.seed(64)
            keep = gp.get_parameter_vector()
            t0 = time.perf_counter()
            gpUtils.optimizeGP(gp, X, y, nGPRestarts=64, method="powell", options={"maxiter": 3}, engine=engine)
            print(json.dumps(dict(config="cfg4-optGP", N=N, restarts=64, engine=engine, seconds=time.perf_counter() - t0,
                                  stats=gpUtils.optimizeGP.last_stats, best_ll=float(gp.log_likelihood(y)),
                                  note="Powell capped at 3 outer iterations per restart")), flush=True)
            gp.set_parameter_vector(keep); gp.recompute()
