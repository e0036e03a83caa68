#!/usr/bin/env python
"""Per-evaluation cost of the device-resident optimisers: wall time of one apgp_minimize_nll / apgp_minimize_utility
launch divided by the longest restart's evaluation count, next to the launch+copy floor (evaluate_only).  JSON lines."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approxposterior_b200 import GP, kernels  # noqa: E402


def best(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); out = fn(); ts.append(time.perf_counter() - t0)
    return min(ts), out


SPLIT_ONLY = "--split" in sys.argv
for N, d in (() if SPLIT_ONLY else ((50, 2), (70, 2), (90, 2), (150, 2), (200, 2), (64, 10), (128, 10), (200, 10))):
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, d))
    y = -0.5 * np.sum((X / 2.0) ** 2, axis=1) + 0.1 * rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, 4.0), ndim=d), fit_mean=True, mean=float(np.median(y)),
            white_noise=-12.0)
    gp.compute(X, y=y)
    np.random.seed(1)
    R = 3 if d == 2 else 64
    P0 = np.array([[np.median(y)] + list(np.random.randn(d)) for _ in range(R)])
    opts = None if d == 2 else {"maxiter": 2}
    t_floor, _ = best(lambda: gp.minimize_nll(P0, y, evaluate_only=True))
    t_run, (p, f, nfev) = best(lambda: gp.minimize_nll(P0, y, method="powell", options=opts), reps=3)
    print(json.dumps(dict(what="minimize_nll powell", N=N, d=d, restarts=R, floor_us=t_floor * 1e6, run_ms=t_run * 1e3,
                          nfev_max=int(nfev.max()), nfev_sum=int(nfev.sum()),
                          us_per_eval=(t_run - t_floor) / max(int(nfev.max()), 1) * 1e6,
                          cycles_per_eval=float(np.mean(gp.last_opt_stats[:, 2] / gp.last_opt_stats[:, 0])))), flush=True)
    bounds = [(-5.0, 5.0)] * d
    S = rng.uniform(-4, 4, size=(5 if d == 2 else 64, d))
    t_floor, _ = best(lambda: gp.minimize_utility(y, S, "bape", bounds=bounds, evaluate_only=True))
    t_run, (x, f, nfev) = best(lambda: gp.minimize_utility(y, S, "bape", bounds=bounds, options={"adaptive": True}), reps=3)
    print(json.dumps(dict(what="minimize_utility nelder-mead", N=N, d=d, starts=len(S), floor_us=t_floor * 1e6,
                          run_ms=t_run * 1e3, nfev_max=int(nfev.max()), nfev_sum=int(nfev.sum()),
                          us_per_eval=(t_run - t_floor) / max(int(nfev.max()), 1) * 1e6,
                          cycles_per_eval=float(np.mean(gp.last_opt_stats[:, 2] / gp.last_opt_stats[:, 0])))), flush=True)

if os.environ.get("APGP_LIB"):          # profiling build: where the cycles of one nll evaluation go
    import ctypes
    from approxposterior_b200 import _lib
    lib = _lib.load()
    for N in (50, 70, 90, 200):
        rng = np.random.default_rng(N)
        X = rng.uniform(-5, 5, size=(N, 2))
        y = -0.5 * np.sum((X / 2.0) ** 2, axis=1) + 0.1 * rng.standard_normal(N)
        gp = GP(kernel=kernels.ExpSquaredKernel(np.full(2, 4.0), ndim=2), fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
        gp.compute(X, y=y)
        P0 = gp.get_parameter_vector()[None, :]
        buf = (ctypes.c_longlong * 16)()
        for rep in range(3):                 # ONE evaluation, exactly (third repetition: warm instruction cache)
            lib.apgp_debug_read_prof(buf)
            gp.minimize_nll(P0, y, evaluate_only=True)
            lib.apgp_debug_read_prof(buf)
        v1 = np.array(list(buf), dtype=float)
        print(json.dumps(dict(what="ONE nll evaluation, cycles (thread 0)", N=N, total=int(gp.last_opt_stats[0, 2]),
                              prologue=v1[6], build_loop=v1[7], phase1=v1[0], phase2=v1[1], p2_load_D=v1[8],
                              p2_factor=v1[9], p2_rowsolve=v1[10], tail=v1[3])), flush=True)
        p, f, nfev = gp.minimize_nll(P0, y, method="powell")
        lib.apgp_debug_read_prof(buf)
        v = np.array(list(buf), dtype=float) / float(nfev[0])
        print(json.dumps(dict(what="nll eval cycle split (thread 0, per evaluation)", N=N, nfev=int(nfev[0]),
                              total=float(gp.last_opt_stats[0, 2] / nfev[0]), build=v[2], chol_phase1_incl_barrier=v[0],
                              chol_phase1_work=v[4], chol_phase2_incl_barrier=v[1], chol_phase2_work=v[5], tail=v[3], prologue=v[6], build_loop=v[7],
                              p2_load_D=v[8], p2_factor=v[9], p2_rowsolve=v[10])), flush=True)
