#!/usr/bin/env python
"""Per-round cost of the tiled batched log-likelihood (N > 220: one launch sequence per optimiser round), and the
launch list of one round under APGP_TRACE via ncu is left to the caller.  JSON lines."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
for N in (256, 512, 1024, 2048):
    d = 2
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, d)); y = -0.5 * np.sum((X / 2) ** 2, axis=1)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, 0.5), ndim=d), fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    for R in (1, 3, 8, 64):
        P = gp.get_parameter_vector()[None, :] + 0.05 * rng.standard_normal((R, len(gp)))
        gp.log_likelihood_batch(P, y)
        n = 20 if N <= 1024 else 8
        t0 = time.perf_counter()
        for _ in range(n):
            ll = gp.log_likelihood_batch(P, y)
        dt = (time.perf_counter() - t0) / n
        print(json.dumps(dict(what="tiled loglik_batch round", N=N, d=d, R=R, ms_per_round=dt * 1e3,
                              launches_per_round=3 * ((N + 63) // 64) + 1, finite=int(np.isfinite(ll).sum()))), flush=True)
