#!/usr/bin/env python
"""Device sampler with a training set too large for one CTA's shared memory: stream it from L2 (round 1,
APGP_SAMPLER_NO_SPLIT=1) vs partition it over a cluster of 2/4/8 CTAs (round 2).  us per sampler step."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
for N, d, nw, nens, nsteps in ((2000, 10, 200, 1, 2000), (2000, 10, 200, 16, 500), (4096, 5, 100, 1, 2000), (8192, 20, 80, 4, 300)):
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, d)); y = -0.5 * np.sum((X / 2) ** 2, axis=1)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, 2.0), ndim=d), fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    p0 = rng.uniform(-2, 2, size=(nens * nw, d))
    bounds = [(-5, 5)] * d
    out = dict(N=N, d=d, nwalkers=nw, nens=nens, nsteps=nsteps)
    for label, env in (("stream_l2", {"APGP_SAMPLER_NO_SPLIT": "1"}), ("auto", {}), ("split2", {"APGP_SAMPLER_SPLIT": "2"}),
                       ("split4", {"APGP_SAMPLER_SPLIT": "4"}), ("split8", {"APGP_SAMPLER_SPLIT": "8"})):
        for k in ("APGP_SAMPLER_NO_SPLIT", "APGP_SAMPLER_SPLIT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        gp.run_ensembles(y, p0, 20, bounds, nens=nens, seed=1, thin=20)
        t0 = time.perf_counter()
        r = gp.run_ensembles(y, p0, nsteps, bounds, nens=nens, seed=2, thin=nsteps)
        dt = time.perf_counter() - t0
        out["us_per_step_" + label] = round(dt / nsteps * 1e6, 2)
        out["evals_per_s_" + label] = round(nens * nw * nsteps / dt)
    print(json.dumps(out), flush=True)
