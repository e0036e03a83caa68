#!/usr/bin/env python
"""One ensemble of 20*ndim walkers (emcee's default through mcmcUtils.py:75) on the device sampler: seconds per run
for several (N, d): one CTA of 8 warps (the first version), one CTA of up to 32 warps, thread-block clusters of 2/4/8
CTAs sharing the ensemble through distributed shared memory, and the automatic choice.  Chains must be identical."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
for N, d, nsteps in ((90, 2, 20000), (500, 5, 5000), (1000, 5, 5000), (2000, 10, 2000)):
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, d)); y = -0.5 * np.sum((X / 2) ** 2, axis=1)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, 2.0), ndim=d), fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    nw = 20 * d
    p0 = rng.uniform(-2, 2, size=(nw, d))
    bounds = [(-5, 5)] * d
    out = dict(N=N, d=d, nwalkers=nw, nsteps=nsteps)
    ref = None
    for label, warps, clus in (("8warps_1cta", "8", "1"), ("auto_warps_1cta", None, "1"), ("cluster2", None, "2"),
                               ("cluster4", None, "4"), ("cluster8", None, "8"), ("auto", None, None)):
        for k, v in (("APGP_SAMPLER_WARPS", warps), ("APGP_SAMPLER_CLUSTER", clus)):
            if v: os.environ[k] = v
            else: os.environ.pop(k, None)
        gp.run_ensembles(y, p0, 50, bounds, nens=1, seed=1)
        t0 = time.perf_counter()
        r = gp.run_ensembles(y, p0, nsteps, bounds, nens=1, seed=2)
        out["s_" + label] = round(time.perf_counter() - t0, 4)
        if ref is None:
            ref = r
        out["same_chain_" + label] = bool(np.array_equal(ref["chain"], r["chain"]) and np.array_equal(ref["log_prob"], r["log_prob"])
                                          and np.array_equal(ref["naccepted"], r["naccepted"]))
    out["speedup"] = out["s_8warps_1cta"] / out["s_auto"]
    print(json.dumps(out), flush=True)
