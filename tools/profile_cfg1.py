#!/usr/bin/env python
"""Where the README configuration's BAPE iteration time goes: wall-clock split of one ApproxPosterior.run
(m0=50, m=20, nmax=2, 20 walkers x 2e4 steps) into minimizeObjective / optimizeGP / append_point / MCMC,
with the number of optimiser rounds each consumed.  One JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approxposterior_b200 import approx, gpUtils, utility as ut, likelihood as lh  # noqa: E402
from approxposterior_b200.gp import GP  # noqa: E402

acc = {}


def wrap(mod, name, key, stats_of=None):
    f = getattr(mod, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            e = acc.setdefault(key, dict(s=0.0, calls=0, rounds=0, evals=0))
            e["s"] += time.perf_counter() - t0
            e["calls"] += 1
            st = getattr(f, "last_stats", None)
            if st:
                e["rounds"] += st.get("batches", 0)
                e["evals"] += st.get("evals", 0)
    g.__dict__.update(f.__dict__)
    setattr(mod, name, g)
    return f


wrap(ut, "minimizeObjective", "minimizeObjective")
wrap(gpUtils, "optimizeGP", "optimizeGP")
wrap(GP, "append_point", "append_point")

for rep in range(2):
    acc.clear()
    np.random.seed(57)
    bounds = [(-5, 5), (-5, 5)]
    theta = lh.rosenbrockSample(50)
    yy = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
    gp = gpUtils.defaultGP(theta, yy, white_noise=-12)
    ap = approx.ApproxPosterior(theta=theta, y=yy, gp=gp, lnprior=lh.BoxPrior(bounds), lnlike=lh.rosenbrockLnlike,
                                priorSample=lh.rosenbrockSample, bounds=bounds, algorithm="bape")
    t0 = time.perf_counter()
    ap.run(m=20, nmax=2, estBurnin=True, nGPRestarts=3, mcmcKwargs={"iterations": int(2.0e4)},
           samplerKwargs={"nwalkers": 20}, cache=False, verbose=False, thinChains=False, onlyLastMCMC=True,
           timing=True, seed=57)
    total = time.perf_counter() - t0
    s = ap.sampler.get_chain(discard=ap.iburns[-1], flat=True)
    print(json.dumps(dict(config="cfg1-breakdown", rep=rep, total_s=total, trainingTime=ap.trainingTime,
                          mcmcTime=ap.mcmcTime, parts=acc, posterior_mean=s.mean(axis=0).tolist(),
                          hyper=ap.gp.get_parameter_vector().tolist())), flush=True)
