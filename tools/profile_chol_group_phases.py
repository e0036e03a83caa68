#!/usr/bin/env python
"""Per-phase SM-cycle breakdown of the fused cluster log-likelihood (chol_group.cuh) from the -DAPGP_PROF build.
Run as  APGP_LIB=approxposterior_b200/libapgp_prof.so APGP_PROF_PRINT=1 python tools/profile_chol_group_phases.py
The library prints one '[cgprof]' line per CTA (the first two CTAs of the grid) per call to stderr: cycles of
build / diag load / diag factor / diag store+publish / deferred updates / wait A / panel / publish+wait B /
urgent updates / final / total."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approxposterior_b200 import GP, kernels  # noqa: E402

os.environ["APGP_LOGLIK_PATH"] = "group"
rng = np.random.default_rng(64)
d = 10
for N in [int(v) for v in (sys.argv[1:] or ["256", "512", "1024"])]:
    X = rng.uniform(-5, 5, size=(N, d))
    y = np.sin(X).sum(axis=1)
    gp = GP(kernel=float(np.var(y)) * kernels.ExpSquaredKernel(np.ones(d), ndim=d), fit_mean=True,
            mean=float(np.median(y)), white_noise=-12.0)
    gp.compute(X, y=y)
    for R, Cs in ((1, (1, 8)), (64, (1, 2))):
        P = np.column_stack([np.full(R, np.median(y)), rng.standard_normal((R, 11))])
        for C in Cs:
            os.environ["APGP_CHOL_CLUSTER"] = str(C)
            print(f"# warm-up N={N} R={R} C={C}", file=sys.stderr, flush=True)
            gp.log_likelihood_batch(P, y)
            print(f"# measured N={N} R={R} C={C}", file=sys.stderr, flush=True)
            t0 = time.perf_counter()
            gp.log_likelihood_batch(P, y)
            print(f"# wall {1e6 * (time.perf_counter() - t0):.1f} us (includes the profile read-back)", file=sys.stderr, flush=True)
