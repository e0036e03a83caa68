#!/usr/bin/env python
"""Latency of a predict call with a handful of queries from host memory (mean + variance + BAPE utility): the
few-query kernel (default for Q <= 16: S CTAs per query against the explicit inverse) against the tiled DMMA kernels
(APGP_PREDICT_FEW=0), each in its own process (the flag is read when a handle is created), and the agreement of the two.
One JSON object per line."""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = ((70, 2), (300, 5), (512, 5), (1024, 5), (2048, 5), (4096, 5))
QS = (1, 5, 16)


def child():
    from approxposterior_b200 import GP, kernels
    for N, d in SHAPES:
        rng = np.random.default_rng(N)
        X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
        gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
        gp.compute(X, y=y)
        out = dict(N=N, d=d)
        for Q in QS:
            q = rng.uniform(-5, 5, size=(Q, d))
            r = gp.predict_utility(y, q, "bape")
            t0 = time.perf_counter()
            for _ in range(50):
                gp.predict_utility(y, q, "bape")
            out["Q%d_us" % Q] = (time.perf_counter() - t0) / 50 * 1e6
            out["Q%d_out" % Q] = [np.asarray(v).tolist() for v in r]
        print(json.dumps(out), flush=True)


if len(sys.argv) > 1 and sys.argv[1] == "--child":
    child()
    sys.exit(0)
runs = {}
for few in ("1", "0"):
    env = dict(os.environ, APGP_PREDICT_FEW=few)
    txt = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True).stdout
    runs[few] = [json.loads(l) for l in txt.splitlines() if l.startswith("{")]
for a, b in zip(runs["1"], runs["0"]):
    out = dict(N=a["N"], d=a["d"])
    for Q in QS:
        out["Q%d_few_us" % Q] = round(a["Q%d_us" % Q], 1)
        out["Q%d_tiled_us" % Q] = round(b["Q%d_us" % Q], 1)
        u, v = np.array(a["Q%d_out" % Q]), np.array(b["Q%d_out" % Q])
        out["Q%d_max_rel_diff" % Q] = float(np.max(np.abs(u - v) / (np.abs(v) + 1e-300)))
    print(json.dumps(out), flush=True)
