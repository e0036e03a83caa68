#!/usr/bin/env python
"""Two batched log-likelihood calls on the fused cluster path (N, R from argv; d = 10) for an ncu capture:
ncu --set full --import-source on -k regex:loglik_group -s 1 -c 1 python tools/ncu_chol_group_once.py 512 64"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from approxposterior_b200 import GP, kernels  # noqa: E402

os.environ["APGP_LOGLIK_PATH"] = "group"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
d = 10
rng = np.random.default_rng(64)
X = rng.uniform(-5, 5, size=(N, d))
y = np.sin(X).sum(axis=1)
gp = GP(kernel=float(np.var(y)) * kernels.ExpSquaredKernel(np.ones(d), ndim=d), fit_mean=True,
        mean=float(np.median(y)), white_noise=-12.0)
gp.compute(X, y=y)
P = np.column_stack([np.full(R, np.median(y)), rng.standard_normal((R, 11))])
for _ in range(2):
    ll = gp.log_likelihood_batch(P, y)
print(int(np.isfinite(ll).sum()), "finite of", R)
