#!/usr/bin/env python
"""Driver for an ncu capture of the device optimiser kernels: one Powell fit of the hyper-parameters (N=70, d=2,
3 restarts) and one Nelder-Mead utility minimisation (5 starts), nothing else on the GPU."""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels  # noqa: E402

N, d = 70, 2
rng = np.random.default_rng(N)
X = rng.uniform(-5, 5, size=(N, d))
y = -0.5 * np.sum((X / 2.0) ** 2, axis=1) + 0.1 * rng.standard_normal(N)
gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, 4.0), ndim=d), fit_mean=True, mean=float(np.median(y)), white_noise=-12.0)
gp.compute(X, y=y)
np.random.seed(1)
P0 = np.array([[np.median(y)] + list(np.random.randn(d)) for _ in range(3)])
for _ in range(2):
    p, f, nfev = gp.minimize_nll(P0, y, method="powell")
S = rng.uniform(-4, 4, size=(5, d))
for _ in range(2):
    x, fu, nu = gp.minimize_utility(y, S, "bape", bounds=[(-5.0, 5.0)] * d, options={"adaptive": True})
print("nll nfev", nfev, "utility nfev", nu)
