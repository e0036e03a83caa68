"""Round-2 ncu targets, one launch each (select with argv[1]):
  group   loglik_group_kernel        cfg4 shape: N=512, d=10, R=64 hyper-parameter vectors (cluster of 2 CTAs per vector)
  acf     acf_partial_kernel         65 536 walkers x 1000 steps x d=2 device-resident chain (synthetic AR(1))
  sampler sampler_kernel             cfg2 shape, 100 steps
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from approxposterior_b200 import GP, kernels
which = sys.argv[1]
rng = np.random.default_rng(0)
if which == "group":
    N, d, R = 512, 10, 64
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    gp = GP(kernel=2.0 * kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    P = np.column_stack([np.zeros(R), 0.3 * rng.standard_normal((R, d + 1)) + np.r_[np.log(2.0 / d), np.full(d, np.log(d))]])
    os.environ["APGP_LOGLIK_PATH"] = "group"
    for _ in range(2):
        ll = gp.log_likelihood_batch(P, y)
    print("finite", int(np.isfinite(ll).sum()))
elif which == "acf":
    gp = GP(kernel=kernels.ExpSquaredKernel([1.0, 1.0], ndim=2), fit_mean=True, mean=0.0, white_noise=-12.0)
    n, W = 1000, 65536
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.empty((n, W, 2), dtype=torch.float64, device="cuda")
    x[0] = torch.randn((W, 2), dtype=torch.float64, device="cuda", generator=g)
    for t in range(1, n):
        x[t] = 0.95 * x[t - 1] + 0.31 * torch.randn((W, 2), dtype=torch.float64, device="cuda", generator=g)
    for _ in range(2):
        print(gp.integrated_time(x))
else:
    from approxposterior_b200 import gpUtils, likelihood as lh
    np.random.seed(57)
    theta = lh.rosenbrockSample(1024)
    y = np.array([lh.rosenbrockLnlike(t) for t in theta])
    gp = gpUtils.defaultGP(theta, y)
    gp.set_parameter_vector([float(np.median(y)), 0.5, 1.2]); gp.recompute()
    p0 = np.random.uniform(-5, 5, size=(2048 * 32, 2))
    for _ in range(2):
        gp.run_ensembles(y, p0, 100, [(-5, 5), (-5, 5)], nens=2048, seed=1, thin=100)
