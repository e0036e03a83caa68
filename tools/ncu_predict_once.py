#!/usr/bin/env python
"""One fused predict pass (cfg3: N=2048, d=5, 2^20 candidates, BAPE) for ncu captures; APGP_PREDICT_GROUP selects
the one-tile-per-CTA kernel (0) or the grouped kernel (G / -1)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from approxposterior_b200 import GP, kernels
X, y, logM, mean = bench.make_problem()
gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=bench.DIM), fit_mean=True, mean=mean, white_noise=-12.0)
gp.compute(X, y=y)
q = -5 + 10 * torch.rand((1 << 20, bench.DIM), dtype=torch.float64, device="cuda")
for _ in range(int(os.environ.get("REPS", "2"))):
    gp._predict_raw(q, True, utility="bape", bounds=bench.BOUNDS, ybest=float(np.max(y)))
torch.cuda.synchronize()
