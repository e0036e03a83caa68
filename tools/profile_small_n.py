"""predict_var at N=256 and N=512 (d=5, Q=2^20) for the ncu capture of the small-N regime."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
dev = torch.device("cuda", 0)
for N in (256, 512):
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, 5)); y = rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(5, 5.0), ndim=5), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    q = -5 + 10 * torch.rand((1 << 20, 5), dtype=torch.float64, device=dev)
    for _ in range(3):
        gp._predict_raw(q, True, utility="bape")
    torch.cuda.synchronize()
