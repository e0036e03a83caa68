import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import synthetic_gp_problem
from approxposterior_b200 import GP, kernels
for N, d, Q, group in [(1024, 2, 40000, -1), (2048, 5, 70001, 16), (2048, 5, 70001, 8), (1100, 3, 30011, 6), (1100, 3, 700, 4), (4096, 2, 20000, -1)]:
    X, y, logM, _ = synthetic_gp_problem(N, d, seed=N + d)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=d), fit_mean=True, mean=0.1, white_noise=-12.0)
    gp.compute(X, y=y)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    q = -5.5 + 11.0 * torch.rand((Q, d), dtype=torch.float64, device="cuda", generator=g)
    gp.set_group(0)
    mu0, var0, u0 = gp._predict_raw(q, True, utility="agp", bounds=[(-5.0, 5.0)] * d)
    gp.set_group(group)
    mu1, var1, u1 = gp._predict_raw(q, True, utility="agp", bounds=[(-5.0, 5.0)] * d)
    torch.cuda.synchronize()
    a = np.abs(gp._alpha()).max()
    print(N, d, Q, group, "max|dmu|", float((mu1 - mu0).abs().max()), "max|dvar|", float((var1 - var0).abs().max()),
          "max|alpha|", a, "nan mu", int(torch.isnan(mu1).sum()), int(torch.isnan(mu0).sum()), flush=True)
