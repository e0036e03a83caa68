import sys, os, time, json
import numpy as np
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
from approxposterior_b200 import GP, kernels
for N in (300, 512, 1024, 2048):
    d = 5
    rng = np.random.default_rng(N)
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    out = dict(N=N)
    for Q in (5, 256, 4096):
        q = rng.uniform(-5, 5, size=(Q, d))
        for group in (0, -1):
            gp.set_group(group)
            gp.predict_utility(y, q, "bape")
            t0 = time.perf_counter()
            for _ in range(20):
                gp.predict_utility(y, q, "bape")
            out["Q%d_%s_us" % (Q, "off" if group == 0 else "auto")] = (time.perf_counter() - t0) / 20 * 1e6
    print(json.dumps(out), flush=True)
