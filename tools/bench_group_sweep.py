#!/usr/bin/env python
"""One-tile-per-CTA (group 0) vs grouped (automatic G) variance kernel over the training-set sizes of cfg5 (d=5)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
dev = torch.device("cuda", 0)
for N in (512, 1024, 1536, 2048, 3000, 4096, 8192):
    d = 5
    rng = np.random.default_rng(N + d)
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    Q = 1 << (20 if N <= 4096 else 19)
    q = -5 + 10 * torch.rand((Q, d), dtype=torch.float64, device=dev)
    out = dict(N=N, d=d, Q=Q)
    for group in (0, -1):
        gp.set_group(group)
        gp._predict_raw(q, True, utility="bape")
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); gp._predict_raw(q, True, utility="bape"); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["ms_group_%s" % ("off" if group == 0 else "auto")] = min(ts)
    out["ratio_auto_over_off"] = out["ms_group_auto"] / out["ms_group_off"]
    print(json.dumps(out), flush=True)
    del gp, q
    torch.cuda.empty_cache()
