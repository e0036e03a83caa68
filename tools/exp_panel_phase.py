"""EXPERIMENT (not product): where does the K* panel phase of predict_var_kernel spend its time?  Times the kernel with
the panel stores skipped (flag 1), with the exponentials replaced by a copy (flag 2), and both (3), using the
instrumented build libapgp_exp.so (APGP_LIB).  Results are wrong by construction; only the times matter."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
dev = torch.device("cuda", 0)
d, Q = 5, 1 << 20
for N in (256, 512):
    rng = np.random.default_rng(N + d)
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    gp.set_group(0)
    q = -5 + 10 * torch.rand((Q, d), dtype=torch.float64, device=dev)
    out = dict(N=N)
    for flags in (0, 1, 2, 3):
        os.environ["APGP_DEBUG_FLAGS"] = str(flags)
        gp._predict_raw(q, True); torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); gp._predict_raw(q, True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out["ms_flags_%d" % flags] = min(ts)
    print(json.dumps(out), flush=True)
