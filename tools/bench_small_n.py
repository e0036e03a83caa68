#!/usr/bin/env python
"""Variance kernel in the N <= 1024 regime (d = 5, Q = 2^20): one tile per CTA vs G CTAs per query tile, device-timed,
with the structural bound of the FP64 pipe next to it (executed DMMA FMAs incl. the diagonal-block remainder + the
(2d + 12) N lane-ops of the K* panel, over the algorithmic N^2 + (3d + 6) N flops), and the end-to-end rate of the
host-buffer call (pinned buffers) with and without the three-stream pipeline.  One JSON object per line."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels
dev = torch.device("cuda", 0)


def dgemm_peak():
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev); b = torch.randn_like(a)
    torch.matmul(a, b); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2 * 8192 ** 3 / best * 1e-9


PEAK = dgemm_peak()            # GFLOP/ms = TFLOP/s
d, Q = 5, 1 << 20
for N in [int(v) for v in (sys.argv[1:] or ["256", "384", "512", "768", "1024"])]:
    rng = np.random.default_rng(N + d)
    X = rng.uniform(-5, 5, size=(N, d)); y = rng.standard_normal(N)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.full(d, float(d)), ndim=d), fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    q = -5 + 10 * torch.rand((Q, d), dtype=torch.float64, device=dev)
    nblk = (N + 63) // 64
    executed_fma = 4096.0 * (nblk * (nblk - 1) / 2.0 + 0.5625 * nblk) + (2 * d + 12) * nblk * 64 / 2.0   # lane-ops/2 ~ FMA slots
    algorithmic = (float(N) * N + (3 * d + 6) * N)
    out = dict(config="small-n", N=N, d=d, Q=Q, dgemm_tflops=PEAK, structural_bound_frac=algorithmic / 2.0 / (4096.0 * (nblk * (nblk - 1) / 2.0 + 0.5625 * nblk) + (2 * d + 12) * nblk * 64))
    for group in (0, 2, 4, -1):
        if group > 0 and 2 * group > nblk:
            continue
        gp.set_group(group)
        gp._predict_raw(q, True); torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); gp._predict_raw(q, True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = min(ts)
        out["ms_group_%s" % {0: "off", -1: "auto"}.get(group, group)] = ms
        out["frac_dgemm_group_%s" % {0: "off", -1: "auto"}.get(group, group)] = algorithmic * Q / ms * 1e-9 / PEAK
    gp.set_group(-1)
    # end to end through host buffers (pinned): pipelined (default) vs serial copies
    hq = torch.empty((Q, d), dtype=torch.float64).pin_memory(); hq.copy_(q.cpu())
    outs = (torch.empty(Q, dtype=torch.float64, pin_memory=True).numpy(), torch.empty(Q, dtype=torch.float64, pin_memory=True).numpy(), None)
    for label, env in (("pipelined", None), ("serial", "1")):
        if env:
            os.environ["APGP_NO_PIPELINE"] = env
        else:
            os.environ.pop("APGP_NO_PIPELINE", None)
        gp._predict_raw(hq.numpy(), True, out=outs)
        t0 = time.perf_counter()
        for _ in range(5):
            gp._predict_raw(hq.numpy(), True, out=outs)
        dt = (time.perf_counter() - t0) / 5
        out["e2e_ms_%s" % label] = dt * 1e3
    os.environ.pop("APGP_NO_PIPELINE", None)
    out["e2e_over_device"] = out["ms_group_auto"] / out["e2e_ms_pipelined"]
    # the pipelined result equals the device-resident one
    mu_d, var_d, _ = gp._predict_raw(q, True)
    assert np.array_equal(outs[0], mu_d.cpu().numpy()) and np.array_equal(outs[1], var_d.cpu().numpy())
    print(json.dumps(out), flush=True)
    del gp, q
    torch.cuda.empty_cache()
