"""Measured forward error of the engine's predictive mean / variance against extended-precision truth, next to the
LAPACK oracle's, on the ill-conditioned parity shapes.  Run once per library build:

    python tools/measure_parity_errors.py                       # current libapgp.so
    APGP_LIB=approxposterior_b200/libapgp_r01_olddiag.so python tools/measure_parity_errors.py
        # the library as of 9d0af3d^ (one-thread-per-row 64x64 diagonal kernel), built from `git archive 9d0af3d^`

Output: one JSON line per shape."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from approxposterior_b200 import _lib  # noqa: E402

if os.environ.get("APGP_LIB"):                     # an older build exports fewer symbols: bind what exists
    import ctypes as C
    lib = C.CDLL(_lib.LIB_PATH)
    for name in list(_lib._SIGNATURES):
        if not hasattr(lib, name):
            del _lib._SIGNATURES[name]

from conftest import extended_truth, synthetic_gp_problem  # noqa: E402
from approxposterior_b200 import GP, kernels  # noqa: E402
from oracle import GPOracle  # noqa: E402

for N, d in [(20, 2), (64, 1), (256, 2), (700, 5), (2100, 20)]:
    X, y, logM, _ = synthetic_gp_problem(N, d, seed=N + d)
    mean = float(np.median(y))
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=d), fit_mean=True, mean=mean, white_noise=-12.0)
    gp.compute(X, y=y)
    orc = GPOracle(d, np.exp(logM), mean=mean, white_noise=-12.0)
    orc.compute(X)
    Xq = np.random.default_rng(1).uniform(-5, 5, size=(777, d))
    mu_g, var_g = gp.predict(y, Xq, return_cov=False, return_var=True)
    mu_o, var_o = orc.predict(y, Xq, return_var=True)
    mu_t, var_t = extended_truth(X, y, logM, Xq, mean=mean, nvar=64)
    print(json.dumps(dict(lib=os.path.basename(_lib.LIB_PATH), N=N, d=d, cond=float(np.linalg.cond(orc._L) ** 2),
                          mu_err_engine_max=float(np.abs(mu_g - mu_t).max()), mu_err_engine_mean=float(np.abs(mu_g - mu_t).mean()),
                          mu_err_oracle_max=float(np.abs(mu_o - mu_t).max()), mu_err_oracle_mean=float(np.abs(mu_o - mu_t).mean()),
                          mu_engine_vs_oracle_max=float(np.abs(mu_g - mu_o).max()), scale=float(np.abs(y).max()),
                          var_err_engine_max=float(np.abs(var_g[:64] - var_t).max()),
                          var_err_oracle_max=float(np.abs(var_o[:64] - var_t).max()))), flush=True)
