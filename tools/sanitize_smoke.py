"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import GP, kernels

rng = np.random.default_rng(0)
for N, d, amp in ((70, 2, None), (200, 3, 2.0), (300, 5, None)):
    X = rng.uniform(-5, 5, size=(N, d)); y = np.sin(X).sum(axis=1)
    k = kernels.ExpSquaredKernel(np.full(d, 3.0), ndim=d)
    if amp:
        k = amp * k
    gp = GP(kernel=k, fit_mean=True, mean=0.0, white_noise=-12.0)
    gp.compute(X, y=y)
    q = rng.uniform(-5, 5, size=(700, d))
    gp.predict(y, q, return_cov=False, return_var=True)
    gp.predict(y, q, return_cov=False, return_var=False)
    gp.predict_utility(y, q, "jones", bounds=[(-5, 5)] * d)
    P = np.vstack([gp.get_parameter_vector() + 0.1 * i for i in range(5)])
    gp.log_likelihood_batch(P, y)
    gp.log_likelihood_batch(P, y, return_grad=(N <= 224))
    gp.grad_log_likelihood(y)
    gp.append_point(rng.uniform(-5, 5, size=d), 0.3)
    y2 = np.concatenate([y, [0.3]])
    gp.predict(y2, q[:100], return_cov=False, return_var=True)
    p0 = rng.uniform(-5, 5, size=(2 * 4 * d, d))
    gp.run_ensembles(y2, p0, 20, [(-5, 5)] * d, nens=2, seed=3)
    os.environ["APGP_LOGLIK_TILED"] = "1"
    gp.log_likelihood_batch(P, y2)
    del os.environ["APGP_LOGLIK_TILED"]
print("sanitize_smoke ok")
