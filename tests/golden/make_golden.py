"""Generate the committed golden vectors from the CPU oracle (run here, in the authoring container).

The reference itself cannot be imported (george/emcee are not installable offline), so these vectors
come from the oracle *after* it has been pinned against the reference's KATs (tests/test_oracle_kat.py).
They freeze the oracle's output so the GPU box compares the CUDA path against committed numbers, not
only against a live oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import GPOracle, agp_utility, bape_utility, jones_utility  # noqa: E402


def case(name, N, d, Q, seed, amp=None, logM=None):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-5, 5, size=(N, d))
    y = np.sin(X).sum(axis=1) - 0.1 * np.sum(X * X, axis=1)
    logM = np.full(d, np.log(d)) if logM is None else np.asarray(logM, float)
    mean = float(np.median(y))
    gp = GPOracle(d, np.exp(logM), amp=amp, mean=mean, white_noise=-12.0)
    gp.compute(X)
    Xq = rng.uniform(-5.5, 5.5, size=(Q, d))
    Xq[:4] = X[:4]
    mu, var = gp.predict(y, Xq, return_var=True)
    ok = np.all(np.abs(Xq) <= 5, axis=1)
    P = np.vstack([gp.get_parameter_vector() + 0.3 * k for k in range(-2, 3)])
    ll = []
    for p in P:
        gp.set_parameter_vector(p)
        ll.append(gp.log_likelihood(y, quiet=True))
    gp.set_parameter_vector(P[2])
    grad = gp.grad_log_likelihood(y)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), X=X, y=y, logM=logM, mean=mean,
                        amp=np.nan if amp is None else amp, Xq=Xq, mu=mu, var=var,
                        agp=agp_utility(mu, var, ok), bape=bape_utility(mu, var, ok),
                        jones=jones_utility(mu, var, y.max(), 0.01, ok), P=P, ll=np.array(ll), grad=grad)


if __name__ == "__main__":
    case("golden_n96_d2", 96, 2, 300, 1)
    case("golden_n256_d2_amp", 256, 2, 300, 2, amp=3.5)
    case("golden_n400_d5", 400, 5, 300, 3)
    case("golden_n130_d10_amp", 130, 10, 200, 4, amp=0.7)
    case("golden_n64_d1", 64, 1, 200, 5, logM=[0.3])
    print("golden vectors written to", HERE)
