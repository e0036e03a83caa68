"""cfg2-shaped device sampler launch for ncu: 2048 ensembles x 32 walkers, N=1024, d=2, 200 steps."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import gpUtils, likelihood as lh
np.random.seed(57)
theta = lh.rosenbrockSample(1024)
y = np.array([lh.rosenbrockLnlike(t) for t in theta])
gp = gpUtils.defaultGP(theta, y)
gp.set_parameter_vector([float(np.median(y)), 0.5, 1.2]); gp.recompute()
p0 = np.random.uniform(-5, 5, size=(2048 * 32, 2))
for _ in range(2):
    gp.run_ensembles(y, p0, 200, [(-5, 5), (-5, 5)], nens=2048, seed=1, thin=20)
