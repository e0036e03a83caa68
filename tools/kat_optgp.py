import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from approxposterior_b200 import gpUtils, likelihood as lh
for fitAmp in (True, False):
    np.random.seed(57)
    theta = np.array(lh.rosenbrockSample(50))
    y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
    gp = gpUtils.defaultGP(theta, y, fitAmp=fitAmp)
    gp = gpUtils.optimizeGP(gp, theta, y, seed=57, nGPRestarts=5, method="powell")
    print(fitAmp, repr(gp.get_parameter_vector()), gp.log_likelihood(y), gpUtils.optimizeGP.last_stats)
