#!/usr/bin/env python
"""approxposterior's README example (reference README.md:80-138) on the B200 engine.

    python examples/quickstart.py

Identical to the reference script except for the import line and the optional BoxPrior (which lets the
whole MCMC chain run inside one CUDA kernel instead of calling the Python prior once per walker)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from approxposterior_b200 import approx, gpUtils, likelihood as lh   # reference: from approxposterior import ...

m0 = 50                            # initial size of the training set
m = 20                             # new design points per iteration
nmax = 2                           # iterations
bounds = [(-5, 5), (-5, 5)]        # prior bounds
algorithm = "bape"                 # Kandasamy et al. (2015)
seed = 57
np.random.seed(seed)

samplerKwargs = {"nwalkers": 20}                 # emcee.EnsembleSampler parameters
mcmcKwargs = {"iterations": int(2.0e4)}          # emcee.EnsembleSampler.run_mcmc parameters

theta = lh.rosenbrockSample(m0)
y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])

gp = gpUtils.defaultGP(theta, y, white_noise=-12)

ap = approx.ApproxPosterior(theta=theta, y=y, gp=gp,
                            lnprior=lh.BoxPrior(bounds),          # or lh.rosenbrockLnprior (host-side prior)
                            lnlike=lh.rosenbrockLnlike,
                            priorSample=lh.rosenbrockSample,
                            bounds=bounds, algorithm=algorithm)

ap.run(m=m, nmax=nmax, estBurnin=True, nGPRestarts=3, mcmcKwargs=mcmcKwargs, cache=False,
       samplerKwargs=samplerKwargs, verbose=True, thinChains=False, onlyLastMCMC=True, timing=True)

samples = ap.sampler.get_chain(discard=ap.iburns[-1], flat=True, thin=ap.ithins[-1])
print("training set:", ap.theta.shape, " posterior samples:", samples.shape)
print("posterior mean:", samples.mean(axis=0), " std:", samples.std(axis=0))
print("seconds per BAPE iteration:", np.round(ap.trainingTime, 2), " MCMC:", np.round(ap.mcmcTime, 2))
