"""approxposterior_b200 -- B200-native GP-surrogate engine for approxposterior's hot path.

Drop-in for the george/emcee calls made by dflemin3/approxposterior v0.4:
``GP`` (george.GP duck type), ``kernels`` (ExpSquaredKernel, constant scaling),
``gpUtils`` / ``utility`` / ``mcmcUtils`` / ``approx`` mirroring the reference modules of the
same names, and ``EnsembleSampler`` (emcee surface used by approx.py:839-847).
All numerics run in hand-written sm_100a CUDA kernels behind the C-ABI in ``include/apgp.h``;
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import kernels  # noqa: F401
from .gp import GP  # noqa: F401
