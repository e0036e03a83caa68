"""CPU oracle for the approxposterior GP-surrogate hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``approxposterior_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker
(or as the timed CPU baseline), never as the product path.

Why a restatement: the arithmetic of the hot path lives in two third-party
packages that are neither vendored under /root/reference nor installable here
(no network): ``george`` (unpinned in reference ``setup.py:67``; release
contemporary with approxposterior 0.4 is george 0.3.1) and ``emcee>=3.0``
(reference ``setup.py:68``; 3.0.x evidenced by the warning captured in
``examples/Notebooks/fittingALine.ipynb``).  The oracle restates their published
algorithms in NumPy/SciPy fp64 and is anchored on the reference's own call sites
and known-answer tests:

* GP part (``gp_oracle.py``, ``utility_oracle.py``) -- PINNED by the reference's
  KATs ``approxposterior/tests/test_GPUtil.py:50,56,62,101,107,113`` and
  ``tests/test_InitGP.py:43,76`` (see ``tests/test_oracle_kat.py``), and through
  the optimiser drivers by ``tests/test_OptimizeGP.py:50,91`` and
  ``tests/test_findNewPoint.py:107``.
* Sampler part (``sampler_oracle.py``) -- PARITY UNPINNED: the reference has no
  bit-level golden for emcee's stretch move (``tests/test_Burnin.py:89`` is a
  1e-1 statistical check on emcee's RNG flow, ``tests/test_APRun.py:66`` a z<1
  check).  The restated move is the only bit-level pin.
"""
from .gp_oracle import GPOracle, default_gp_oracle  # noqa: F401
from .utility_oracle import (logsubexp, agp_utility, bape_utility,  # noqa: F401
                             jones_utility, UTILITY_BY_NAME)
from .sampler_oracle import stretch_move_oracle  # noqa: F401
