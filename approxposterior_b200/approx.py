"""``ApproxPosterior``: mirror of reference ``approxposterior/approx.py`` on the B200 engine.

Same constructor, ``run`` / ``findNextPoint`` / ``runMCMC`` / ``findMAP`` / ``bayesOpt`` / ``optGP``
methods, keyword names, defaults, cache files and (as long as no utility restart is retried, see
``utility.minimizeObjective``) RNG consumption order as the reference class
(approx.py:30-1151); george.GP becomes ``approxposterior_b200.GP`` and emcee.EnsembleSampler becomes
``approxposterior_b200.sampler.EnsembleSampler``.  What changes is *how* the three hot loops run:

  * every lnprob call of the MCMC (approx.py:148-189) is batched -- per half-step on the "host-rng"
    engine, or the whole chain in one kernel on the "device" engine (box priors);
  * the utility restarts of findNextPoint (approx.py:664-672) advance in lock step, one fused
    predict+utility launch per optimiser round, or are replaced by a device-side candidate scan
    (``scanCandidates=``);
  * optGP's restarts (approx.py:222-225) share one batched Cholesky per optimiser round.

Deviations from the reference, all deliberate and listed in DESIGN.md: emcee's HDF5 chain files are
written by the package's own minimal HDF5 writer (``hdf5min``, emcee's layout; h5py is not available
offline) next to an ``.npz`` twin; ``bayesOpt(cache=True)`` no longer dies on the missing ``self.gpPar``
(approx.py:709-710).
"""
import time

import numpy as np

from . import gpUtils
from . import mcmcUtils
from . import utility as ut
from .gp import GP
from .sampler import EnsembleSampler

__all__ = ["ApproxPosterior"]

try:
    import tqdm as _tqdm

    def _progress(it):
        return _tqdm.tqdm(it)
except Exception:      # pragma: no cover
    def _progress(it):
        return it


class ApproxPosterior(object):
    """GP-surrogate posterior estimation / Bayesian optimisation (reference approx.py:30-76)."""

    def __init__(self, theta, y, lnprior, lnlike, priorSample, bounds, gp=None, algorithm="bape"):
        if theta is None or y is None:
            raise ValueError("Must supply both theta and y for initial GP training set.")
        self.theta = np.array(theta).squeeze()
        self.y = np.array(y).squeeze()
        self.ndim = 1 if self.theta.ndim <= 1 else np.asarray(theta).shape[-1]
        if np.any(~np.isfinite(self.theta)) or np.any(~np.isfinite(self.y)):
            print("theta, y:", theta, y)
            raise ValueError("All theta and y values must be finite!")
        if len(bounds) != self.ndim:
            raise ValueError("ERROR: bounds provided but len(bounds) != ndim.\nndim = %d, len(bounds) = %d"
                             % (self.ndim, len(bounds)))
        self.bounds = bounds
        self._lnprior = lnprior
        self._lnlike = lnlike
        self.priorSample = priorSample
        self.algorithm = str(algorithm).lower()
        table = {"bape": ut.BAPEUtility, "agp": ut.AGPUtility, "alternate": ut.AGPUtility, "jones": ut.JonesUtility}
        if self.algorithm not in table:
            raise ValueError("Unknown algorithm. Valid options: bape, agp, jones, or alternate.")
        self.utility = table[self.algorithm]
        self.iburns, self.ithins, self.backends = list(), list(), list()
        self.sampler = None
        if gp is None:
            print("INFO: No GP specified. Initializing GP using ExpSquaredKernel.")
            self.gp = gpUtils.defaultGP(self.theta, self.y)
        else:
            self.gp = gp

    # ------------------------------------------------------------------ surrogate log-probability
    def _gpll(self, theta, *args, **kwargs):
        """(mu, lnprior) at one point; (-inf, nan) when theta is non-finite, the prior rejects it,
        or the GP mean is not finite (reference approx.py:148-189)."""
        if not np.any(np.isfinite(theta)):
            return -np.inf, np.nan
        lnprior = self._lnprior(theta)
        if not np.isfinite(lnprior):
            return -np.inf, np.nan
        try:
            mu = self.gp.predict(self.y, np.array(theta).reshape(1, -1), return_cov=False, return_var=False)
        except ValueError:
            return -np.inf, np.nan
        if not np.isfinite(mu):
            return -np.inf, np.nan
        return mu, lnprior

    def _gpll_batch(self, thetas, *args, **kwargs):
        """Row-wise ``_gpll`` with ONE batched mean-only predict for all rows the prior admits."""
        T = np.asarray(thetas, dtype=np.float64).reshape(-1, self.ndim)
        lp = np.full(T.shape[0], -np.inf)
        blob = np.full(T.shape[0], np.nan)
        pri = np.array([self._lnprior(t) if np.any(np.isfinite(t)) else -np.inf for t in T], dtype=np.float64)
        ok = np.isfinite(pri) & np.all(np.isfinite(T), axis=1)
        if np.any(ok):
            mu = np.asarray(self.gp.predict(self.y, T[ok], return_cov=False, return_var=False))
            fin = np.isfinite(mu)
            idx = np.nonzero(ok)[0][fin]
            lp[idx] = mu[fin]
            blob[idx] = pri[idx]
        return lp, blob

    def optGP(self, seed=None, method="powell", options=None, p0=None, nGPRestarts=1,
              gpHyperPrior=gpUtils.defaultHyperPrior):
        """Re-optimise the GP hyper-parameters (reference approx.py:192-226)."""
        self.gp = gpUtils.optimizeGP(self.gp, self.theta, self.y, seed=seed, method=method, options=options,
                                     p0=p0, nGPRestarts=nGPRestarts, gpHyperPrior=gpHyperPrior)

    # ------------------------------------------------------------------ BAPE / AGP main loop
    def run(self, m=10, nmax=2, seed=None, timing=False, verbose=True, mcmcKwargs=None, samplerKwargs=None,
            estBurnin=False, thinChains=False, runName="apRun", cache=True, gpMethod="powell", gpOptions=None,
            gpP0=None, optGPEveryN=1, nGPRestarts=1, nMinObjRestarts=5, onlyLastMCMC=False, initGPOpt=True,
            kmax=3, gpHyperPrior=gpUtils.defaultHyperPrior, eps=1.0, convergenceCheck=False,
            minObjMethod="nelder-mead", minObjOptions=None, args=None, scanCandidates=None, **kwargs):
        """Reference approx.py:229-524: nmax iterations of {find m new design points, retrain, MCMC}."""
        if cache:
            np.savez(str(runName) + "APFModelCache.npz", theta=self.theta, y=self.y)
            self.gpPar = list()
        if seed is not None:
            np.random.seed(seed)
        if timing:
            self.trainingTime, self.mcmcTime = list(), list()
        if convergenceCheck:
            self.marginalMeans, self.marginalStds, self.marginalZScores = list(), list(), list()
        if initGPOpt:
            self.optGP(seed=seed, method=gpMethod, options=gpOptions, p0=gpP0, nGPRestarts=nGPRestarts,
                       gpHyperPrior=gpHyperPrior)
        kk = 0
        if convergenceCheck and onlyLastMCMC:
            raise RuntimeError("If convergenceCheck is True, must run an MCMC each iteration.\n"
                               "convergenceCheck = %d onlyLastMCMC = %d" % (convergenceCheck, onlyLastMCMC))

        for nn in range(nmax):
            if verbose:
                print("Iteration: %d" % nn)
            start = time.time()
            self.findNextPoint(computeLnLike=True, seed=seed, cache=cache, gpMethod=gpMethod, gpOptions=gpOptions,
                               nGPRestarts=nGPRestarts, nMinObjRestarts=nMinObjRestarts, optGPEveryN=optGPEveryN,
                               numNewPoints=m, gpHyperPrior=gpHyperPrior, minObjMethod=minObjMethod,
                               minObjOptions=minObjOptions, runName=runName, theta0=None, args=args,
                               verbose=verbose, scanCandidates=scanCandidates, **kwargs)
            if timing:
                self.trainingTime.append(time.time() - start)
            if cache:
                np.savez(str(runName) + "APGP.npz", gpParamNames=self.gp.get_parameter_names(),
                         gpParamValues=self.gpPar)
            if onlyLastMCMC and nn != (nmax - 1):
                self.sampler = None
                continue

            start = time.time()
            self.sampler, iburn, ithin = self.runMCMC(samplerKwargs=samplerKwargs, mcmcKwargs=mcmcKwargs,
                                                      runName=str(runName) + str(nn), cache=cache,
                                                      estBurnin=estBurnin, thinChains=thinChains, verbose=verbose,
                                                      args=args, kwargs=kwargs)
            self.iburns.append(iburn)
            self.ithins.append(ithin)
            if timing:
                self.mcmcTime.append(time.time() - start)
                if cache:
                    np.savez(str(runName) + "APTiming.npz", trainingTime=self.trainingTime, mcmcTime=self.mcmcTime)

            if convergenceCheck:
                samples = self.sampler.get_chain(discard=self.iburns[-1], flat=True, thin=self.ithins[-1])
                meanNN, stdNN = np.mean(samples, axis=0), np.std(samples, axis=0)
                self.marginalMeans.append(meanNN)
                self.marginalStds.append(stdNN)
                if nn > 0:
                    zScore = np.fabs((meanNN - meanPrev) / stdPrev)
                    kk = kk + 1 if np.all(zScore < eps) else 0
                meanPrev, stdPrev = meanNN, stdNN
                if cache:
                    np.savez(str(runName) + "ConvergenceCache.npz", means=self.marginalMeans,
                             stds=self.marginalStds, zscores=self.marginalZScores, eps=eps, kmax=kmax,
                             finalIteration=kk)
                if kk >= kmax:
                    # the reference only leaves the loop when verbose (approx.py:518-523); kept
                    if verbose:
                        print("Approximate marginal posterior distributions converged.")
                        print("Delta zScore threshold, eps: %e" % eps)
                        print("kk, kmax: %d, %d" % (kk, kmax))
                        print("Final abs(zScore):", zScore)
                        break

    # ------------------------------------------------------------------ design-point selection
    def _selectPoint(self, theta0, nMinObjRestarts, minObjMethod, minObjOptions, scanCandidates, scanPolish="device"):
        """Next design point.  Default: the reference's multistart local minimisation (lock-step batched).
        ``scanCandidates=n``: score n uniform candidates from ``bounds`` on the device in one launch, then
        polish the winner either with a device-side shrinking-box search (``scanPolish="device"``) or with
        the reference's local optimiser (``scanPolish="nelder-mead"`` etc.)."""
        if scanCandidates:
            kind = {ut.AGPUtility: "agp", ut.BAPEUtility: "bape", ut.JonesUtility: "jones"}[self.utility]
            seed = int(np.random.randint(0, 2 ** 31 - 1))
            if scanPolish == "device":
                best, ubest, _, _ = ut.scanUtility(self.gp, self.y, kind, self.bounds, nCandidates=int(scanCandidates),
                                                   seed=seed, device_out=True, refineRounds=12)
                if np.isfinite(self._lnprior(best)):
                    return best, ubest
                scanPolish = minObjMethod      # the Python prior is stricter than the box: fall through
            best, ubest, _, _ = ut.scanUtility(self.gp, self.y, kind, self.bounds, nCandidates=int(scanCandidates),
                                               seed=seed, device_out=True)
            thetaT, uT = ut.minimizeObjective(self.utility, self.y, self.gp, sampleFn=self.priorSample,
                                              priorFn=self._lnprior, nRestarts=1, method=scanPolish,
                                              options=minObjOptions, bounds=self.bounds, theta0=None,
                                              args=(self.y, self.gp, self._lnprior), _start=best)
            if not (uT <= ubest):
                thetaT, uT = best, ubest
            return thetaT, uT
        return ut.minimizeObjective(self.utility, self.y, self.gp, sampleFn=self.priorSample,
                                    priorFn=self._lnprior, nRestarts=nMinObjRestarts, method=minObjMethod,
                                    options=minObjOptions, bounds=self.bounds, theta0=theta0,
                                    args=(self.y, self.gp, self._lnprior))

    def findNextPoint(self, theta0=None, computeLnLike=True, seed=None, cache=True, gpOptions=None, gpP0=None,
                      verbose=True, nGPRestarts=1, nMinObjRestarts=5, gpMethod="powell",
                      minObjMethod="nelder-mead", minObjOptions=None, runName="apRun", numNewPoints=1,
                      optGPEveryN=1, gpHyperPrior=gpUtils.defaultHyperPrior, args=None, scanCandidates=None,
                      scanPolish="device", **kwargs):
        """Pick numNewPoints design points by minimising the (negative) utility, optionally evaluate the
        forward model there, grow the training set and refactor / re-optimise the GP
        (reference approx.py:527-754)."""
        assert (isinstance(numNewPoints, int) and (numNewPoints >= 1))
        assert (isinstance(optGPEveryN, int) and (optGPEveryN >= 1))
        if verbose and numNewPoints < optGPEveryN:
            print("WARNING: numNewPoints < optGPEveryN. GP hyperparameters will not be re-optimized.")
        if args is None:
            args = ()
        newTheta, newY = list(), list()

        for ii in (_progress(range(numNewPoints)) if verbose else range(numNewPoints)):
            if self.algorithm == "alternate":
                self.utility = ut.AGPUtility if ii % 2 == 0 else ut.BAPEUtility
            thetaT, uT = self._selectPoint(theta0, nMinObjRestarts, minObjMethod, minObjOptions, scanCandidates,
                                           scanPolish)
            newTheta.append(thetaT)
            if not computeLnLike:
                continue

            loglikeT = self._lnlike(thetaT, *args, **kwargs)
            if hasattr(loglikeT, "__iter__"):
                yT = np.array([loglikeT[0] + self._lnprior(thetaT)])
            else:
                yT = np.array([loglikeT + self._lnprior(thetaT)])
            newY.append(yT)
            if self.theta.ndim > 1:
                self.theta = np.vstack([self.theta, np.array(thetaT)])
            else:
                self.theta = np.hstack([self.theta, thetaT])
            self.y = np.hstack([self.y, yT])

            try:
                currentHype = self.gp.get_parameter_vector()
                if verbose:
                    print("hyperparameters", currentHype)
                if cache:
                    if not hasattr(self, "gpPar"):
                        self.gpPar = list()
                    self.gpPar.append(currentHype)
                if isinstance(self.gp, GP):
                    # the reference builds a fresh george.GP around the same kernel (approx.py:712-717);
                    # re-computing in place is equivalent and keeps the device handle and its buffers
                    # ... and, with unchanged hyper-parameters, a bordered O(N^2) update replaces the refactor
                    # (append_point falls back to a full compute when it cannot take the fast path)
                    self.gp.set_parameter_vector(currentHype)
                    self.gp.append_point(thetaT, yT)
                else:
                    # any other george-like object: rebuild it the way the reference does (approx.py:712-717)
                    self.gp = type(self.gp)(kernel=self.gp.kernel, fit_mean=True, mean=self.gp.mean,
                                            white_noise=self.gp.white_noise, fit_white_noise=False)
                    self.gp.set_parameter_vector(currentHype)
                    self.gp.compute(self.theta)
                if ii % optGPEveryN == 0:
                    self.optGP(seed=seed, method=gpMethod, options=gpOptions, p0=gpP0, nGPRestarts=nGPRestarts,
                               gpHyperPrior=gpHyperPrior)
            except ValueError:
                print("theta:", self.theta)
                print("y:", self.y)
                print("gp parameters names:", self.gp.get_parameter_names())
                print("gp parameters:", self.gp.get_parameter_vector())
                raise ValueError("GP couldn't optimize!")
            if cache:
                np.savez(str(runName) + "APFModelCache.npz", theta=self.theta, y=self.y)

        if numNewPoints == 1:
            newTheta = newTheta[0]
            if computeLnLike:
                newY = newY[0]
        if computeLnLike:
            return np.asarray(newTheta), np.asarray(newY)
        return np.asarray(newTheta)

    # ------------------------------------------------------------------ MCMC on the surrogate
    def runMCMC(self, samplerKwargs=None, mcmcKwargs=None, runName="apRun", cache=True, estBurnin=True,
                thinChains=True, verbose=False, args=None, **kwargs):
        """Sample the surrogate posterior (reference approx.py:757-859).  Returns (sampler, iburn, ithin).

        ``samplerKwargs["engine"]``: "device" (default when the log-prior exposes ``.bounds``, e.g.
        ``likelihood.BoxPrior``) or "host-rng" (emcee's NumPy RNG flow, batched lnprob)."""
        samplerKwargs, mcmcKwargs = mcmcUtils.validateMCMCKwargs(self, samplerKwargs, mcmcKwargs, verbose)
        samplerKwargs = dict(samplerKwargs)
        engine = samplerKwargs.pop("engine", None)
        box = getattr(self._lnprior, "bounds", None)
        if engine is None:
            engine = "device" if box is not None else "host-rng"
        backend = None
        if cache:
            bname = str(runName) + ".h5"           # as the reference (approx.py:830); written by hdf5min, with an .npz twin
            self.backends.append(bname)
            backend = bname
        samplerKwargs["log_prob_fn"] = self._gpll_batch
        self.sampler = EnsembleSampler(**samplerKwargs, backend=backend, args=args, kwargs=kwargs,
                                       blobs_dtype=[("lnprior", float)], engine=engine, gp=self.gp, y=self.y,
                                       bounds=box if box is not None else self.bounds,
                                       lnprior_const=getattr(self._lnprior, "value", 0.0))
        for _ in self.sampler.sample(**mcmcKwargs):
            pass
        if verbose:
            print("mcmc finished")
        iburn, ithin = mcmcUtils.estimateBurnin(self.sampler, estBurnin=estBurnin, thinChains=thinChains,
                                                verbose=verbose)
        return self.sampler, iburn, ithin

    # ------------------------------------------------------------------ MAP and Bayesian optimisation
    def findMAP(self, theta0=None, method="nelder-mead", options=None, nRestarts=15):
        """Maximise the GP mean under the prior (reference approx.py:862-926)."""
        if theta0 is not None:
            theta0 = np.array(theta0).reshape(1, self.theta.shape[-1])
        else:
            theta0 = self.theta[np.argmax(self.y)]
        if str(method).lower() == "nelder-mead" and options is None:
            options = {"adaptive": True}

        def fn(x):
            if not np.isfinite(self._lnprior(x)):
                return np.inf
            return -(self._gpll(x)[0])

        fn.batch = lambda T: -self._gpll_batch(T)[0]      # all restarts' calls in one mean-only predict
        fn.device_kind = "negmean"                        # ... or the whole multistart on the device (box priors)

        MAP, MAPVal = ut.minimizeObjective(fn, self.y, self.gp, self.priorSample, self._lnprior,
                                           nRestarts=nRestarts, args=None, method=method, options=options,
                                           bounds=self.bounds, theta0=theta0)
        return MAP, -MAPVal

    def bayesOpt(self, nmax, theta0=None, tol=1.0e-3, kmax=3, seed=None, verbose=True, runName="apRun",
                 cache=True, gpMethod="powell", gpOptions=None, gpP0=None, optGPEveryN=1, nGPRestarts=1,
                 nMinObjRestarts=5, initGPOpt=True, minObjMethod="nelder-mead",
                 gpHyperPrior=gpUtils.defaultHyperPrior, minObjOptions=None, findMAP=True, args=None,
                 scanCandidates=None, **kwargs):
        """Bayesian optimisation loop (reference approx.py:929-1151): one new design point per
        iteration, stop when |delta best y| < tol for kmax consecutive iterations."""
        thetas, vals = list(), list()
        thetasMAP, valsMAP = list(), list()
        if cache:
            np.savez(str(runName) + "APFModelCache.npz", theta=self.theta, y=self.y)
        if seed is not None:
            np.random.seed(seed)
        if initGPOpt:
            self.optGP(seed=seed, method=gpMethod, options=gpOptions, p0=gpP0, nGPRestarts=nGPRestarts,
                       gpHyperPrior=gpHyperPrior)
        kk = 0
        for nn in range(nmax):
            if verbose:
                print("Iteration: %d" % nn)
            optN = 1 if nn % optGPEveryN == 0 else 99999999
            thetaT, yT = self.findNextPoint(computeLnLike=True, seed=seed, cache=cache, gpMethod=gpMethod,
                                            gpOptions=gpOptions, nGPRestarts=nGPRestarts,
                                            nMinObjRestarts=nMinObjRestarts, optGPEveryN=optN, numNewPoints=1,
                                            gpHyperPrior=gpHyperPrior, minObjMethod=minObjMethod,
                                            minObjOptions=minObjOptions, runName=runName, args=args,
                                            verbose=verbose, scanCandidates=scanCandidates, **kwargs)
            if verbose:
                print("Forward model evaluation at: ", thetaT, ", function value: ", yT)
            if cache:
                np.savez(str(runName) + "APGP.npz", gpParamNames=self.gp.get_parameter_names(),
                         gpParamValues=self.gp.get_parameter_vector())
            ibest = np.argmax(self.y)
            thetas.append(self.theta[ibest])
            vals.append(self.y[ibest])
            if findMAP:
                thetaN, valN = self.findMAP(theta0=theta0, method=minObjMethod, options=minObjOptions,
                                            nRestarts=nMinObjRestarts)
                if verbose:
                    print("Current MAP solution: ", thetaN, valN)
                thetasMAP.append(thetaN)
                valsMAP.append(valN)
            if nn > 0:
                kk = kk + 1 if np.fabs(vals[-1] - vals[-2]) < tol else 0
                if kk >= kmax:
                    break

        soln = {"thetaBest": thetas[-1], "valBest": vals[-1], "thetas": np.asarray(thetas).squeeze(),
                "vals": np.asarray(vals).squeeze(), "nev": nn + 1}
        if findMAP:
            soln["thetasMAP"] = np.asarray(thetasMAP).squeeze()
            soln["valsMAP"] = np.asarray(valsMAP).squeeze()
            soln["thetaMAPBest"] = soln["thetasMAP"][np.argmax(soln["valsMAP"])]
            soln["valMAPBest"] = soln["valsMAP"][np.argmax(soln["valsMAP"])]
        return soln
