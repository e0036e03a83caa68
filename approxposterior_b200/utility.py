"""Acquisition utilities and their multistart minimiser: mirror of reference ``utility.py``.

Scalar functions keep the reference's signatures ``fn(theta, y, gp, priorFn)`` and return values
(+inf outside the prior, the naive ``logsubexp``, Jones returning 0.0 when std <= 0).  On an
``approxposterior_b200.GP`` the (mu, var) -> utility epilogue is evaluated inside the fused predict
kernel; ``utilityBatch`` and ``minimizeObjective`` push all live optimiser restarts -- or a whole
candidate cloud -- through one launch.
"""
import numpy as np
from scipy.optimize import minimize
from scipy.special import ndtr

from . import _optimizers as _opt
from ._lockstep import run_lockstep

__all__ = ["logsubexp", "AGPUtility", "BAPEUtility", "JonesUtility", "minimizeObjective", "klNumerical",
           "utilityBatch", "scanUtility"]


def klNumerical(x, p, q):
    """Monte-Carlo KL estimate mean(log p/q) over samples x ~ p (reference utility.py:26-66)."""
    try:
        res = np.sum(np.log(p(x) / q(x))) / len(x)
    except ValueError:
        raise ValueError("ERROR: inf/NaN encountered.  q(x) = 0 likely occured.")
    return res


def logsubexp(x1, x2):
    """log(exp(x1) - exp(x2)), written as the reference writes it (utility.py:69-89)."""
    if x1 <= x2:
        return -np.inf
    else:
        return x1 + np.log(1.0 - np.exp(x2 - x1))


def _mu_var(theta, y, gp, kind):
    """(mu, var, util-or-None) of one point; util comes from the device epilogue when available."""
    if not gp.computed:
        raise RuntimeError("ERROR: Need to compute GP before using it!")
    t = np.asarray(theta, dtype=np.float64).reshape(1, -1)
    if hasattr(gp, "predict_utility"):
        mu, var, u = gp.predict_utility(y, t, kind)
        return mu[0], var[0], u[0]
    mu, var = gp.predict(y, t, return_var=True)
    return np.asarray(mu).ravel()[0], np.asarray(var).ravel()[0], None


def AGPUtility(theta, y, gp, priorFn):
    """Negative AGP utility -(mu + 0.5 ln(2 pi e var)) (reference utility.py:99-142)."""
    if not np.isfinite(priorFn(theta)):
        return np.inf
    mu, var, u = _mu_var(theta, y, gp, "agp")
    if u is None:
        u = -(mu + 0.5 * np.log(2.0 * np.pi * np.e * var))
    return u


def BAPEUtility(theta, y, gp, priorFn):
    """Negative log BAPE utility -((2 mu + var) + logsubexp(var, 0)) (reference utility.py:145-189)."""
    if not np.isfinite(priorFn(theta)):
        return np.inf
    mu, var, u = _mu_var(theta, y, gp, "bape")
    if u is None:
        u = -((2.0 * mu + var) + logsubexp(var, 0.0))
    return u


def JonesUtility(theta, y, gp, priorFn, zeta=0.01):
    """Negative expected improvement (reference utility.py:192-250)."""
    if not np.isfinite(priorFn(theta)):
        return np.inf
    if not gp.computed:
        raise RuntimeError("ERROR: Need to compute GP before using it!")
    t = np.asarray(theta, dtype=np.float64).reshape(1, -1)
    if hasattr(gp, "predict_utility"):
        return gp.predict_utility(y, t, "jones", zeta=zeta)[2][0]
    mu, var = gp.predict(y, t, return_var=True)
    mu, var = np.asarray(mu).ravel()[0], np.asarray(var).ravel()[0]
    std = np.sqrt(var)
    yBest = np.max(y)
    if std > 0:
        z = (mu - yBest - zeta) / std
    else:
        return 0.0
    pdf = np.exp(-0.5 * z * z) / np.sqrt(2.0 * np.pi)          # scipy.stats.norm.pdf's own expression: the two terms
    return -((mu - yBest - zeta) * ndtr(z) + std * pdf)         # cancel in the far tail, so the grouping shows in the bits


_KIND = {AGPUtility: "agp", BAPEUtility: "bape", JonesUtility: "jones"}
ASSUME_BOX_PRIOR = False     # set by compat.accelerate(box_prior=True): a plain-function prior is the box over `bounds`


def _kind_by_name(fn):
    """The reference's own utility functions (approxposterior.utility.{AGP,BAPE,Jones}Utility), recognised by name when
    the reference package calls this minimiser through ``compat.accelerate()``: same formulas (utility.py:99-250), so
    they are evaluated by the fused predict + utility kernel instead of one Python call per point."""
    name = getattr(fn, "__name__", "")
    mod = getattr(fn, "__module__", "") or ""
    if mod.split(".")[-1] == "utility" and name in ("AGPUtility", "BAPEUtility", "JonesUtility"):
        return {"AGPUtility": "agp", "BAPEUtility": "bape", "JonesUtility": "jones"}[name]
    return None


def utilityBatch(thetas, y, gp, priorFn, kind, bounds=None, zeta=0.01):
    """Utility of many points with one launch.  ``priorFn`` (any Python callable) gates on the host
    exactly as the scalar functions do; pass ``bounds`` instead to gate inside the kernel."""
    T = np.ascontiguousarray(np.atleast_2d(np.asarray(thetas, dtype=np.float64)))
    _, _, u = gp.predict_utility(y, T, kind, bounds=bounds, zeta=zeta)
    if priorFn is not None:
        bad = np.array([not np.isfinite(priorFn(t)) for t in T], dtype=bool)
        u = np.where(bad, np.inf, u)
    return u


def scanUtility(gp, y, kind, bounds, nCandidates=1 << 20, seed=None, zeta=0.01, device_out=False, candidates=None,
                refineRounds=0, refineN=4096, refineShrink=0.5):
    """Evaluate the utility on a cloud of ``nCandidates`` uniform draws from the box ``bounds`` on the
    device and return (thetaBest, utilBest, candidates, utilities).  This is the data-parallel
    replacement for the handful of Nelder-Mead restarts of minimizeObjective (BASELINE config 3:
    1M multistart candidates per iteration).

    ``refineRounds`` > 0 adds a device-side local search around the winner: each round scores ``refineN``
    uniform draws from a box of shrinking half-width (start: the mean candidate spacing, then
    x ``refineShrink`` per round) clipped to ``bounds`` and keeps the best -- a batched pattern search that
    replaces the sequential SciPy polish (one launch per round instead of ~80 dependent predict calls)."""
    import torch
    dev = torch.device("cuda", gp._device)
    d = gp.ndim
    g = torch.Generator(device=dev)
    if seed is not None:
        g.manual_seed(int(seed))
    lo = torch.tensor([b[0] for b in bounds], dtype=torch.float64, device=dev)
    hi = torch.tensor([b[1] for b in bounds], dtype=torch.float64, device=dev)
    if candidates is None:
        cand = lo + (hi - lo) * torch.rand((int(nCandidates), d), dtype=torch.float64, device=dev, generator=g)
    else:
        cand = candidates
    gp._sync_y(y)
    gp.recompute()
    ybest = float(np.max(y))

    def score(c):
        _, _, u = gp._predict_raw(c, True, utility=kind, bounds=bounds, ybest=ybest, zeta=zeta)
        uc = torch.where(torch.isnan(u), torch.full_like(u, float("inf")), u)
        i = int(torch.argmin(uc))
        return u, c[i].clone(), float(uc[i])

    u, best, ubest = score(cand)
    if refineRounds > 0:
        half = 0.5 * (hi - lo) * float(cand.shape[0]) ** (-1.0 / d)       # ~ mean spacing of the cloud
        for _ in range(int(refineRounds)):
            local = best + half * (2.0 * torch.rand((int(refineN), d), dtype=torch.float64, device=dev, generator=g) - 1.0)
            local = torch.minimum(torch.maximum(local, lo), hi)
            local[0] = best                                                # never lose the incumbent
            _, b2, u2 = score(local)
            if u2 <= ubest:
                best, ubest = b2, u2
            half = half * refineShrink
    best_np = best.cpu().numpy()
    if device_out:
        return best_np, ubest, cand, u
    return best_np, ubest, cand.cpu().numpy(), u.cpu().numpy()


DEVICE_OPTIMIZER = True      # module default for minimizeObjective(engine=None)


def _solve_one(f, t0, redraw, priorFn, bounds, method, options, maxIters):
    """One restart of the reference's protocol (utility.py:332-366): minimise from t0; while the optimum is not finite or
    is rejected by the prior, redraw the start and try again, at most maxIters times."""
    ii = 0
    while True:
        if ii >= maxIters:
            errMsg = "ERROR: Cannot find a valid solution. Current iterations: %d\n" % ii
            errMsg += "Maximum iterations: %d\n" % maxIters
            raise RuntimeError(errMsg)
        tmp = minimize(f, t0, bounds=bounds, method=method, options=options)["x"]
        if np.all(np.isfinite(tmp)):
            if np.isfinite(priorFn(tmp)):
                return tmp, f(tmp)
        t0 = redraw()
        ii += 1


def minimizeObjective(fn, y, gp, sampleFn, priorFn, nRestarts=5, method="nelder-mead", options=None, bounds=None,
                      theta0=None, args=None, maxIters=100, batched=True, _start=None, engine=None):
    """Multistart local minimisation of ``fn`` (reference utility.py:253-372).

    Protocol kept from the reference: Nelder-Mead ``{"adaptive": True}`` by default; bounds are only
    forwarded for methods " l-bfgs-b" (sic) and "tnc"; each restart starts from ``sampleFn(1)`` (or
    a perturbed ``theta0``), and is retried from a fresh prior sample, up to ``maxIters`` times,
    until its optimum is finite and allowed by ``priorFn``; the best of the restarts is returned.
    When ``fn`` is one of the three utilities and ``gp`` is the B200 GP, the restarts do not go through
    SciPy one objective call at a time:

    * ``engine="device"`` (default when ``priorFn`` is a box prior exposing ``.bounds`` and the method/options
      are covered): every restart is minimised by ONE CTA of ``apgp_minimize_utility`` -- SciPy's
      Nelder-Mead / Powell restated on the device, the whole multistart in a single launch;
    * ``engine="lockstep"``: the restarts advance in lock step on the host (coroutine restatements of the
      same SciPy algorithms) and each round of objective calls is one fused predict+utility launch.

    RNG consumption: the side-by-side engines draw all ``nRestarts`` start points up front and the retries afterwards.
    The reference draws start r, optimises, redraws on rejection, and only then draws start r+1 (utility.py:336,364):
    the two orders consume ``np.random`` identically as long as NO restart is retried.  The sequential path
    (``batched=False``, or an objective without a batched form) draws each start lazily and is the reference's order
    exactly, retries included (``tests/test_reference_dropin.py::test_sequential_minimize_objective_is_the_references_rng_order``).
    """
    if str(method).lower() == "nelder-mead" and options is None:
        options = {"adaptive": True}
    _box_hint = bounds                              # before the reference's rule below drops it
    if str(method).lower() in [" l-bfgs-b", "tnc"]:
        pass
    else:
        bounds = None
    if args is None:
        args = ()

    if theta0 is not None:
        theta0 = np.asarray(theta0).squeeze()
        ndim = theta0.ndim
        if ndim <= 0:
            ndim = 1

    def draw():
        return np.asarray(sampleFn(1), dtype=np.float64).ravel()

    def start_point():
        if theta0 is None:
            return draw()
        return np.atleast_1d(theta0 + np.min(theta0) * 1.0e-3 * np.random.randn(ndim)).ravel()

    kind = _KIND.get(fn) or getattr(fn, "device_kind", None) or _kind_by_name(fn)
    fn_batch = getattr(fn, "batch", None)      # any objective may bring its own batched form
    box = getattr(priorFn, "bounds", None)
    if box is None and ASSUME_BOX_PRIOR and _box_hint is not None:
        box = [(float(a), float(b)) for a, b in _box_hint]      # the caller vouches: priorFn is the box over `bounds`
    if engine is None:
        engine = "device" if DEVICE_OPTIMIZER else "lockstep"
    on_device = (engine == "device" and batched and kind is not None and box is not None and hasattr(gp, "minimize_utility")
                 and _opt.supported(method, options, bounds))
    use_batch = batched and nRestarts > 1 and (fn_batch is not None or
                                               (kind is not None and hasattr(gp, "predict_utility")))
    if not on_device and not use_batch:
        # one restart after the other, as the reference runs them: start r is drawn only after restart r - 1 (and its
        # retries) has finished, so np.random is consumed in the reference's order even when a restart is retried
        out = []
        for r in range(nRestarts):
            t0 = np.asarray(_start, dtype=np.float64).ravel() if (r == 0 and _start is not None) else start_point()
            out.append(_solve_one(lambda x: fn(x, *args), t0, draw, priorFn, bounds, method, options, maxIters))
        objective = [o[1] for o in out]
        bestInd = np.argmin(objective)
        return np.array([o[0] for o in out])[bestInd], objective[bestInd]

    starts = [start_point() for _ in range(nRestarts)]          # side-by-side restarts: all starts up front
    if _start is not None:          # engine extension: polish a known-good candidate (scanUtility)
        starts[0] = np.asarray(_start, dtype=np.float64).ravel()

    if on_device:
        # one CTA per restart, the whole multistart in one launch; restarts whose optimum is not finite or is
        # rejected by the prior are redrawn from the prior and re-run (utility.py:342-366), together
        res, objective = [None] * nRestarts, [None] * nRestarts
        todo, t0s, nlaunch, nev = list(range(nRestarts)), list(starts), 0, 0
        for ii in range(maxIters + 1):
            if ii >= maxIters:
                raise RuntimeError("ERROR: Cannot find a valid solution. Current iterations: %d\n"
                                   "Maximum iterations: %d\n" % (ii, maxIters))
            xs, fs, nfev = gp.minimize_utility(y, np.array(t0s), kind, bounds=box, method=method, options=options)
            nlaunch += 1
            nev += int(np.sum(nfev))
            again, t0n = [], []
            for r, x, f in zip(todo, xs, fs):
                if np.all(np.isfinite(x)) and np.isfinite(priorFn(x)):
                    res[r], objective[r] = x, f
                else:
                    again.append(r)
                    t0n.append(draw())
            todo, t0s = again, t0n
            if not todo:
                break
        minimizeObjective.last_stats = dict(batches=nlaunch, evals=nev, scheduler="device")
        bestInd = np.argmin(objective)
        return np.array(res)[bestInd], objective[bestInd]

    def solve(f, t0, redraw):
        return _solve_one(f, t0, redraw, priorFn, bounds, method, options, maxIters)

    # lock-step restarts on the host: every round of objective calls is one fused predict + utility launch
    zeta = 0.01

    def batch_fn(thetas):
        if fn_batch is not None:
            return fn_batch(np.array(thetas))
        return utilityBatch(np.array(thetas), y, gp, priorFn, kind, zeta=zeta)

    if _opt.supported(method, options, bounds):
        # thread-free lock step: SciPy's algorithm restated as a coroutine per restart
        make = _opt.nelder_mead_gen if str(method).lower() == "nelder-mead" else _opt.powell_gen

        def restart(t0):
            ii = 0
            while True:
                if ii >= maxIters:
                    raise RuntimeError("ERROR: Cannot find a valid solution. Current iterations: %d\n"
                                       "Maximum iterations: %d\n" % (ii, maxIters))
                tmp, _ = yield from make(t0, **(options or {}))
                if np.all(np.isfinite(tmp)) and np.isfinite(priorFn(tmp)):
                    ftmp = yield np.copy(tmp)            # the reference re-evaluates fn at the optimum
                    return tmp, ftmp
                t0 = draw()
                ii += 1

        out, rounds, evals = _opt.run_generators([restart(t0) for t0 in starts], batch_fn)
        minimizeObjective.last_stats = dict(batches=rounds, evals=evals, scheduler="generators")
    else:
        import threading
        rng_lock = threading.Lock()

        def redraw():
            with rng_lock:
                return draw()

        out, ev = run_lockstep(nRestarts, batch_fn, lambda wid, f: solve(f, starts[wid], redraw))
        minimizeObjective.last_stats = dict(batches=ev.nbatches, evals=ev.nevals, scheduler="threads")

    res = [o[0] for o in out]
    objective = [o[1] for o in out]
    bestInd = np.argmin(objective)
    return np.array(res)[bestInd], objective[bestInd]


minimizeObjective.last_stats = None
