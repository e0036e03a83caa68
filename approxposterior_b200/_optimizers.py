"""Coroutine (generator) forms of SciPy's Nelder-Mead and Powell minimisers.

``scipy.optimize.minimize`` owns the control flow, so batching the objective calls of several restarts
needs one thread per restart (``_lockstep.py``) and every round pays a Python thread hand-off per
restart (~50 us each, measured).  For the two methods the reference actually uses -- Nelder-Mead
``{"adaptive": True}`` in ``utility.minimizeObjective`` (utility.py:307-308) and Powell in
``gpUtils.optimizeGP`` (gpUtils.py:184) -- the algorithms are restated here as generators that *yield*
the next point to evaluate and are *sent* its value, so a plain loop can collect one point from every
live restart, evaluate them in ONE device launch and resume them, with no threads at all.

The restatements follow SciPy 1.18's ``_minimize_neldermead`` / ``_minimize_powell`` (+ ``bracket`` and
Brent's line search) statement for statement, in the same floating-point order, so each restart visits
exactly the points SciPy would visit (``tests/test_host_logic.py::test_generator_optimisers_match_scipy``
checks the evaluated-point sequences for equality).  Options outside the subset below fall back to the
threaded SciPy path.
"""

# ---------------------------------------------------------------------------------------------------------------------
# The optimiser bodies below (Nelder-Mead, Powell, bracket, Brent) restate, statement for statement, the algorithms of
# SciPy 1.18's scipy/optimize/_optimize.py so that iterates match SciPy's bit for bit.  SciPy is distributed under the
# BSD 3-Clause licence:
#
#   Copyright (c) 2001-2002 Enthought, Inc. 2003, SciPy Developers.  All rights reserved.
#
#   Redistribution and use in source and binary forms, with or without modification, are permitted provided that the
#   following conditions are met:
#   1. Redistributions of source code must retain the above copyright notice, this list of conditions and the following
#      disclaimer.
#   2. Redistributions in binary form must reproduce the above copyright notice, this list of conditions and the
#      following disclaimer in the documentation and/or other materials provided with the distribution.
#   3. Neither the name of the copyright holder nor the names of its contributors may be used to endorse or promote
#      products derived from this software without specific prior written permission.
#
#   THIS SOFTWARE IS PROVIDED BY THE COPYRIGHT HOLDERS AND CONTRIBUTORS "AS IS" AND ANY EXPRESS OR IMPLIED WARRANTIES,
#   INCLUDING, BUT NOT LIMITED TO, THE IMPLIED WARRANTIES OF MERCHANTABILITY AND FITNESS FOR A PARTICULAR PURPOSE ARE
#   DISCLAIMED.  IN NO EVENT SHALL THE COPYRIGHT HOLDER OR CONTRIBUTORS BE LIABLE FOR ANY DIRECT, INDIRECT, INCIDENTAL,
#   SPECIAL, EXEMPLARY, OR CONSEQUENTIAL DAMAGES (INCLUDING, BUT NOT LIMITED TO, PROCUREMENT OF SUBSTITUTE GOODS OR
#   SERVICES; LOSS OF USE, DATA, OR PROFITS; OR BUSINESS INTERRUPTION) HOWEVER CAUSED AND ON ANY THEORY OF LIABILITY,
#   WHETHER IN CONTRACT, STRICT LIABILITY, OR TORT (INCLUDING NEGLIGENCE OR OTHERWISE) ARISING IN ANY WAY OUT OF THE USE
#   OF THIS SOFTWARE, EVEN IF ADVISED OF THE POSSIBILITY OF SUCH DAMAGE.
# ---------------------------------------------------------------------------------------------------------------------

import numpy as np

__all__ = ["nelder_mead_gen", "powell_gen", "run_generators", "supported"]


class _MaxFun(Exception):
    pass


def supported(method, options, bounds=None, jac=None):
    """True when (method, options) is covered by the generator restatements."""
    if bounds is not None or jac not in (None, False):
        return False
    m = str(method).lower()
    opts = dict(options or {})
    if m == "nelder-mead":
        return set(opts) <= {"adaptive", "xatol", "fatol", "maxiter", "maxfev"}
    if m == "powell":
        return set(opts) <= {"xtol", "ftol", "maxiter", "maxfev"}
    return False


# ----------------------------------------------------------------------------------------------
# Nelder-Mead (scipy.optimize._optimize._minimize_neldermead)
# ----------------------------------------------------------------------------------------------
def nelder_mead_gen(x0, adaptive=False, xatol=1e-4, fatol=1e-4, maxiter=None, maxfev=None, _stable=False):
    """``_stable`` orders tied simplex values with a stable sort, as the device restatement (csrc/optimize.cu)
    does; SciPy's plain ``np.argsort`` leaves the order of ties to NumPy's introsort / SIMD dispatch."""
    argsort = (lambda a: np.argsort(a, kind="stable")) if _stable else np.argsort
    x0 = np.atleast_1d(np.asarray(x0, dtype=np.float64)).flatten()
    N = len(x0)
    if adaptive:
        dim = float(N)
        rho, chi, psi, sigma = 1, 1 + 2 / dim, 0.75 - 1 / (2 * dim), 1 - 1 / dim
    else:
        rho, chi, psi, sigma = 1, 2, 0.5, 0.5
    nonzdelt, zdelt = 0.05, 0.00025
    sim = np.empty((N + 1, N), dtype=x0.dtype)
    sim[0] = x0
    for k in range(N):
        y = np.array(x0, copy=True)
        if y[k] != 0:
            y[k] = (1 + nonzdelt) * y[k]
        else:
            y[k] = zdelt
        sim[k + 1] = y
    if maxiter is None and maxfev is None:
        maxiter, maxfun = N * 200, N * 200
    elif maxiter is None:
        maxfun = maxfev
        maxiter = N * 200 if maxfev == np.inf else np.inf
    elif maxfev is None:
        maxfun = N * 200 if maxiter == np.inf else np.inf
    else:
        maxfun = maxfev
    ncalls = [0]

    def ev(x):                       # generator helper: counts calls like SciPy's maxfun wrapper
        if ncalls[0] >= maxfun:
            raise _MaxFun()
        ncalls[0] += 1
        fx = yield np.copy(x)
        return fx

    one2np1 = list(range(1, N + 1))
    fsim = np.full((N + 1,), np.inf, dtype=float)
    try:
        for k in range(N + 1):
            fsim[k] = yield from ev(sim[k])
    except _MaxFun:
        pass
    ind = argsort(fsim)
    sim = np.take(sim, ind, 0)
    fsim = np.take(fsim, ind, 0)
    iterations = 1
    while ncalls[0] < maxfun and iterations < maxiter:
        try:
            if (np.max(np.ravel(np.abs(sim[1:] - sim[0]))) <= xatol and
                    np.max(np.abs(fsim[0] - fsim[1:])) <= fatol):
                break
            xbar = np.add.reduce(sim[:-1], 0) / N
            xr = (1 + rho) * xbar - rho * sim[-1]
            fxr = yield from ev(xr)
            doshrink = 0
            if fxr < fsim[0]:
                xe = (1 + rho * chi) * xbar - rho * chi * sim[-1]
                fxe = yield from ev(xe)
                if fxe < fxr:
                    sim[-1] = xe
                    fsim[-1] = fxe
                else:
                    sim[-1] = xr
                    fsim[-1] = fxr
            else:
                if fxr < fsim[-2]:
                    sim[-1] = xr
                    fsim[-1] = fxr
                else:
                    if fxr < fsim[-1]:
                        xc = (1 + psi * rho) * xbar - psi * rho * sim[-1]
                        fxc = yield from ev(xc)
                        if fxc <= fxr:
                            sim[-1] = xc
                            fsim[-1] = fxc
                        else:
                            doshrink = 1
                    else:
                        xcc = (1 - psi) * xbar + psi * sim[-1]
                        fxcc = yield from ev(xcc)
                        if fxcc < fsim[-1]:
                            sim[-1] = xcc
                            fsim[-1] = fxcc
                        else:
                            doshrink = 1
                    if doshrink:
                        for j in one2np1:
                            sim[j] = sim[0] + sigma * (sim[j] - sim[0])
                            fsim[j] = yield from ev(sim[j])
            iterations += 1
        except _MaxFun:
            pass
        finally:
            ind = argsort(fsim)
            sim = np.take(sim, ind, 0)
            fsim = np.take(fsim, ind, 0)
    return sim[0], fsim[0]


# ----------------------------------------------------------------------------------------------
# Powell (scipy.optimize._optimize._minimize_powell, unbounded) with bracket() + Brent line search
# ----------------------------------------------------------------------------------------------
def _bracket_gen(f, xa=0.0, xb=1.0, grow_limit=110.0, maxiter=1000):
    """scipy.optimize.bracket; returns (xa, xb, xc, fa, fb, fc, valid)."""
    _gold = 1.618034
    _verysmall_num = 1e-21
    xa, xb = np.asarray([xa, xb])
    fa = yield from f(xa)
    fb = yield from f(xb)
    if fa < fb:
        xa, xb = xb, xa
        fa, fb = fb, fa
    xc = xb + _gold * (xb - xa)
    fc = yield from f(xc)
    it = 0
    while fc < fb:
        tmp1 = (xb - xa) * (fb - fc)
        tmp2 = (xb - xc) * (fb - fa)
        val = tmp2 - tmp1
        if np.abs(val) < _verysmall_num:
            denom = 2.0 * _verysmall_num
        else:
            denom = 2.0 * val
        w = xb - ((xb - xc) * tmp2 - (xb - xa) * tmp1) / denom
        wlim = xb + grow_limit * (xc - xb)
        if it > maxiter:
            raise RuntimeError("No valid bracket was found before the iteration limit was reached.")
        it += 1
        if (w - xc) * (xb - w) > 0.0:
            fw = yield from f(w)
            if fw < fc:
                xa = xb
                xb = w
                fa = fb
                fb = fw
                break
            elif fw > fb:
                xc = w
                fc = fw
                break
            w = xc + _gold * (xc - xb)
            fw = yield from f(w)
        elif (w - wlim) * (wlim - xc) >= 0.0:
            w = wlim
            fw = yield from f(w)
        elif (w - wlim) * (xc - w) > 0.0:
            fw = yield from f(w)
            if fw < fc:
                xb = xc
                xc = w
                w = xc + _gold * (xc - xb)
                fb = fc
                fc = fw
                fw = yield from f(w)
        else:
            w = xc + _gold * (xc - xb)
            fw = yield from f(w)
        xa = xb
        xb = xc
        xc = w
        fa = fb
        fb = fc
        fc = fw
    cond1 = (fb < fc and fb <= fa) or (fb < fa and fb <= fc)
    cond2 = (xa < xb < xc or xc < xb < xa)
    cond3 = np.isfinite(xa) and np.isfinite(xb) and np.isfinite(xc)
    return xa, xb, xc, fa, fb, fc, bool(cond1 and cond2 and cond3)


def _brent_gen(f, tol, maxiter=500):
    """Brent.optimize after bracket(), wrapped as _recover_from_bracket_error does; returns (xmin, fval)."""
    xa, xb, xc, fa, fb, fc, valid = yield from _bracket_gen(f)
    if not valid:
        xs, fs = [xa, xb, xc], [fa, fb, fc]
        if np.any(np.isnan([xs, fs])):
            return np.nan, np.nan
        imin = np.argmin(fs)
        return xs[imin], fs[imin]
    _mintol = 1.0e-11
    _cg = 0.3819660
    x = w = v = xb
    fw = fv = fx = fb
    if xa < xc:
        a = xa
        b = xc
    else:
        a = xc
        b = xa
    deltax = 0.0
    it = 0
    rat = None
    while it < maxiter:
        tol1 = tol * np.abs(x) + _mintol
        tol2 = 2.0 * tol1
        xmid = 0.5 * (a + b)
        if np.abs(x - xmid) < (tol2 - 0.5 * (b - a)):
            break
        if np.abs(deltax) <= tol1:
            if x >= xmid:
                deltax = a - x
            else:
                deltax = b - x
            rat = _cg * deltax
        else:
            tmp1 = (x - w) * (fx - fv)
            tmp2 = (x - v) * (fx - fw)
            p = (x - v) * tmp2 - (x - w) * tmp1
            tmp2 = 2.0 * (tmp2 - tmp1)
            if tmp2 > 0.0:
                p = -p
            tmp2 = np.abs(tmp2)
            dx_temp = deltax
            deltax = rat
            if ((p > tmp2 * (a - x)) and (p < tmp2 * (b - x)) and
                    (np.abs(p) < np.abs(0.5 * tmp2 * dx_temp))):
                rat = p * 1.0 / tmp2
                u = x + rat
                if (u - a) < tol2 or (b - u) < tol2:
                    if xmid - x >= 0:
                        rat = tol1
                    else:
                        rat = -tol1
            else:
                if x >= xmid:
                    deltax = a - x
                else:
                    deltax = b - x
                rat = _cg * deltax
        if np.abs(rat) < tol1:
            if rat >= 0:
                u = x + tol1
            else:
                u = x - tol1
        else:
            u = x + rat
        fu = yield from f(u)
        if fu > fx:
            if u < x:
                a = u
            else:
                b = u
            if (fu <= fw) or (w == x):
                v = w
                w = u
                fv = fw
                fw = fu
            elif (fu <= fv) or (v == x) or (v == w):
                v = u
                fv = fu
        else:
            if u >= x:
                a = x
            else:
                b = x
            v = w
            w = x
            x = u
            fv = fw
            fw = fx
            fx = fu
        it += 1
    return x, fx


def powell_gen(x0, xtol=1e-4, ftol=1e-4, maxiter=None, maxfev=None):
    x = np.asarray(x0, dtype=np.float64).flatten()
    N = len(x)
    maxfun = maxfev
    if maxiter is None and maxfun is None:
        maxiter, maxfun = N * 1000, N * 1000
    elif maxiter is None:
        maxiter = N * 1000 if maxfun == np.inf else np.inf
    elif maxfun is None:
        maxfun = N * 1000 if maxiter == np.inf else np.inf
    ncalls = [0]

    def ev(xx):
        if ncalls[0] >= maxfun:
            raise _MaxFun()
        ncalls[0] += 1
        fx = yield np.copy(xx)
        return fx

    def linesearch(p, xi, tol, fval):
        # _linesearch_powell, unbounded branch
        if not np.any(xi):
            return fval, p, xi

        def myfunc(alpha):
            return (yield from ev(p + alpha * xi))

        alpha_min, fret = yield from _brent_gen(myfunc, tol)
        xi = alpha_min * xi
        return fret, p + xi, xi

    direc = np.eye(N, dtype=float)
    fval = yield from ev(x)
    x1 = x.copy()
    it = 0
    while True:
        try:
            fx = fval
            bigind = 0
            delta = 0.0
            for i in range(N):
                direc1 = direc[i]
                fx2 = fval
                fval, x, direc1 = yield from linesearch(x, direc1, xtol * 100, fval)
                if (fx2 - fval) > delta:
                    delta = fx2 - fval
                    bigind = i
            it += 1
            bnd = ftol * (np.abs(fx) + np.abs(fval)) + 1e-20
            if 2.0 * (fx - fval) <= bnd:
                break
            if ncalls[0] >= maxfun:
                break
            if it >= maxiter:
                break
            if np.isnan(fx) and np.isnan(fval):
                break
            direc1 = x - x1
            x1 = x.copy()
            x2 = x + direc1
            fx2 = yield from ev(x2)
            if fx > fx2:
                t = 2.0 * (fx + fx2 - 2.0 * fval)
                temp = (fx - fval - delta)
                t *= temp * temp
                temp = fx - fx2
                t -= delta * temp * temp
                if t < 0.0:
                    fval, x, direc1 = yield from linesearch(x, direc1, xtol * 100, fval)
                    if np.any(direc1):
                        direc[bigind] = direc[-1]
                        direc[-1] = direc1
        except _MaxFun:
            break
    return x, fval


# ----------------------------------------------------------------------------------------------
def run_generators(gens, batch_fn):
    """Drive optimiser generators in lock step: each round gathers one pending point from every live
    generator, evaluates them with ONE call ``batch_fn(list_of_points) -> values`` and resumes them.
    Returns ([(x, f) per generator], rounds, evals)."""
    n = len(gens)
    results = [None] * n
    pending = {}
    for i, g in enumerate(gens):
        try:
            pending[i] = next(g)
        except StopIteration as s:
            results[i] = s.value
    rounds = evals = 0
    while pending:
        ids = sorted(pending)
        vals = batch_fn([pending[i] for i in ids])
        rounds += 1
        evals += len(ids)
        nxt = {}
        for i, v in zip(ids, vals):
            try:
                nxt[i] = gens[i].send(np.float64(v))
            except StopIteration as s:
                results[i] = s.value
        pending = nxt
    return results, rounds, evals
