"""Restatement of emcee 3.0.x EnsembleSampler + StretchMove(a=2) (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED by the reference (see ``oracle/__init__.py``): emcee is an
un-vendored dependency (reference ``setup.py:68``) driven from reference
``approx.py:839-847``.  Restated from emcee 3.0.x's published algorithm
(Goodman & Weare 2010 stretch move, red/blue split):

  per iteration
    move  = random.choice(moves, p=weights)     -> consumes one uniform
    inds  = arange(nw) % 2 ; random.shuffle(inds)
    for split in (0, 1):
      s = coords[inds == split] ; c = coords[inds != split]
      zz      = ((a-1) * random.rand(Ns) + 1)**2 / a
      factors = (ndim-1) * log(zz)
      rint    = random.randint(Nc, size=Ns)
      q       = c[rint] - (c[rint] - s) * zz[:, None]
      nlp     = log_prob(q)
      accept_j = factors_j + nlp_j - lp_j > log(random.rand())   (one draw per walker, in order)

The function can *record* every draw so the CUDA sampler replays them.
"""
import numpy as np


def stretch_move_oracle(log_prob_fn, p0, nsteps, a=2.0, rng=None, record=False):
    """Run one ensemble.  ``log_prob_fn(q[Ns,d]) -> (lp[Ns], blob[Ns])`` is vectorised.

    Returns dict(chain[nsteps,nw,d], log_prob[nsteps,nw], blobs[nsteps,nw],
    accepted[nw] counts) and, if ``record``, the replay buffers
    inds[nsteps,nw] (0/1 colour), zz[nsteps,2,Ns], rint[nsteps,2,Ns], logu[nsteps,2,Ns].
    """
    rng = np.random.RandomState(0) if rng is None else rng
    coords = np.array(p0, dtype=np.float64, copy=True)
    nw, nd = coords.shape
    assert nw % 2 == 0 and nw >= 2
    Ns = nw // 2
    lp, blob = log_prob_fn(coords)
    lp = np.array(lp, dtype=np.float64)
    blob = np.array(blob, dtype=np.float64)
    chain = np.empty((nsteps, nw, nd))
    lps = np.empty((nsteps, nw))
    blobs = np.empty((nsteps, nw))
    nacc = np.zeros(nw, dtype=np.int64)
    if record:
        r_inds = np.empty((nsteps, nw), dtype=np.int32)
        r_zz = np.empty((nsteps, 2, Ns))
        r_rint = np.empty((nsteps, 2, Ns), dtype=np.int32)
        r_logu = np.empty((nsteps, 2, Ns))
    all_inds = np.arange(nw)
    for it in range(nsteps):
        rng.choice(1, p=[1.0])                       # move selection draw
        inds = all_inds % 2
        rng.shuffle(inds)
        if record:
            r_inds[it] = inds
        for split in (0, 1):
            S1 = inds == split
            s = coords[S1]
            c = coords[~S1]
            zz = ((a - 1.0) * rng.rand(Ns) + 1.0) ** 2.0 / a
            factors = (nd - 1.0) * np.log(zz)
            rint = rng.randint(len(c), size=(Ns,))
            q = c[rint] - (c[rint] - s) * zz[:, None]
            nlp, nblob = log_prob_fn(q)
            logu = np.empty(Ns)
            acc = np.zeros(Ns, dtype=bool)
            for i, j in enumerate(all_inds[S1]):
                logu[i] = np.log(rng.rand())
                acc[i] = (factors[i] + nlp[i] - lp[j]) > logu[i]
            jj = all_inds[S1][acc]
            coords[jj] = q[acc]
            lp[jj] = np.asarray(nlp)[acc]
            blob[jj] = np.asarray(nblob)[acc]
            nacc[jj] += 1
            if record:
                r_zz[it, split] = zz
                r_rint[it, split] = rint
                r_logu[it, split] = logu
        chain[it] = coords
        lps[it] = lp
        blobs[it] = blob
    out = dict(chain=chain, log_prob=lps, blobs=blobs, naccepted=nacc)
    if record:
        out.update(inds=r_inds, zz=r_zz, rint=r_rint, logu=r_logu)
    return out


def gpll_batch(gp, y, q, lo, hi, lnprior_const=0.0):
    """Vectorised restatement of reference ApproxPosterior._gpll (approx.py:148-189)
    with a box prior (reference likelihood.py:60-63 style): returns (mu or -inf, lnprior or nan)."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    ok = np.all(np.isfinite(q), axis=1) & np.all((q >= lo) & (q <= hi), axis=1)
    lp = np.full(q.shape[0], -np.inf)
    blob = np.full(q.shape[0], np.nan)
    if np.any(ok):
        mu = gp.predict(y, q[ok], return_cov=False, return_var=False)
        fin = np.isfinite(mu)
        idx = np.nonzero(ok)[0]
        lp[idx[fin]] = mu[fin]
        blob[idx[fin]] = lnprior_const
    return lp, blob
