"""Multi-GPU sharding for the three embarrassingly parallel workloads.

One process per GPU (``torchrun``), GP state replicated by refactorising on every rank (cheaper than a
broadcast at N <= 2048), independent units sharded as contiguous blocks, and exactly ONE collective
per call: an all-gather of candidate scores / chains / restart results.  ``torch.distributed`` is the
plumbing (NCCL on GPUs, gloo in the CPU tests); there is no per-step communication.
"""
import numpy as np

__all__ = ["world", "shard_bounds", "gather_best", "gather_concat", "scan_utility_sharded",
           "run_ensembles_sharded", "best_restart_sharded", "predict_sharded", "optimize_gp_sharded"]


def world():
    """(rank, world_size) -- (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous block [lo, hi) of ``n`` units owned by ``rank``; blocks differ by at most one unit."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device_for_collective():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def gather_concat(local, axis=0):
    """All-gather equally-shaped NumPy arrays and concatenate along ``axis`` (one collective)."""
    rank, ws = world()
    local = np.ascontiguousarray(local)
    if ws == 1:
        return local
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(local).to(_device_for_collective())
    out = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy() for o in out], axis=axis)


def gather_best(score, theta):
    """Each rank contributes its best (score, theta); every rank gets the global arg-min.
    NaN scores never win.  One all-gather of 1+d doubles per rank."""
    theta = np.asarray(theta, dtype=np.float64).ravel()
    s = float(score)
    pack = np.concatenate([[s if np.isfinite(s) or s == -np.inf else np.inf], theta])
    allp = gather_concat(pack[None, :], axis=0)
    i = int(np.argmin(allp[:, 0]))
    return allp[i, 1:], float(allp[i, 0]), i


def scan_utility_sharded(gp, y, kind, bounds, nCandidates, seed=0, zeta=0.01, candidates=None):
    """Utility scan of ``nCandidates`` uniform candidates split over the ranks (BASELINE config 3):
    rank r draws and scores its own block on its GPU, then one all-gather picks the winner.
    ``candidates``: this rank's block as a torch CUDA tensor, when the caller already holds it on the device."""
    from .utility import scanUtility
    rank, ws = world()
    lo, hi = shard_bounds(nCandidates, rank, ws)
    best, ubest, _, _ = scanUtility(gp, y, kind, bounds, nCandidates=hi - lo, seed=int(seed) * 1000003 + rank,
                                    zeta=zeta, device_out=True, candidates=candidates)
    return gather_best(ubest, best)[:2]


def run_ensembles_sharded(gp, y, p0, nsteps, bounds, nens, **kw):
    """Independent ensembles split over the ranks (never one ensemble across GPUs: the stretch move
    couples each walker to the complementary half every half-step).  Returns the gathered
    (chain, log_prob, blobs, naccepted) with walkers ordered by ensemble."""
    rank, ws = world()
    p0 = np.asarray(p0, dtype=np.float64)
    nw = p0.shape[0] // nens
    if nens % ws:
        raise ValueError("nens must be a multiple of the world size (equal shards keep one all-gather)")
    lo, hi = shard_bounds(nens, rank, ws)
    seed = kw.pop("seed", 0)
    if ws > 1 and _device_for_collective().type == "cuda" and hasattr(gp, "_device") and kw.get("replay") is None:
        # NCCL: the shards stay on the GPUs, meet over NVLink and cross PCIe once, already in walker order -- instead of
        # D2H of the shard, H2D for the collective, D2H of every rank's piece and a strided host concatenation
        kw.pop("device_out", None)
        import os, sys, time, torch
        timing = bool(os.environ.get("APGP_DIST_TIMING"))       # per-phase wall times of rank 0 on stderr (adds two syncs)
        t0 = time.perf_counter()
        out = gp.run_ensembles(y, p0[lo * nw:hi * nw], nsteps, bounds, nens=hi - lo, seed=int(seed) + 7919 * rank,
                               device_out=True, **kw)
        if timing:
            torch.cuda.synchronize(); t1 = time.perf_counter()
        dev = {k: _gather_concat_device(out[k], ax)
               for k, ax in (("chain", 1), ("log_prob", 1), ("blobs", 1), ("naccepted", 0))}
        if timing:
            torch.cuda.synchronize(); t2 = time.perf_counter()
        res = _to_host(dev)
        if timing and rank == 0:
            print("run_ensembles_sharded: sample %.1f ms, gather %.1f ms, to host %.1f ms"
                  % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (time.perf_counter() - t2) * 1e3), file=sys.stderr)
        return res
    out = gp.run_ensembles(y, p0[lo * nw:hi * nw], nsteps, bounds, nens=hi - lo, seed=int(seed) + 7919 * rank, **kw)
    return dict(chain=gather_concat(out["chain"], axis=1), log_prob=gather_concat(out["log_prob"], axis=1),
                blobs=gather_concat(out["blobs"], axis=1), naccepted=gather_concat(out["naccepted"], axis=0))


def _gather_concat_device(t, axis):
    """``gather_concat`` for a CUDA tensor, result left on the device: one all-gather into a [world, ...] tensor and the
    concatenation along ``axis``."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size()
    t = t.contiguous()
    buf = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)   # rank-major concatenation
    dist.all_gather_into_tensor(buf, t)
    if axis == 0:
        return buf
    return torch.cat(list(buf.reshape((ws,) + tuple(t.shape)).unbind(0)), dim=axis)


def _to_host(tensors):
    """Device tensors -> NumPy arrays through pinned host memory from torch's caching host allocator: the copies run at
    PCIe speed instead of the pageable path's staging + page faults (105 MB: ~2 ms instead of ~50 ms), and a loop that
    drops the previous result gets its blocks back without a new cudaHostAlloc.  The arrays keep their tensors alive.
    Falls back to pageable copies when the pinned allocation is refused."""
    import torch
    host = {}
    try:
        for k, t in tensors.items():
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t, non_blocking=True)
            host[k] = h
        for t in tensors.values():
            torch.cuda.current_stream(t.device).synchronize()
            break
    except RuntimeError:
        host = {k: t.cpu() for k, t in tensors.items()}
    return {k: h.numpy() for k, h in host.items()}


def best_restart_sharded(params, mll):
    """optimizeGP restarts split over ranks: all-gather (mll, p) rows, arg-max on every rank."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    mll = np.asarray(mll, dtype=np.float64).reshape(-1, 1)
    allr = gather_concat(np.hstack([mll, params]), axis=0)
    m = np.where(np.isfinite(allr[:, 0]), allr[:, 0], -np.inf)
    i = int(np.argmax(m))
    return allr[i, 1:], float(allr[i, 0])


def predict_sharded(gp, y, n_queries_total, make_queries, chunk=1 << 22, want_var=True, utility=None, bounds=None):
    """BASELINE config 5: ``n_queries_total`` query points split over the ranks as contiguous blocks; rank r
    produces its block chunk by chunk with ``make_queries(first_index, count) -> torch CUDA tensor [count, d]``
    (e.g. a counter-based generator, so the points do not depend on the number of ranks), evaluates the fused predict
    kernel on each chunk and keeps running sums; ONE all-gather of (count, sum mu, sum var, min utility) per call
    closes it.  Returns dict(count, sum_mu, sum_var, min_util) over ALL ranks."""
    import torch
    rank, ws = world()
    lo, hi = shard_bounds(n_queries_total, rank, ws)
    gp._sync_y(y)
    gp.recompute()
    ybest = float(np.max(gp._y))
    acc = torch.zeros(3, dtype=torch.float64, device=torch.device("cuda", gp._device))
    umin = float("inf")
    for s in range(lo, hi, chunk):
        c = min(chunk, hi - s)
        q = make_queries(s, c)
        mu, var, u = gp._predict_raw(q, want_var, utility=utility, bounds=bounds, ybest=ybest)
        acc[0] += c
        acc[1] += mu.sum()
        if var is not None:
            acc[2] += var.sum()
        if u is not None:
            umin = min(umin, float(torch.nan_to_num(u, nan=float("inf")).min()))
    pack = np.concatenate([acc.cpu().numpy(), [umin]])
    allp = gather_concat(pack[None, :], axis=0)
    return dict(count=int(allp[:, 0].sum()), sum_mu=float(allp[:, 1].sum()), sum_var=float(allp[:, 2].sum()),
                min_util=float(allp[:, 3].min()))


def optimize_gp_sharded(gp, y, x0s, method="powell", options=None, default_prior=True):
    """gpUtils.optimizeGP's restarts (reference gpUtils.py:223-254) split over the ranks: every rank holds the same
    ``x0s`` [R, P] (same seed), minimises its contiguous block on its GPU with ONE ``apgp_minimize_nll`` launch, and
    one all-gather of (mll, p) rows picks the best restart on every rank.  Returns (p_best, mll_best, nfev_total)."""
    rank, ws = world()
    x0s = np.atleast_2d(np.asarray(x0s, dtype=np.float64))
    lo, hi = shard_bounds(x0s.shape[0], rank, ws)
    P = x0s.shape[1]
    if hi > lo:
        p, f, nfev = gp.minimize_nll(x0s[lo:hi], y, method=method, options=options, default_prior=default_prior)
        rows = np.hstack([np.where(np.isfinite(f), -f, -np.inf)[:, None], p, nfev[:, None].astype(np.float64)])
    else:
        rows = np.empty((0, P + 2))
    # equal-shape gather: pad every rank's block to the largest block
    width = -(-x0s.shape[0] // ws)
    pad = np.full((width, P + 2), -np.inf)
    pad[:, -1] = 0.0
    pad[:rows.shape[0]] = rows
    allr = gather_concat(pad, axis=0)
    m = np.where(np.isfinite(allr[:, 0]), allr[:, 0], -np.inf)
    i = int(np.argmax(m))
    return allr[i, 1:1 + P], float(allr[i, 0]), int(allr[:, -1].sum())
