"""Multi-GPU sharding for the three embarrassingly parallel workloads.

One process per GPU (``torchrun``), GP state replicated by refactorising on every rank (cheaper than a
broadcast at N <= 2048), independent units sharded as contiguous blocks, and exactly ONE collective
per call: an all-gather of candidate scores / chains / restart results.  ``torch.distributed`` is the
plumbing (NCCL on GPUs, gloo in the CPU tests); there is no per-step communication.
"""
import numpy as np

__all__ = ["world", "shard_bounds", "gather_best", "gather_concat", "scan_utility_sharded",
           "run_ensembles_sharded", "best_restart_sharded"]


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


def scan_utility_sharded(gp, y, kind, bounds, nCandidates, seed=0, zeta=0.01):
    """Utility scan of ``nCandidates`` uniform candidates split over the ranks (BASELINE config 3):
    rank r draws and scores its own block on its GPU, then one all-gather picks the winner."""
    from .utility import scanUtility
    rank, ws = world()
    lo, hi = shard_bounds(nCandidates, rank, ws)
    best, ubest, _, _ = scanUtility(gp, y, kind, bounds, nCandidates=hi - lo, seed=int(seed) * 1000003 + rank,
                                    zeta=zeta, device_out=True)
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
    out = gp.run_ensembles(y, p0[lo * nw:hi * nw], nsteps, bounds, nens=hi - lo, seed=int(seed) + 7919 * rank, **kw)
    return dict(chain=gather_concat(out["chain"], axis=1), log_prob=gather_concat(out["log_prob"], axis=1),
                blobs=gather_concat(out["blobs"], axis=1), naccepted=gather_concat(out["naccepted"], axis=0))


def best_restart_sharded(params, mll):
    """optimizeGP restarts split over ranks: all-gather (mll, p) rows, arg-max on every rank."""
    params = np.atleast_2d(np.asarray(params, dtype=np.float64))
    mll = np.asarray(mll, dtype=np.float64).reshape(-1, 1)
    allr = gather_concat(np.hstack([mll, params]), axis=0)
    m = np.where(np.isfinite(allr[:, 0]), allr[:, 0], -np.inf)
    i = int(np.argmax(m))
    return allr[i, 1:], float(allr[i, 0])
