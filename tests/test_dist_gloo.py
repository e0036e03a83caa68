"""world_size-2 gloo tests (CPU) of the sharding helpers used on the N>1 GPU path."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from approxposterior_b200 import dist as apd
        assert apd.world() == (rank, ws)
        # candidate scan: every rank scores its own block of one global candidate list
        rng = np.random.default_rng(5)
        cand = rng.uniform(-5, 5, size=(1001, 3))
        score = np.sum((cand - 1.0) ** 2, axis=1)
        score[17] = np.nan
        lo, hi = apd.shard_bounds(len(cand), rank, ws)
        s_loc = np.where(np.isnan(score[lo:hi]), np.inf, score[lo:hi])
        i_loc = int(np.argmin(s_loc))
        theta, best, owner = apd.gather_best(s_loc[i_loc], cand[lo + i_loc])
        i_glob = int(np.nanargmin(score))
        ok1 = np.array_equal(theta, cand[i_glob]) and best == score[i_glob]
        # chains: walkers concatenate in rank order
        chain_loc = np.full((4, 6, 2), float(rank))
        full = apd.gather_concat(chain_loc, axis=1)
        ok2 = full.shape == (4, 6 * ws, 2) and all(np.all(full[:, 6 * r:6 * (r + 1)] == r) for r in range(ws))
        # restarts: arg-max of the marginal likelihood over all ranks, -inf/NaN never win
        p_loc = np.array([[rank, 1.0], [rank, 2.0]])
        mll = np.array([-10.0 + rank, np.nan if rank == 0 else -np.inf])
        pbest, mbest = apd.best_restart_sharded(p_loc, mll)
        ok3 = mbest == -10.0 + (ws - 1) and pbest[0] == ws - 1 and pbest[1] == 1.0
        # optimizeGP restarts sharded over ranks (uneven split: 5 restarts on 2 ranks), fake engine on CPU
        class FakeGP(object):
            def minimize_nll(self, x0s, y, method="powell", options=None, default_prior=True):
                x0s = np.atleast_2d(x0s)
                f = np.sum((x0s - 0.3) ** 2, axis=1)
                f[np.any(x0s > 1.9, axis=1)] = np.inf           # a failed restart never wins
                return x0s * 0.5, f, np.full(len(x0s), 7, dtype=np.int64)
        x0s = np.array([[2.0, 0.0], [1.0, 1.0], [0.5, 0.25], [0.3, 0.31], [-1.0, 0.0]])
        pb, mb, nfev = apd.optimize_gp_sharded(FakeGP(), None, x0s)
        ok4 = np.array_equal(pb, x0s[3] * 0.5) and np.isclose(mb, -np.sum((x0s[3] - 0.3) ** 2)) and nfev == 35
        # the tensor form of the chain gather (the NCCL path of run_ensembles_sharded; CPU tensors over gloo here):
        # one all-gather into [world, ...], concatenation along the walker axis -- same result as the NumPy form
        import torch
        t_loc = torch.arange(4 * 6 * 2, dtype=torch.float64).reshape(4, 6, 2) + 1000.0 * rank
        g1 = apd._gather_concat_device(t_loc, 1).numpy()
        ok5 = np.array_equal(g1, apd.gather_concat(t_loc.numpy(), axis=1))
        n_loc = torch.arange(6, dtype=torch.int32) + 10 * rank
        ok5 = ok5 and np.array_equal(apd._gather_concat_device(n_loc, 0).numpy(), apd.gather_concat(n_loc.numpy(), axis=0))
        # run_ensembles_sharded on the host path (no CUDA): contiguous ensemble blocks, rank-dependent seed, walker order
        class FakeSampler(object):
            def run_ensembles(self, y, p0, nsteps, bounds, nens=1, seed=0, **kw):
                W = p0.shape[0]
                ch = np.broadcast_to(p0[None], (nsteps, W, p0.shape[1])) + float(seed)
                return dict(chain=ch.copy(), log_prob=np.full((nsteps, W), float(nens)), blobs=np.zeros((nsteps, W)),
                            naccepted=np.arange(W, dtype=np.int32))
        p0 = np.arange(4 * 3 * 2, dtype=np.float64).reshape(12, 2)        # 4 ensembles of 3 walkers
        o = apd.run_ensembles_sharded(FakeSampler(), None, p0, 5, None, 4, seed=2)
        want = np.concatenate([p0[:6] + 2.0, p0[6:] + 2.0 + 7919.0])
        ok6 = (o["chain"].shape == (5, 12, 2) and np.array_equal(o["chain"][3], want) and np.all(o["log_prob"] == 2.0)
               and np.array_equal(o["naccepted"], np.tile(np.arange(6, dtype=np.int32), 2)))
        q.put((rank, bool(ok1), bool(ok2), bool(ok3 and ok4 and ok5 and ok6)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharding_helpers_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(r[0] for r in res) == [0, 1]
    for r in res:
        assert r[1:] == (True, True, True), r
