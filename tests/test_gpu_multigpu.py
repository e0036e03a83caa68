"""N>1 GPU path: candidates / ensembles / restarts sharded over ranks with ONE all-gather (NCCL).
The single-GPU result is the oracle: shards are independent, so gathered outputs must be bit-identical
to running the same shards one after another on one device.  Skipped on single-GPU boxes."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _problem():
    rng = np.random.default_rng(3)
    X = rng.uniform(-5, 5, size=(200, 3))
    y = -0.5 * np.sum(X * X, axis=1) / 4.0
    return X, y


def _make_gp(device):
    from approxposterior_b200 import GP, kernels
    X, y = _problem()
    gp = GP(kernel=kernels.ExpSquaredKernel([3.0, 3.0, 3.0], ndim=3), fit_mean=True, mean=float(np.median(y)),
            white_noise=-12.0, device=device)
    gp.compute(X, y=y)
    return gp, y


def _worker(rank, ws, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        from approxposterior_b200 import dist as apd
        gp, y = _make_gp(rank)
        bounds = [(-5.0, 5.0)] * 3
        theta, u = apd.scan_utility_sharded(gp, y, "bape", bounds, nCandidates=40000, seed=5)
        rng = np.random.default_rng(11)
        nens, nw = 4, 12
        p0 = rng.uniform(-3, 3, size=(nens * nw, 3))
        out = apd.run_ensembles_sharded(gp, y, p0, 60, bounds, nens, seed=21)
        pb, mb = apd.best_restart_sharded(np.array([[rank, 1.0]]), np.array([-5.0 - rank]))
        q.put((rank, theta, u, out["chain"], out["naccepted"], pb, mb))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_equals_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(30)
    # every rank sees the same gathered results
    for a, b in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(np.asarray(a), np.asarray(b))
    # single-GPU oracle: run the two shards one after the other on device 0
    from approxposterior_b200 import utility as ut
    gp, y = _make_gp(0)
    bounds = [(-5.0, 5.0)] * 3
    bests = [ut.scanUtility(gp, y, "bape", bounds, nCandidates=20000, seed=5 * 1000003 + r, device_out=True)[:2]
             for r in range(2)]
    ib = int(np.argmin([b[1] for b in bests]))
    assert np.array_equal(res[0][1], bests[ib][0]) and res[0][2] == bests[ib][1]
    rng = np.random.default_rng(11)
    p0 = rng.uniform(-3, 3, size=(4 * 12, 3))
    halves = [gp.run_ensembles(y, p0[r * 24:(r + 1) * 24], 60, bounds, nens=2, seed=21 + 7919 * r) for r in range(2)]
    assert np.array_equal(res[0][3], np.concatenate([h["chain"] for h in halves], axis=1))
    assert np.array_equal(res[0][4], np.concatenate([h["naccepted"] for h in halves]))
    assert res[0][6] == -5.0 and res[0][5][0] == 0


def test_two_handles_on_two_devices_in_one_process():
    """A process may hold GP handles on several GPUs (apgp_create(device)).  The > 48 KB dynamic shared-memory opt-in
    of every kernel is a per-device function attribute: each device must get it (the guards were
    process-global, so the second GPU's launches failed with invalid-value).  Same problem on both devices ->
    identical results, kernel by kernel."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    outs = []
    rng = np.random.default_rng(2)
    Xq = rng.uniform(-5, 5, size=(3000, 3))
    p0 = rng.uniform(-3, 3, size=(24, 3))
    bounds = [(-5.0, 5.0)] * 3
    gps = [_make_gp(dev) for dev in (0, 1)]            # both handles alive at once; device 1 used first
    for gp, y in reversed(gps):
        mu, var, u = gp.predict_utility(y, Xq, "bape", bounds=bounds)           # predict_var (+ grouped) kernels
        m2 = gp.predict(y, Xq, return_cov=False)                                # predict_mean
        ch = gp.run_ensembles(y, p0, 40, bounds, nens=2, seed=9)["chain"]       # sampler
        P = np.array([gp.get_parameter_vector(), gp.get_parameter_vector() + 0.2])
        ll, g = gp.log_likelihood_batch(P, y, return_grad=True)                 # loglik_small (+ gradient)
        xs, fs, _ = gp.minimize_utility(y, Xq[:4], "bape", bounds=bounds, options={"adaptive": True})   # optimisers
        ps, fn, _ = gp.minimize_nll(P, y, method="powell", options={"maxiter": 1})
        gl = gp.grad_log_likelihood(y)                                           # tiled GEMM path
        mf, vf = gp.predict(y, Xq[:3], return_cov=False, return_var=True)        # few-query kernel
        tau, win = gp.integrated_time(ch)                                        # autocorrelation kernel
        outs.append((mu, var, u, m2, ch, ll, g, xs, fs, ps, fn, gl, mf, vf, tau, win))
    # the cluster kernels (N beyond one CTA's shared memory): log-likelihood batch and the cluster-cooperative optimiser
    from approxposterior_b200 import GP, kernels
    Xb = rng.uniform(-5, 5, size=(260, 3))
    yb = -0.5 * np.sum(Xb * Xb, axis=1) / 4.0
    big = []
    for dev in (1, 0):
        gp = GP(kernel=kernels.ExpSquaredKernel([3.0, 3.0, 3.0], ndim=3), fit_mean=True, mean=float(np.median(yb)),
                white_noise=-12.0, device=dev)
        gp.compute(Xb, y=yb)
        P = np.array([gp.get_parameter_vector(), gp.get_parameter_vector() + 0.2, gp.get_parameter_vector() - 0.1])
        llb = gp.log_likelihood_batch(P, yb)
        ps, fn, _ = gp.minimize_nll(P[:2], yb, method="powell", options={"maxiter": 1})
        big.append((llb, ps, fn))
    outs = [o + b for o, b in zip(outs, big)]
    for k, (a, b) in enumerate(zip(*outs)):
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True), k


def _comm_worker(rank, ws, idq, outq):
    """C-ABI multi-GPU entry points (include/apgp.h: apgp_comm_*), no torch.distributed involved."""
    import ctypes as C
    from approxposterior_b200 import GP, kernels, _lib
    lib = _lib.load()
    if rank == 0:
        buf = C.create_string_buffer(128)
        _lib.check(lib.apgp_comm_unique_id(buf), "apgp_comm_unique_id")
        for _ in range(ws - 1):
            idq.put(buf.raw)
        idb = buf.raw
    else:
        idb = idq.get(timeout=120)
    if rank == 0:
        gp, y = _make_gp(0)
    else:                                    # same model shape, never computed: the factorisation arrives by broadcast
        gp = GP(kernel=kernels.ExpSquaredKernel([1.0, 1.0, 1.0], ndim=3), fit_mean=True, mean=0.0, white_noise=-12.0,
                device=rank)
    _lib.check(lib.apgp_comm_init(gp._h, idb, rank, ws), "apgp_comm_init")
    _lib.check(lib.apgp_comm_broadcast_factor(gp._h, 0), "apgp_comm_broadcast_factor")
    q = np.random.default_rng(7).uniform(-5, 5, size=(5000, 3))
    lo, hi = (len(q) * rank) // ws, (len(q) * (rank + 1)) // ws
    mu, var, u = gp._predict_raw(q[lo:hi], True, utility="bape", bounds=[(-5.0, 5.0)] * 3, ybest=1.0)
    send = np.ascontiguousarray(np.stack([mu, var, u], axis=1).ravel())
    recv = np.empty(send.size * ws)
    _lib.check(lib.apgp_comm_allgather(gp._h, _lib.ptr(send), _lib.ptr(recv), send.size, 1), "apgp_comm_allgather")
    _lib.check(lib.apgp_comm_destroy(gp._h), "apgp_comm_destroy")
    outq.put((rank, recv))


@pytest.mark.timeout(300)
def test_c_abi_comm_broadcast_factor_and_allgather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    idq, outq = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_comm_worker, args=(r, 2, idq, outq)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([outq.get(timeout=240) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(30)
    assert np.array_equal(res[0][1], res[1][1], equal_nan=True)              # every rank holds the same gathered result
    gp, y = _make_gp(0)                                                        # single-GPU oracle: the whole batch at once
    q = np.random.default_rng(7).uniform(-5, 5, size=(5000, 3))
    mu, var, u = gp._predict_raw(q, True, utility="bape", bounds=[(-5.0, 5.0)] * 3, ybest=1.0)
    ref = np.stack([mu, var, u], axis=1).ravel()
    assert np.array_equal(res[0][1], ref, equal_nan=True)                     # rank 1 never factorised: broadcast state is exact
