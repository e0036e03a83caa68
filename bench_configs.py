"""bench.py --config {cfg1, cfg2, cfg4, cfg5}: the other BASELINE.json configurations, each with the same JSON line
contract as the headline (cfg3, in bench.py itself): `value` device-timed, `e2e` through the host-buffer public API,
`roofline` for the dominant kernel, `cpu_baseline`, and an `--impl reference` arm on the host cores.

  cfg1  README Rosenbrock 2-D BAPE run (m0=50, m=20, nmax=2, 20 walkers x 2e4): BAPE iteration time.  Reference arm =
        the reference's OWN ApproxPosterior.run (baseline/_ref, unmodified) on the oracle-backed george/emcee shim.
  cfg2  65 536 walkers = 2048 ensembles x 32 on the N=1024 Rosenbrock surrogate, 1000 steps per step, sharded over the
        ranks through dist.run_ensembles_sharded (strong scaling): mean-only lnprob evals/s.
  cfg4  bayesOpt on the 10-D Branin-style objective, 64 optGP restarts (sharded over ranks): bayesOpt iteration time.
  cfg5  predict sweep point (--N, --d; default N=512, d=5): 1e8 queries per step sharded through dist.predict_sharded
        (strong scaling): mean+var evals/s.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


# ------------------------------------------------------------------------------------------------ problems
def readme_problem():
    from approxposterior_b200 import likelihood as lh
    np.random.seed(57)
    theta = lh.rosenbrockSample(50)
    y = np.array([lh.rosenbrockLnlike(t) + lh.rosenbrockLnprior(t) for t in theta])
    return theta, y


def cfg2_problem():
    """Rosenbrock 2-D surrogate grown to N = 1024 training points (hyper-parameters fixed)."""
    from scipy.optimize import rosen
    rng = np.random.RandomState(57)
    theta = rng.uniform(-5, 5, size=(1024, 2))
    y = np.array([-rosen(t) / 100.0 for t in theta])
    return theta, y, np.array([0.5, 1.2]), float(np.median(y))


def branin10(X):
    def branin(u, v):
        return (v - 5.1 / (4 * np.pi ** 2) * u ** 2 + 5 / np.pi * u - 6) ** 2 + 10 * (1 - 1 / (8 * np.pi)) * np.cos(u) + 10
    X = np.atleast_2d(X)
    U = (X + 5) / 10
    return -sum(branin(15 * U[:, 2 * i] - 5, 15 * U[:, 2 * i + 1]) for i in range(5)) / 100.0


def cfg5_problem(N, d):
    rng = np.random.default_rng(N * 100 + d)
    X = rng.uniform(-5, 5, size=(N, d))
    y = rng.standard_normal(N)
    return X, y, np.full(d, np.log(d)), 0.0


def F_ll(N, d):
    return N ** 3 / 3.0 + 2.0 * N * N + (3 * d + 1) * N * N / 2.0


def F_mv(N, d):
    return float(N) * N + (3 * d + 6) * N


def F_m(N, d):
    return (3.0 * d + 3.0) * N


# ------------------------------------------------------------------------------------------------ harness
class Harness(object):
    def __init__(self, args, bench):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.bench, self.args = torch, dist, bench, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.clocks = bench.ClockSampler(self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, step, warmup, steps):
        """W untimed + exactly K timed steps, CUDA events on torch's current stream (the engine runs on it for device
        tensors; host-API steps are synchronous, so the events bracket them as well), max over ranks.  Returns ms."""
        torch = self.torch
        if self.rank == 0:
            self.clocks.start(); self.clocks.wait_first_sample()
        for i in range(warmup):
            step(i)
        self.barrier()
        self.clocks.mark()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for i in range(steps):
            step(warmup + i)
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        self.clk = self.clocks.stop() if self.rank == 0 else None
        return float(t.item())

    def wall_max(self, seconds):
        t = self.torch.tensor([seconds], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()

    def line(self, **kw):
        a = self.args
        base = {"n_gpus": self.world, "steps": a.steps, "warmup": a.warmup, "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "clocks": self.clk}
        base.update(kw)
        return base


def reference_root():
    """The unmodified reference package as installed by `pip install --no-deps --target baseline/_ref`;
    /root/reference is never read at run time."""
    cand = os.path.join(ROOT, "baseline", "_ref")
    return cand if os.path.isfile(os.path.join(cand, "approxposterior", "approx.py")) else None


# ------------------------------------------------------------------------------------------------ cfg1
CFG1_METRIC = "BAPE iteration time (README config: m0=50, m=20, nmax=2, 20 walkers x 2e4 steps)"
CFG1_UNIT = "s per BAPE iteration"


def cfg1_config(extra=None):
    cfg = {"workload": "cfg1: README Rosenbrock 2-D BAPE run, m0=50, m=20 new design points per iteration, nGPRestarts=3, "
                       "ExpSquared GP; one step = one BAPE iteration (20 x {utility multistart, forward model, GP "
                       "refit + optGP}) (BASELINE.json configs[0])",
           "l2_policy": "working set (N <= 90) is far below L2; latency-bound by construction"}
    if extra:
        cfg.update(extra)
    return cfg


def _run_readme(ap_module, gputils, lh, prior, nmax, iterations, seed=57):
    theta, y = readme_problem()
    gp = gputils.defaultGP(theta, y, white_noise=-12)
    ap = ap_module.ApproxPosterior(theta=theta, y=y, gp=gp, lnprior=prior, lnlike=lh.rosenbrockLnlike,
                                   priorSample=lh.rosenbrockSample, bounds=[(-5, 5), (-5, 5)], algorithm="bape")
    t0 = time.perf_counter()
    ap.run(m=20, nmax=nmax, estBurnin=True, nGPRestarts=3, mcmcKwargs={"iterations": int(iterations)},
           samplerKwargs={"nwalkers": 20}, cache=False, verbose=False, thinChains=False, onlyLastMCMC=True,
           timing=True, seed=seed)
    return ap, time.perf_counter() - t0


def run_cfg1(args, bench):
    from approxposterior_b200 import approx, gpUtils, likelihood as lh
    H = Harness(args, bench)
    times, mcmc, launches = [], [], [0]

    def one_run(i):
        ap, tot = _run_readme(approx, gpUtils, lh, lh.BoxPrior([(-5, 5), (-5, 5)]), 2, 2.0e4, seed=57 + H.rank)
        times.extend(ap.trainingTime); mcmc.append(ap.mcmcTime[-1])
        launches[0] += ap.gp.launch_count
        return ap
    nruns = (args.steps + 1) // 2
    for _ in range(max(1, (args.warmup + 1) // 2)):
        one_run(0)
    del times[:], mcmc[:]
    launches[0] = 0
    if H.rank == 0:
        H.clocks.start(); H.clocks.wait_first_sample()
    H.barrier(); H.clocks.mark()
    ap = None
    for i in range(nruns):
        ap = one_run(i)
    H.barrier()
    H.clk = H.clocks.stop() if H.rank == 0 else None
    per_iter = H.wall_max(float(np.sum(times[:args.steps]) / args.steps))
    # roofline of the dominant kernel (device Powell over the hyper-parameters, minimize_nll_kernel): algorithmic
    # log-likelihood flops of all its objective evaluations over its wall time, at the run's mid-size training set
    theta, y = ap.theta[:70], ap.y[:70]
    g = gpUtils.defaultGP(theta, y)
    np.random.seed(1)
    x0 = np.array([[np.median(y)] + list(np.random.randn(2)) for _ in range(3)])
    g.minimize_nll(x0, y)
    t0 = time.perf_counter(); _, _, nfev = g.minimize_nll(x0, y); dt = time.perf_counter() - t0
    # e2e: the reference's own unmodified driver on the engine (drop-in shims), when the reference install travelled
    e2e = None
    ref = reference_root()
    if H.rank == 0 and ref is not None:
        from approxposterior_b200 import compat
        from oracle.refshim import scipy_x0_compat
        compat.install(box_prior_sampler=True)      # rosenbrockLnprior IS the box over ap.bounds: device sampler
        sys.path.insert(0, ref)
        try:
            import approxposterior.approx as rap, approxposterior.gpUtils as rgu, approxposterior.likelihood as rlh, approxposterior.utility as rut
            scipy_x0_compat(rut)
            rp0, tot0 = _run_readme(rap, rgu, rlh, rlh.rosenbrockLnprior, 2, 2.0e4)      # level 0: scalar SciPy loops
            compat.accelerate(box_prior=True)                                            # level 1: batched multistarts
            _run_readme(rap, rgu, rlh, rlh.rosenbrockLnprior, 1, 2000)
            rp, tot = _run_readme(rap, rgu, rlh, rlh.rosenbrockLnprior, 2, 2.0e4)
            e2e = {"value": float(np.mean(rp.trainingTime)), "unit": CFG1_UNIT, "mcmc_s": float(rp.mcmcTime[-1]), "run_total_s": tot,
                   "scalar_loops": {"value": float(np.mean(rp0.trainingTime)), "mcmc_s": float(rp0.mcmcTime[-1]), "run_total_s": tot0,
                                    "note": "compat.install() only: the reference's own SciPy loops, one GPU call per "
                                            "objective evaluation (~3000 launches per BAPE iteration)"},
                   "h2d_bytes_per_step": int(20 * 3 * 90 * 8 * 400), "d2h_bytes_per_step": int(20 * 400 * 8 * 2),
                   "bytes_note": "approximate: one training-set upload per hyper-parameter evaluation and one scalar back, "
                                 "~400 evaluations per refit through the reference's scalar SciPy loop",
                   "api": "the reference's own approxposterior.ApproxPosterior.run (baseline/_ref, unmodified) with george/"
                          "emcee resolved to the engine by approxposterior_b200.compat.install(box_prior_sampler=True) and its "
                          "two multistart drivers by compat.accelerate(box_prior=True)"}
        finally:
            sys.path.remove(ref); compat.uninstall()
    H.finish()
    if H.rank != 0:
        return
    peak = bench.measure_dgemm_peak(H.torch, H.dev)
    ach = float(np.sum(nfev)) * F_ll(70, 2) / dt * 1e-12
    s = ap.sampler.get_chain(discard=ap.iburns[-1], flat=True)
    if e2e is None:
        e2e = {"value": per_iter, "unit": CFG1_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "api": "ApproxPosterior.run (host arrays in, host arrays out); reference install not present"}
    bench.emit_line(H.line(metric=CFG1_METRIC, value=per_iter, unit=CFG1_UNIT, ms_per_step=per_iter * 1e3,
                           higher_is_better=False, scaling="weak",
                           config=cfg1_config({"parallelism": "replicas only (a BAPE run is sequential): %d independent runs" % H.world,
                                               "mcmc_s": float(np.mean(mcmc)), "posterior_mean": s.mean(axis=0).tolist()}),
                           e2e=e2e, gpu_launches=int(launches[0]),
                           roofline={"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                     "traffic": None, "kernel": "minimize_nll_kernel (device Powell, one CTA per restart, N=70)",
                                     "note": "latency-bound: 3 CTAs, each a serial chain of ~150-700 Cholesky evaluations; the "
                                             "fraction of the FP64 peak is not the figure of merit here, the 14 us per "
                                             "evaluation is (profiles/r02_device_optimizers_profile_block8.jsonl)",
                                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"},
                           cpu_baseline=None if args.no_cpu else cfg1_reference(bounded=True)))


def cfg1_reference(bounded=True):
    """Reference arm: the reference's own ApproxPosterior.run on the oracle-backed shims (CPU).  The MCMC leg is a
    bounded sample (2000 of the 2e4 steps) -- BAPE iteration time is the training half (ap.trainingTime)."""
    from oracle import refshim
    ref = reference_root()
    refshim.install()
    try:
        if ref is not None:
            sys.path.insert(0, ref)
            import approxposterior.approx as rap, approxposterior.gpUtils as rgu, approxposterior.likelihood as rlh, approxposterior.utility as rut
            refshim.scipy_x0_compat(rut)
            ap, tot = _run_readme(rap, rgu, rlh, rlh.rosenbrockLnprior, 1 if bounded else 2, 2000)
            kind, what = "reference", ("the reference's own ApproxPosterior.run (baseline/_ref, unmodified) over the NumPy/SciPy "
                                       "restatement of george/emcee (oracle/refshim.py)")
            sys.path.remove(ref)
        else:
            from approxposterior_b200 import approx, gpUtils, likelihood as lh
            theta, y = readme_problem()
            g = refshim.GP(kernel=refshim._ExpSquared(np.fabs(np.random.randn(2)), 2), fit_mean=True, mean=np.median(y), white_noise=-12)
            g.compute(theta)
            ap = approx.ApproxPosterior(theta=theta, y=y, gp=g, lnprior=lh.rosenbrockLnprior, lnlike=lh.rosenbrockLnlike,
                                        priorSample=lh.rosenbrockSample, bounds=[(-5, 5), (-5, 5)], algorithm="bape")
            ap.run(m=20, nmax=1, estBurnin=True, nGPRestarts=3, mcmcKwargs={"iterations": 2000},
                   samplerKwargs={"nwalkers": 20, "engine": "host-rng"}, cache=False, verbose=False, thinChains=False,
                   onlyLastMCMC=True, timing=True, seed=57, batched=False)
            kind, what = "port", "this repo's restated drivers over the oracle (reference install absent)"
    finally:
        refshim.uninstall()
        for n in [n for n in sys.modules if n == "approxposterior" or n.startswith("approxposterior.")]:
            del sys.modules[n]
    return {"value": float(np.mean(ap.trainingTime)), "unit": CFG1_UNIT, "cores": 1, "kind": kind,
            "sample": "%d BAPE iteration(s) of 20 design points; MCMC leg bounded to 2000 of 2e4 steps (%.2f s, i.e. ~%.0f s "
                      "at full length); %s" % (len(ap.trainingTime), ap.mcmcTime[-1], ap.mcmcTime[-1] * 10, what),
            "mcmc_2000_steps_s": float(ap.mcmcTime[-1])}


def run_cfg1_reference(args, bench):
    cb = cfg1_reference(bounded=args.steps <= 1)
    bench.emit_line({"impl": "reference", "metric": CFG1_METRIC, "value": cb["value"], "unit": CFG1_UNIT, "n_gpus": args.gpus,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["value"] * 1e3, "higher_is_better": False,
                     "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg1_config(),
                     "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": CFG1_UNIT, "h2d_bytes_per_step": 0,
                                                 "d2h_bytes_per_step": 0}, "gpu_launches": 0})


# ------------------------------------------------------------------------------------------------ cfg2
CFG2_METRIC = "GP-surrogate lnprob evals/s (mean-only, device ensemble sampler, N=1024, d=2, 65536 walkers)"
NENS2, NW2, NSTEPS2 = 2048, 32, 1000


def cfg2_config(extra=None):
    cfg = {"workload": "cfg2: Rosenbrock 2-D surrogate on N=1024 training points, 65 536 walkers = 2048 independent ensembles x 32 "
                       "walkers (stretch move), 1000 sampler steps per bench step (BASELINE.json configs[1])",
           "N_train": 1024, "d": 2, "walkers": NENS2 * NW2, "sampler_steps_per_step": NSTEPS2,
           "l2_policy": "state lives in shared memory for the whole chain; the 24 KB scaled training set is staged once per CTA"}
    if extra:
        cfg.update(extra)
    return cfg


def run_cfg2(args, bench):
    from approxposterior_b200 import GP, kernels, dist as apd
    H = Harness(args, bench)
    torch = H.torch
    theta, y, logM, mean = cfg2_problem()
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=2), fit_mean=True, mean=mean, white_noise=-12.0, device=H.local)
    gp.compute(theta, y=y)
    bounds = [(-5.0, 5.0)] * 2
    lo, hi = apd.shard_bounds(NENS2, H.rank, H.world)
    nloc = hi - lo
    rng = np.random.default_rng(2)
    p0_all = rng.uniform(-5, 5, size=(NENS2 * NW2, 2))
    p0_dev = torch.from_numpy(p0_all[lo * NW2:hi * NW2]).to(H.dev)
    l0 = gp.launch_count

    def step(i):       # device-resident: p0 and the (thinned) chain stay on the GPU
        gp.run_ensembles(y, p0_dev, NSTEPS2, bounds, nens=nloc, seed=100 + i + 7919 * H.rank, thin=NSTEPS2, device_out=True)
    ms = H.timed(step, args.warmup, args.steps)
    launches = gp.launch_count - l0
    evals = float(NENS2 * NW2) * NSTEPS2 * args.steps
    # e2e: the public sharded API with host buffers -- p0 H2D, chain (every 20th step) / log_prob / blobs D2H, one all-gather
    e2e_steps = max(1, min(args.steps, 3))
    for i in range(3):   # warm-up with the caller's own pattern (the previous result is alive during the next call): at
        out = apd.run_ensembles_sharded(gp, y, p0_all, NSTEPS2, bounds, NENS2, seed=3 + i, thin=20)   # N > 1 the two sets of
    H.barrier()          # pinned result blocks then both sit in torch's host cache and no timed call pays a cudaHostAlloc
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = apd.run_ensembles_sharded(gp, y, p0_all, NSTEPS2, bounds, NENS2, seed=6 + i, thin=20)
    H.barrier()
    e2e_s = H.wall_max(time.perf_counter() - t0)
    # burn-in estimate on the device-resident chain of this rank (N2): no chain D2H
    dev_out = gp.run_ensembles(y, p0_dev, NSTEPS2, bounds, nens=nloc, seed=3, device_out=True)
    torch.cuda.synchronize(H.dev)
    t0 = time.perf_counter(); tau = gp.integrated_time(dev_out["chain"]); t_tau = time.perf_counter() - t0
    del dev_out
    H.finish()
    if H.rank != 0:
        return
    peak = bench.measure_dgemm_peak(torch, H.dev)
    kernel_ms = ms / args.steps
    ach = float(nloc * NW2) * NSTEPS2 * F_m(1024, 2) / (kernel_ms * 1e-3) * 1e-12
    nst = NSTEPS2 // 20
    bench.emit_line(H.line(metric=CFG2_METRIC, value=evals / (ms * 1e-3), unit="evals/s", ms_per_step=ms / args.steps,
                           higher_is_better=True, scaling="strong",
                           config=cfg2_config({"parallelism": "2048 ensembles sharded over %d GPU(s) (dist.run_ensembles_sharded), one "
                                                              "all-gather of chains" % H.world,
                                               "integrated_time_on_device_s": t_tau,
                                               "tau": None if tau is None else [float(v) for v in tau[0]]}),
                           e2e={"value": float(NENS2 * NW2) * NSTEPS2 * e2e_steps / e2e_s, "unit": "evals/s", "steps": e2e_steps,
                                "h2d_bytes_per_step": int(nloc * NW2 * 2 * 8),
                                # every rank receives the gathered result: at N > 1 the whole chain crosses each rank's PCIe link
                                "d2h_bytes_per_step": int(nst * (NENS2 if H.world > 1 else nloc) * NW2 * (2 + 2) * 8
                                                          + (NENS2 if H.world > 1 else nloc) * NW2 * 4),
                                "api": "dist.run_ensembles_sharded(gp, y, p0 [host], 1000, bounds, 2048, thin=20) -> gathered host "
                                       "chain / log_prob / blobs / naccepted"},
                           gpu_launches=int(launches),
                           roofline={"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                     "traffic": None, "kernel": "sampler_kernel (one CTA per ensemble, whole chain in one launch)",
                                     "kernel_ms": kernel_ms, "flops_per_eval": F_m(1024, 2),
                                     "note": "algorithmic flops (3d+3)N per mean-only evaluation (SURVEY 8d) count the exponential "
                                             "as ONE flop; it costs 10 FP64 instructions on the same pipe, so the executed "
                                             "FP64 instruction rate is (2d+12)/(3d+3) x 2 of the achieved figure. This rank's "
                                             "share of the walkers over this rank's kernel time",
                                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (the FP64 pipe; DFMA issue peak is "
                                                    "half the DMMA flop rate)"},
                           cpu_baseline=None if args.no_cpu else cfg2_reference(bounded_s=10.0)))


def _cfg2_cpu_worker(task):
    """One 32-walker ensemble of the cfg2 workload on the oracle (one process = one core; BLAS limited to one thread)."""
    seed, nsteps = task
    from threadpoolctl import threadpool_limits
    from oracle import GPOracle, stretch_move_oracle
    from oracle.sampler_oracle import gpll_batch
    with threadpool_limits(limits=1):
        theta, y, logM, mean = cfg2_problem()
        orc = GPOracle(2, np.exp(logM), mean=mean, white_noise=-12.0)
        orc.compute(theta)
        yo = y.copy(); yo.setflags(write=False)
        lo, hi = np.full(2, -5.0), np.full(2, 5.0)
        rs = np.random.RandomState(seed)
        p0 = rs.uniform(-5, 5, size=(NW2, 2))
        t0 = time.perf_counter()
        stretch_move_oracle(lambda q: gpll_batch(orc, yo, q, lo, hi), p0, nsteps, rng=rs)
        return time.perf_counter() - t0


def cfg2_reference(bounded_s=10.0):
    """CPU arm: emcee's restatement (oracle/sampler_oracle.py) on the oracle GP, per-half-step batched lnprob (a
    best-effort CPU form: the reference itself makes one Python call per walker).  The workload is 2048 INDEPENDENT
    ensembles, so the arm runs one ensemble per host core side by side (spawned processes, one BLAS thread each) and
    reports their aggregate rate."""
    import multiprocessing as mp
    from oracle import GPOracle
    cores = max(1, os.cpu_count() or 1)
    ctx = mp.get_context("spawn")            # the parent may hold a CUDA context: no fork
    with ctx.Pool(cores) as pool:
        per100 = max(pool.map(_cfg2_cpu_worker, [(1000 + i, 100) for i in range(cores)]))      # also warms the workers up
        nsteps = int(max(100, min(20000, 100 * bounded_s / per100)))
        t0 = time.perf_counter()
        pool.map(_cfg2_cpu_worker, [(i, nsteps) for i in range(cores)])
        dt = time.perf_counter() - t0
    theta, y, logM, mean = cfg2_problem()
    orc = GPOracle(2, np.exp(logM), mean=mean, white_noise=-12.0)
    orc.compute(theta)
    yo = y.copy(); yo.setflags(write=False)
    p0 = np.random.RandomState(0).uniform(-5, 5, size=(NW2, 2))
    t0 = time.perf_counter()
    for i in range(300):
        orc.predict(yo, p0[i % NW2:i % NW2 + 1], return_var=False)
    per_call = 300 / (time.perf_counter() - t0)
    return {"value": cores * NW2 * nsteps / dt, "unit": "evals/s", "cores": cores, "kind": "port",
            "sample": "%d 32-walker ensembles side by side (one per host core) x %d steps (of 2048 ensembles x 1000) through the "
                      "emcee restatement with per-half-step batched lnprob on the oracle GP, %.1f s" % (cores, nsteps, dt),
            "per_call_evals_per_s": per_call,
            "note": "per_call = one predict per walker per step on one core, the shape of approx.py:178"}


def run_cfg2_reference(args, bench):
    cb = cfg2_reference(bounded_s=min(60.0, 6.0 * max(1, args.steps)))
    bench.emit_line({"impl": "reference", "metric": CFG2_METRIC, "value": cb["value"], "unit": "evals/s", "n_gpus": args.gpus,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                     "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg2_config(), "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})


# ------------------------------------------------------------------------------------------------ cfg4
CFG4_METRIC = "bayesOpt iteration time (10-D Branin-style objective, 64 optGP restarts, N=64 growing)"
CFG4_UNIT = "s per bayesOpt iteration"


def cfg4_config(extra=None):
    cfg = {"workload": "cfg4: bayesOpt on the 10-D synthetic Branin-style objective (sum of 5 Branin pairs on [-5,5]^10), N0=64 "
                       "design points, fitAmp (P=12), 64 optGP restarts (Powell) per iteration, Jones utility, findMAP; one step "
                       "= one bayesOpt iteration (BASELINE.json configs[3])",
           "d": 10, "P": 12, "restarts": 64, "l2_policy": "working set below L2; latency-bound optimiser chains"}
    if extra:
        cfg.update(extra)
    return cfg


def _cfg4_ap(approx_mod, gputils, GPcls, kernels, seed=64):
    from approxposterior_b200 import likelihood as lh
    rng = np.random.RandomState(seed)
    X = rng.uniform(-5, 5, size=(64, 10))
    y = branin10(X)
    bounds = [(-5.0, 5.0)] * 10
    gp = GPcls(kernel=float(np.var(y)) * kernels.ExpSquaredKernel(np.ones(10), ndim=10), fit_mean=True, mean=float(np.median(y)),
               white_noise=-12.0)
    gp.compute(X)
    prior = lh.BoxPrior(bounds)
    ap = approx_mod.ApproxPosterior(theta=X, y=y, gp=gp, lnprior=prior, lnlike=lambda t, *a, **k: float(branin10(t)[0]),
                                    priorSample=lambda n=1: np.random.uniform(-5, 5, size=(n, 10)).squeeze(), bounds=bounds,
                                    algorithm="jones")
    return ap


def run_cfg4(args, bench):
    from approxposterior_b200 import GP, approx, gpUtils, kernels
    H = Harness(args, bench)
    ap = _cfg4_ap(approx, gpUtils, GP, kernels)
    iters, evals = [], []
    orig = gpUtils.optimizeGP

    def timed_opt(*a, **k):
        t0 = time.perf_counter(); r = orig(*a, **k)
        st = getattr(gpUtils.optimizeGP, "last_stats", None) or getattr(orig, "last_stats", None) or {"evals": 0}
        evals.append((time.perf_counter() - t0, st["evals"], len(ap.y)))
        return r
    gpUtils.optimizeGP = timed_opt
    approx.gpUtils.optimizeGP = timed_opt
    np.random.seed(64)                           # same stream on every rank: the sharded restarts need identical starts
    total = args.warmup + args.steps
    if H.rank == 0:
        H.clocks.start(); H.clocks.wait_first_sample()
    l0 = None
    for it in range(total):
        if it == args.warmup:
            H.barrier(); H.clocks.mark(); l0 = ap.gp.launch_count; del evals[:]
        t0 = time.perf_counter()
        # one bayesOpt iteration = findNextPoint + optGP(64 restarts) + findMAP (approx.py:1074-1116); findMAP is called
        # explicitly because bayesOpt(nmax=1, findMAP=True) trips over its own squeeze()d one-element history
        # (approx.py:1146-1150 index a 0-d array)
        ap.bayesOpt(nmax=1, verbose=False, cache=False, nGPRestarts=64, nMinObjRestarts=5, initGPOpt=False, findMAP=False,
                    seed=None, kmax=10 ** 6)
        ap.findMAP(nRestarts=5)
        if it >= args.warmup:
            iters.append(time.perf_counter() - t0)
    H.barrier()
    H.clk = H.clocks.stop() if H.rank == 0 else None
    gpUtils.optimizeGP = orig
    approx.gpUtils.optimizeGP = orig
    per_iter = H.wall_max(float(np.mean(iters)))
    launches = ap.gp.launch_count - (l0 or 0)
    H.finish()
    if H.rank != 0:
        return
    peak = bench.measure_dgemm_peak(H.torch, H.dev)
    t_opt = float(np.sum([e[0] for e in evals])); n_ev = float(np.sum([e[1] for e in evals]))
    Nmid = int(np.mean([e[2] for e in evals]))
    ach = n_ev * F_ll(Nmid, 10) / t_opt * 1e-12
    bench.emit_line(H.line(metric=CFG4_METRIC, value=per_iter, unit=CFG4_UNIT, ms_per_step=per_iter * 1e3, higher_is_better=False,
                           scaling="strong",
                           config=cfg4_config({"parallelism": "64 optGP restarts sharded over %d GPU(s) (dist.optimize_gp_sharded), one "
                                                              "all-gather of (mll, p)" % H.world,
                                               "optGP_s_per_iteration": t_opt / len(evals), "nll_evals_per_optGP": n_ev / len(evals),
                                               "nll_evals_per_s": n_ev / t_opt, "N_final": int(len(ap.y)),
                                               "best_y": float(np.max(ap.y))}),
                           e2e={"value": per_iter, "unit": CFG4_UNIT, "h2d_bytes_per_step": int(64 * 12 * 8 + Nmid * 11 * 8),
                                "d2h_bytes_per_step": int(64 * 14 * 8),
                                "api": "ApproxPosterior.bayesOpt(nmax=1, nGPRestarts=64, findMAP=True): host arrays in and out; the "
                                       "step IS the public call, so e2e = value"},
                           gpu_launches=int(launches),
                           roofline={"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                                     "kernel": "minimize_nll_kernel (64 CTAs, one restart each; device Powell over P=12)",
                                     "flops_per_eval": F_ll(Nmid, 10), "N": Nmid,
                                     "note": "latency-bound serial Cholesky chains on 64 of 148 SMs",
                                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"},
                           cpu_baseline=None if args.no_cpu else cfg4_reference(restarts=2)))


def _cfg4_cpu_worker(task):
    """`restarts` optGP restarts (SciPy Powell on the oracle GP) in one process = one core."""
    seed, restarts = task
    from threadpoolctl import threadpool_limits
    from oracle import refshim
    from approxposterior_b200 import approx, gpUtils
    with threadpool_limits(limits=1):
        ap = _cfg4_ap(approx, gpUtils, refshim.GP,
                      type("K", (), {"ExpSquaredKernel": staticmethod(lambda m, ndim: refshim._ExpSquared(m, ndim))}))
        np.random.seed(seed)
        t0 = time.perf_counter()
        gpUtils.optimizeGP(ap.gp, ap.theta, ap.y, nGPRestarts=restarts, method="powell", batched=False)
        return time.perf_counter() - t0


def cfg4_reference(restarts=2):
    """CPU arm: the same bayesOpt iteration through this repo's drivers on the oracle GP with SciPy's Powell.  The 64 optGP
    restarts are independent, so they are spread over the host cores: every core runs `restarts` of them side by side
    (spawned processes, one BLAS thread each) and the measured restart rate is scaled to 64; the sequential rest of the
    iteration (findNextPoint, findMAP) is timed once on the parent."""
    import multiprocessing as mp
    from oracle import refshim
    from approxposterior_b200 import approx, gpUtils
    cores = max(1, os.cpu_count() or 1)
    ctx = mp.get_context("spawn")            # the parent may hold a CUDA context: no fork
    with ctx.Pool(cores) as pool:
        pool.map(_cfg4_cpu_worker, [(1000 + i, 1) for i in range(cores)])             # start-up and imports
        t0 = time.perf_counter()
        pool.map(_cfg4_cpu_worker, [(64 + i, restarts) for i in range(cores)])
        t_round = time.perf_counter() - t0
    t_opt64 = 64.0 * t_round / (cores * restarts)
    ap = _cfg4_ap(approx, gpUtils, refshim.GP, type("K", (), {"ExpSquaredKernel": staticmethod(lambda m, ndim: refshim._ExpSquared(m, ndim))}))
    np.random.seed(64)
    t0 = time.perf_counter()
    ap.bayesOpt(nmax=1, verbose=False, cache=False, nGPRestarts=1, nMinObjRestarts=5, initGPOpt=False, findMAP=False, kmax=10 ** 6,
                batched=False)
    ap.findMAP(nRestarts=5)
    t_rest = time.perf_counter() - t0
    t0 = time.perf_counter()
    gpUtils.optimizeGP(ap.gp, ap.theta, ap.y, nGPRestarts=1, method="powell", batched=False)     # the refit inside bayesOpt above
    t_one = time.perf_counter() - t0
    per_iter = t_opt64 + max(0.0, t_rest - t_one)
    return {"value": per_iter, "unit": CFG4_UNIT, "cores": cores, "kind": "port",
            "sample": "%d x %d of the 64 optGP restarts side by side on %d cores (%.1f s; restart rate scaled to 64: %.1f s) + one full "
                      "findNextPoint/findMAP pass (%.1f s, one core) on the oracle GP through SciPy's Powell / Nelder-Mead; "
                      "%.1f s on a single core" % (cores, restarts, cores, t_round, t_opt64, t_rest, t_round / restarts * 64 + t_rest)}


def run_cfg4_reference(args, bench):
    cb = cfg4_reference(restarts=2)
    bench.emit_line({"impl": "reference", "metric": CFG4_METRIC, "value": cb["value"], "unit": CFG4_UNIT, "n_gpus": args.gpus,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["value"] * 1e3, "higher_is_better": False,
                     "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg4_config(),
                     "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": CFG4_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "gpu_launches": 0})


# ------------------------------------------------------------------------------------------------ cfg5
QTOT5 = 100_000_000


def cfg5_metric(N, d):
    return "GP-surrogate lnprob evals/s (fp64 mean+var, N=%d, d=%d, 1e8 queries)" % (N, d)


def cfg5_config(N, d, extra=None):
    cfg = {"workload": "cfg5: synthetic predict sweep point N=%d, d=%d: 1e8 query points ~ U(-5,5)^d per step, mean + variance "
                       "(BASELINE.json configs[4])" % (N, d),
           "N_train": N, "d": d, "queries_per_step": QTOT5,
           "l2_policy": "query chunks of 2^22 rotate through 4 buffers (%.0f MB each, > L2 with the outputs)" % ((1 << 22) * d * 8 / 1e6)}
    if extra:
        cfg.update(extra)
    return cfg


def run_cfg5(args, bench):
    from approxposterior_b200 import GP, kernels, dist as apd
    H = Harness(args, bench)
    torch = H.torch
    N, d = args.N, args.d
    X, y, logM, mean = cfg5_problem(N, d)
    gp = GP(kernel=kernels.ExpSquaredKernel(np.exp(logM), ndim=d), fit_mean=True, mean=mean, white_noise=-12.0, device=H.local)
    gp.compute(X, y=y)
    chunk = 1 << 22
    gen = torch.Generator(device=H.dev); gen.manual_seed(5 + H.rank)
    bufs = [(-5.0 + 10.0 * torch.rand((chunk, d), dtype=torch.float64, device=H.dev, generator=gen)) for _ in range(4)]
    make = lambda first, count: bufs[(first // chunk) % 4][:count]
    l0 = gp.launch_count
    out = {}

    def step(i):
        out["s"] = apd.predict_sharded(gp, y, QTOT5, make, chunk=chunk)
    ms = H.timed(step, args.warmup, args.steps)
    launches = gp.launch_count - l0
    assert out["s"]["count"] == QTOT5
    # kernel-only time of one chunk (CUDA events on the launching stream = torch's current stream)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    gp._predict_raw(bufs[0], True); torch.cuda.synchronize(H.dev)
    e0.record(); gp._predict_raw(bufs[1], True); e1.record(); torch.cuda.synchronize(H.dev)
    kernel_ms = e0.elapsed_time(e1)
    # e2e: the reference-facing call on pinned HOST buffers (pipelined H2D / kernel / D2H inside the library)
    Qh = 1 << 23
    hq = torch.empty((Qh, d), dtype=torch.float64).pin_memory()
    hq.copy_(torch.cat([bufs[0], bufs[1]]).cpu())
    outs = tuple(torch.empty(Qh, dtype=torch.float64, pin_memory=True).numpy() for _ in range(2))
    gp._predict_raw(hq.numpy(), True, out=(outs[0], outs[1], None))
    H.barrier()
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        gp._predict_raw(hq.numpy(), True, out=(outs[0], outs[1], None))
    H.barrier()
    e2e_s = H.wall_max(time.perf_counter() - t0)
    H.finish()
    if H.rank != 0:
        return
    peak = bench.measure_dgemm_peak(torch, H.dev)
    ach = F_mv(N, d) * chunk / (kernel_ms * 1e-3) * 1e-12
    hbm = None
    try:
        import json
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    bench.emit_line(H.line(metric=cfg5_metric(N, d), value=QTOT5 * args.steps / (ms * 1e-3), unit="evals/s", ms_per_step=ms / args.steps,
                           higher_is_better=True, scaling="strong",
                           config=cfg5_config(N, d, {"parallelism": "1e8 queries sharded over %d GPU(s) (dist.predict_sharded), one all-gather "
                                                                    "of per-rank sums" % H.world,
                                                     "checksum": out["s"]}),
                           e2e={"value": Qh * H.world * reps / e2e_s, "unit": "evals/s", "steps": reps, "queries_per_call": Qh,
                                "h2d_bytes_per_step": Qh * d * 8, "d2h_bytes_per_step": 2 * Qh * 8,
                                "api": "GP.predict(y, pinned host ndarray [2^23, d], return_var=True) -> host (mu, var); the library "
                                       "overlaps the H2D copy, the kernel and the D2H copies of successive slices"},
                           gpu_launches=int(launches),
                           roofline={"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                                     "kernel": "predict_var kernels (fused K* panel + DMMA triangular GEMM), one 2^22-query chunk",
                                     "kernel_ms": kernel_ms, "flops_per_eval": F_mv(N, d),
                                     "algorithmic_hbm_gbs": (8.0 * d + 16.0) * chunk / (kernel_ms * 1e-3) * 1e-9,
                                     "hbm_peak_gbs": hbm,
                                     "note": "FP64-pipe bound at every N of the sweep (SURVEY 8d): compulsory HBM traffic is 8d+16 bytes "
                                             "per evaluation; algorithmic_hbm_gbs is that figure over the kernel time",
                                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run"},
                           cpu_baseline=None if args.no_cpu else cfg5_reference(N, d, target_s=10.0)))


def cfg5_reference(N, d, target_s=10.0):
    from oracle import GPOracle
    X, y, logM, mean = cfg5_problem(N, d)
    orc = GPOracle(d, np.exp(logM), mean=mean, white_noise=-12.0)
    orc.compute(X)
    rng = np.random.default_rng(5)
    q = rng.uniform(-5, 5, size=(8192, d))
    orc.predict(y, q, return_var=True)
    t0 = time.perf_counter(); orc.predict(y, q, return_var=True); dt = time.perf_counter() - t0
    nblk = int(max(1, min(64, target_s / max(dt, 1e-6))))
    t0 = time.perf_counter()
    for _ in range(nblk):
        orc.predict(y, q, return_var=True)
    dt = time.perf_counter() - t0
    return {"value": nblk * 8192 / dt, "unit": "evals/s", "cores": __import__("bench").blas_threads(), "kind": "port",
            "sample": "%d of the 1e8 queries through oracle.GPOracle.predict(return_var=True) in 8192-query blocks (cho_solve, all "
                      "BLAS threads), %.1f s" % (nblk * 8192, dt)}


def run_cfg5_reference(args, bench):
    cb = cfg5_reference(args.N, args.d, target_s=min(60.0, 6.0 * max(1, args.steps)))
    bench.emit_line({"impl": "reference", "metric": cfg5_metric(args.N, args.d), "value": cb["value"], "unit": "evals/s", "n_gpus": args.gpus,
                     "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": "strong",
                     "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg5_config(args.N, args.d), "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})


GPU = {"cfg1": run_cfg1, "cfg2": run_cfg2, "cfg4": run_cfg4, "cfg5": run_cfg5}
REFERENCE = {"cfg1": run_cfg1_reference, "cfg2": run_cfg2_reference, "cfg4": run_cfg4_reference, "cfg5": run_cfg5_reference}
