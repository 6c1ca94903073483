#!/usr/bin/env python
"""bench.py -- HMC leapfrog chain-steps/s on BASELINE.json configs[1]:
linear regression Y ~ Normal(X*beta, 1), N = 1 000 000 rows, M = 64 columns, 4096 chains per GPU, HMC isteps = 2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : libsimplemcmc_cuda.so through the host mirror of the reference API.
            value = device-resident throughput (CUDA events inside the library, on its launching stream);
            e2e   = one complete simpleHMC call with HOST buffers (tape lowering, H2D of X/Y from pinned memory,
                    K steps, D2H of the draws) timed with the wall clock.
            labelled_extras (never the headline): the FP64 DMMA kernel the int8 path replaces, the sufficient-statistics
            rewrite, configs[0] with 2^20 chains per GPU (the north-star model), and for --gpus N > 1 the strong-scaling
            reading of configs[1] (4096 chains in total) and configs[2] with its rows sharded over the ranks.
reference : the reference cannot run here (Julia <= 0.2 source, no Julia in the image), so this arm times the
            C++ restatement of its generated closure + simpleHMC loop (oracle/cpu_ref.cpp, kind "port") on all host
            cores, on a bounded sample of the same workload (one chain per core, few steps).
One step = one HMC iteration (isteps leapfrogs) of every chain.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

N_ROWS, N_COLS, CHAINS_PER_GPU, ISTEPS = 1_000_000, 64, 4096, 2
STEPSIZE = 5e-4   # accept rate ~0.83 at N=1e6, M=64 (tools/calibrate_stepsize.py, gpurun_out/calib.log)
MODEL = """
    beta ~ Normal(0, 1.0)
    resid = Y - X * beta
    resid ~ Normal(0, 1.0)
"""
METRIC = "HMC leapfrog chain-steps/sec (all chains)"
UNIT = "leapfrog chain-steps/s"
FP64_PEAK_TFLOPS_FALLBACK = 37.15
I8_PEAK_TOPS = 4580.0     # tcgen05.mma kind::i8 issue rate measured on this pool's B200 (tools/i8_umma_probe.cu, profiles/i8_umma_probe_r01.jsonl)
I8_SLICE_PAIRS = 28       # slice-pair products per FP64 contraction (seven balanced 8-bit slices, pairs with i + j <= 6)
LOGISTIC = """
    vars ~ Normal(0, 1.0)
    prob = 1 / (1. + exp(X * vars))
    Y ~ Bernoulli(prob)
"""
CFG1_N, CFG1_M, CFG1_CHAINS = 1000, 10, 1 << 20
CFG3_N, CFG3_M = 10_000_000, 10


def make_data(n=N_ROWS, m=N_COLS, seed=2, pinned=False):
    """generator of examples/linear.jl:6-11 at the cfg2 size, NumPy default_rng(2)"""
    rng = np.random.default_rng(seed)
    if pinned:
        import torch
        Xt = torch.empty((m, n), dtype=torch.float64, pin_memory=True)   # column-major N x M
        Yt = torch.empty((n,), dtype=torch.float64, pin_memory=True)
        X = Xt.numpy().T
        Y = Yt.numpy()
        hold = (Xt, Yt)
    else:
        X = np.empty((m, n)).T
        Y = np.empty(n)
        hold = None
    X[:, 0] = 1.0
    for j in range(1, m):
        X[:, j] = rng.standard_normal(n)
    beta0 = rng.standard_normal(m)
    Y[:] = X @ beta0 + rng.standard_normal(n)
    return X, Y, beta0, hold


def posterior_mode(X, Y):
    m = X.shape[1]
    return np.linalg.solve(X.T @ X + np.eye(m), X.T @ Y)


class ClockSampler(object):
    """one long-lived `nvidia-smi -lms` process (the recipe of B200_PROFILING.md): started before the warm-up, so that
    no process is spawned inside the timed region, killed after it"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.t_start = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=timestamp," + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def mark(self):
        """samples taken after this call belong to the timed region"""
        self.t_start = time.time()

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.time()
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        rows = []
        import datetime
        for ln in out.strip().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except Exception:
                ts = None
            rows.append((ts, parts[1:]))
        timed = [r for ts, r in rows if ts is not None and self.t_start is not None and self.t_start - 0.05 <= ts <= t_end + 0.05]
        use = timed if timed else [r for _, r in rows[-5:]]
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in use if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in use)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(use[0][1]), "reasons": reasons,
                "samples_in_timed_region": len(timed), "samples_total": len(rows)}


def fp64_peak():
    p = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
    try:
        d = json.load(open(p))
        return max(v for k, v in d.items() if k.startswith("dmma_") and isinstance(v, float)), \
            "measured: tools/fp64_peak.cu DMMA issue-rate microbenchmark on this pool's B200 (profiles/fp64_peak_r01.json); MEASURED_PEAKS.json has no FP64 figure"
    except Exception:
        return FP64_PEAK_TFLOPS_FALLBACK, "fallback: 37.15 TFLOP/s (earlier DMMA microbenchmark)"


def cpu_baseline(X, Y, mode, steps, threads=0):
    """bounded sample of the same workload on the host cores: one chain per core"""
    from oracle import cpu_ref
    cores = threads or cpu_ref.max_threads()
    rng = np.random.default_rng(7)
    init = mode[None, :] + 1e-3 * rng.standard_normal((cores, X.shape[1]))
    t0 = time.time()
    r = cpu_ref.hmc("linear", X, Y, init, steps, ISTEPS, STEPSIZE, seed=11, threads=cores)
    dt = time.time() - t0
    extra = {}
    try:    # the NumPy restatement of the generated closure (BLAS gemv + vectorised logpdf), one thread
        from threadpoolctl import threadpool_limits
        from oracle.refmodel import generate_model_function
        ref = generate_model_function(MODEL, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(X.shape[1]))
        with threadpool_limits(limits=1):
            ref(mode)
            t1 = time.time()
            for _ in range(3):
                ref(mode)
            extra = {"numpy_1thread_evals_per_s": 3.0 / (time.time() - t1),
                     "numpy_1thread_note": "oracle/refmodel.py closure (statement list evaluated with NumPy, BLAS limited to one thread): "
                                           "logp+grad evaluations/s of one chain = its leapfrog chain-steps/s"}
    except Exception as e:      # the baseline above stands on its own
        extra = {"numpy_1thread_error": str(e)[:200]}
    try:    # context: what a CPU code that BATCHED the chains into DGEMMs would reach on these cores (not the reference's algorithm)
        cb = 64
        Bm = np.ascontiguousarray((mode[None, :] + 1e-3 * rng.standard_normal((cb, X.shape[1]))).T)
        def batched():
            E = X @ Bm
            E -= Y[:, None]
            return -0.5 * np.einsum("ij,ij->j", E, E), X.T @ E
        batched()
        t2 = time.time()
        for _ in range(3):
            batched()
        extra["numpy_batched_dgemm_chain_evals_per_s"] = 3.0 * cb / (time.time() - t2)
        extra["numpy_batched_dgemm_note"] = ("context only: 64 chains per call as two DGEMMs (X @ B, X' @ W) with NumPy's BLAS on all host "
                                             "threads it uses by default; logp+grad evaluations/s summed over the chains")
    except Exception as e:
        extra["numpy_batched_dgemm_error"] = str(e)[:200]
    return {"value": r["n_leapfrogs"] / dt, "unit": UNIT, "cores": cores, "kind": "port", **extra,
            "sample": "%d chains (one per host core) x %d HMC steps x %d leapfrogs at the full N=%d, M=%d (oracle/cpu_ref.cpp, "
                      "statement-by-statement C++ restatement of the generated closure), %.1f s" % (cores, steps, ISTEPS, X.shape[0], X.shape[1], dt),
            "caveat": "the reference's one-chain-at-a-time algorithm on the host cores: its two mat-vecs are plain -O3 loops here, not the "
                      "BLAS dgemv Julia called, and %d chains are not 4096 (a rate, not the same job); a CPU code that batched the chains "
                      "into DGEMMs is faster than this arm by the factor numpy_batched_dgemm_chain_evals_per_s / value, measured in the same run" % cores}, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, Y, _, _ = make_data()
    mode = posterior_mode(X, Y)
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(X, Y, mode, 1)
    t0 = time.time()
    base, dt = cpu_baseline(X, Y, mode, max(1, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: linear regression N=1e6 x M=64, HMC isteps=2 (bounded sample: one chain per host core)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def glm_path_extra(K, W, i8):
    """configs[1] through the other GLM kernel, in a child process (SMC_GLM_I8 is read once per process)"""
    import subprocess
    note = ("SMC_GLM_I8=0: the FP64 DMMA kernel (mma.sync.m8n8k4.f64, csrc/glm_fused.cu k_glm) that the int8 tensor-core path "
            "replaces for Normal GLM statements with 33-64 columns and >= 32 chains; same model, data, sampler settings") if not i8 else \
           ("SMC_GLM_I8=1: Normal GLM statement from seven balanced 8-bit slices per operand on tcgen05.mma kind::i8 "
            "(int32 accumulators in TMEM), csrc/glm_i8.cu")
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "i8_bench_extra.py"), str(K), str(W)],
                             capture_output=True, text=True, timeout=300, env=dict(os.environ, SMC_GLM_I8="1" if i8 else "0"))
        lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if out.returncode != 0 or not lines:
            return {"error": (out.stderr or out.stdout)[-400:], "note": note}
        res = json.loads(lines[-1])
        res["note"] = note
        return res
    except Exception as e:      # the extra must never take the benchmark line down
        return {"error": str(e)[:400], "note": note}


def run_ours(args):
    # stdout carries exactly one line (the JSON record): anything libraries print while the benchmark runs -- e.g. the
    # "NCCL version ..." banner of the first communicator -- is sent to stderr at file-descriptor level
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line))
        sys.stdout.flush()


def _allmax(x, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _allsum(x, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def _same_on_all_ranks(arr, world):
    import torch
    import torch.distributed as dist
    if world == 1:
        return True
    t = torch.tensor(np.ascontiguousarray(arr, dtype=np.float64).ravel(), device="cuda")
    hi, lo = t.clone(), t.clone()
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return bool(torch.equal(hi, lo))


def cfg1_extra(api, ctx, local, rank, world, barrier):
    """configs[0] (examples/linear.jl:6-18,29: N = 1000, M = 10, HMC isteps = 2, stepsize = 0.05) with 2^20 chains per GPU:
    the model BASELINE.json's 1e9 leapfrog chain-steps/s target is stated on"""
    rng = np.random.default_rng(1)
    X = np.column_stack([np.ones(CFG1_N), rng.standard_normal((CFG1_N, CFG1_M - 1))])
    Y = X @ rng.standard_normal(CFG1_M) + rng.standard_normal(CFG1_N)
    mode = np.linalg.solve(X.T @ X + np.eye(CFG1_M), X.T @ Y)
    C, steps = CFG1_CHAINS, 100
    low, dev = api._build(MODEL, [("beta", np.zeros(CFG1_M))], dict(X=X, Y=Y), True, C, local, "assign", comm=ctx)
    try:
        kw = dict(isteps=2, stepsize=0.05, chain_offset=rank * C, store_draws=False, want=())
        dev.run("hmc", mode, C, 5, 0, seed=1, **kw)
        barrier()
        raw = dev.run("hmc", mode, C, steps, 0, seed=2, **kw)
        barrier()
        dev.eval(np.tile(mode, (C, 1)))
        ms_eval, ms_glm = dev.eval_timed(5, False)
    finally:
        dev.close()
    ms = _allmax(raw["device_ms"], world)
    leap = _allsum(raw["n_leapfrogs"], world)
    peak, _ = fp64_peak()
    tf = 4.0 * CFG1_N * CFG1_M * C / (ms_glm * 1e-3) / 1e12
    return {"value": leap / (ms * 1e-3), "unit": UNIT, "chains_per_gpu": C, "steps": steps, "accept_rate": float(np.mean(raw["accept_rate"])),
            "ms_per_eval": ms_eval, "glm_ms_per_launch": ms_glm,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                         "note": "k_glm (FP64 DMMA), F = 4*N*M*C algorithmic flop; X (80 KB) stays in L2 / shared memory, so there is no HBM bound"},
            "note": "configs[0] examples/linear.jl (N=1000, M=10, HMC isteps=2, stepsize=0.05) with 2^20 chains per GPU, chains sharded "
                    "over the GPUs; device-resident, CUDA events in the library, max over ranks"}


SLEEPSTUDY = """
    r_bar = (intercept + submatrix * rand_int) + (slope + submatrix * rand_slope) .* days
    rand_int ~ Normal(0,100)
    rand_slope ~ Normal(0,100)
    resid = r_bar - reaction
    resid ~ Normal(0,10)
"""
HIERARCHICAL = """
    mu ~ Normal(0, 1)
    sigma ~ Weibull(2, 1)
    beta ~ Normal(oneD * mu, oneD * sigma)
    effect = (beta[ll,:] .* X) * oneL
    prob = 1. / ( 1. + exp(- effect) )
    Y ~ Bernoulli(prob)
"""


def nuts_extra(api, ctx, local):
    """configs[3] (examples/sleepstudy.jl:25-45, examples/hierarchical.jl:7-37 with synthetic data of those shapes): simpleNUTS,
    16 384 chains as BASELINE.json states, a BOUNDED 400 iterations (the stated 2000 take 75 s per model: tools/nuts_baseline_run.py,
    profiles/nuts_baseline_size_r02.jsonl); nadapt is a literal 1000 in the reference (src/samplers.jl:214), so all of these
    iterations adapt the step size and the number is NOT the one of the full-size run"""
    C, steps = 16384, 400
    rng = np.random.default_rng(6)
    nsub, ndays = 18, 10
    days = np.tile(np.arange(ndays, dtype=np.float64), nsub)
    sub = np.repeat(np.arange(nsub), ndays)
    submatrix = np.zeros((nsub * ndays, nsub))
    submatrix[np.arange(nsub * ndays), sub] = 1.0
    ri, rs = 25 * rng.standard_normal(nsub), 6 * rng.standard_normal(nsub)
    reaction = 250 + 10 * days + ri[sub] + rs[sub] * days + 25 * rng.standard_normal(nsub * ndays)
    rng = np.random.default_rng(4)
    N, D, L = 50, 4, 5
    mu0, sigma0 = rng.standard_normal((1, L)), rng.random((1, L))
    beta0 = rng.standard_normal((D, L)) * sigma0 + mu0
    ll = rng.integers(1, D + 1, N)
    Xh = rng.standard_normal((N, L))
    Yh = (rng.random(N) < 1.0 / (1.0 + np.exp(-(beta0[ll - 1, :] * Xh).sum(axis=1)))).astype(np.float64)
    cases = [("sleepstudy", SLEEPSTUDY, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))],
              dict(days=days, reaction=reaction, submatrix=submatrix), "assign"),
             ("hierarchical", HIERARCHICAL, [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))],
              dict(oneD=np.ones(D), oneL=np.ones(L), ll=ll, X=Xh, Y=Yh), "add")]
    out = {"chains": C, "iterations": steps, "unit": "leapfrogs/s",
           "note": "simpleNUTS, device-resident, CUDA events in the library; bounded sample of configs[3] (see the docstring of nuts_extra)"}
    for name, model, init, data, adj in cases:
        low, dev = api._build(model, init, data, True, C, local, adj, comm=ctx)
        try:
            dev.run("nuts", np.array(low.init), C, 4, 0, seed=3, store_draws=False, want=())
            raw = dev.run("nuts", np.array(low.init), C, steps, steps // 2, seed=1, store_draws=False, want=("epsilon", "jmax"))
        finally:
            dev.close()
        rounds = max(int(raw["n_evals"]) - 1, 1)
        out[name] = {"value": raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3), "rounds": rounds,
                     "launches_per_round": raw["n_kernel_launches"] / rounds, "active_lane_fraction": raw["n_leapfrogs"] / (rounds * C),
                     "epsilon_mean_last": float(raw["epsilon"][-1].mean()), "mean_tree_depth_second_half": float(raw["jmax"][steps // 2:].mean()),
                     "accept_rate": float(raw["accept_rate"].mean())}
    return out


def strong_scaling_extra(api, ctx, local, rank, world, data, mode, K, W, barrier):
    """configs[1] read as strong scaling: 4096 chains IN TOTAL, 4096 / N per GPU (BASELINE.json: "4096 chains, chain-sharded
    at 1/2/4/8 B200")"""
    C = CHAINS_PER_GPU // world
    rng = np.random.default_rng(100)
    init = (mode[None, :] + 1e-3 * rng.standard_normal((CHAINS_PER_GPU, N_COLS)))[rank * C:(rank + 1) * C]
    low, dev = api._build(MODEL, [("beta", np.zeros(N_COLS))], data, True, C, local, "assign", comm=ctx)
    try:
        kw = dict(isteps=ISTEPS, stepsize=STEPSIZE, chain_offset=rank * C, store_draws=False, want=())
        dev.run("hmc", init, C, W, 0, seed=5, **kw)
        barrier()
        raw = dev.run("hmc", init, C, K, 0, seed=6, **kw)
        barrier()
    finally:
        dev.close()
    ms = _allmax(raw["device_ms"], world)
    leap = _allsum(raw["n_leapfrogs"], world)
    return {"value": leap / (ms * 1e-3), "unit": UNIT, "chains_total": C * world, "chains_per_gpu": C, "steps": K, "ms_per_step": ms / K,
            "scaling": "strong", "accept_rate": float(np.mean(raw["accept_rate"])),
            "note": "the same model and data with 4096 chains in total; no collective while sampling, X replicated"}


def row_sharded_extra(api, capi, parallel, local, rank, world, barrier):
    """configs[2] (examples/binomial.jl:13-17): logistic GLM, N = 1e7 rows sharded over the ranks, every rank advances the same
    chains; (L, G) of the likelihood statement is summed over the ranks every evaluation -- over peer memory by one
    reduce + exchange + publish kernel (csrc/smc_p2p.cu), and by the in-stream NCCL all-reduce for comparison"""
    import torch
    rng = np.random.default_rng(3)
    N, M = CFG3_N, CFG3_M
    r0, r1 = parallel.row_shard(N, rank, world)
    Xs = np.empty((M, r1 - r0)).T                 # column-major shard
    b0 = None
    # every rank draws the same stream column by column and keeps its rows (no rank ever holds more than its shard, rank 0 also
    # keeps the full matrix for the single-GPU check)
    Xfull = np.empty((M, N)).T if rank == 0 else None
    col = np.ones(N)
    for j in range(M):
        if j > 0:
            col = rng.standard_normal(N)
        Xs[:, j] = col[r0:r1]
        if rank == 0:
            Xfull[:, j] = col
    b0 = rng.standard_normal(M)
    u = rng.random(N)
    out = {"rows": N, "columns": M, "gpus": world, "model": "examples/binomial.jl: vars ~ Normal(0,1); Y ~ Bernoulli(1/(1+exp(X*vars)))"}
    if rank == 0:
        Yfull = (u < 1.0 / (1.0 + np.exp(Xfull @ b0))).astype(np.float64)
    else:
        Yfull = np.empty(N)
    t = torch.from_numpy(Yfull).cuda()
    import torch.distributed as dist
    dist.broadcast(t, src=0)
    Ys = t.cpu().numpy()[r0:r1].copy()
    del t
    eps = 0.1 / np.sqrt(N / 1000)
    ref = {}
    for exch in ("p2p", "nccl"):
        ctx = capi.Context(local)
        parallel.init_library_comm(ctx, rank, world)
        if exch == "p2p":
            parallel.init_p2p(ctx, rank, world)
        for C in (2, 64):
            betas = b0[None, :] + 1e-3 * np.random.default_rng(9).standard_normal((C, M))
            low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=Xs, Y=Ys), True, C, local, "assign", shard="rows", comm=ctx)
            try:
                lp, g = dev.eval(betas)
                row = {"identical_logp_grad_on_all_ranks": _same_on_all_ranks(np.concatenate([lp, g.ravel()]), world)}
                if rank == 0:
                    if C not in ref:      # the whole data on rank 0's GPU alone
                        ctx1 = capi.Context(local)
                        low1, dev1 = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=Xfull, Y=Yfull), True, C, local, "assign", comm=ctx1)
                        ref[C] = dev1.eval(betas)
                        dev1.run("hmc", betas, C, 3, 0, isteps=2, stepsize=eps, seed=1, store_draws=False, want=())
                        r1g = dev1.run("hmc", betas, C, 20, 0, isteps=2, stepsize=eps, seed=2, store_draws=False, want=())
                        ref[(C, "ms")] = r1g["device_ms"] / 20
                        dev1.close()
                    lp1, g1 = ref[C]
                    row["logp_relerr_vs_1gpu"] = float(np.max(np.abs(lp - lp1) / np.abs(lp1)))
                    row["grad_relerr_vs_1gpu"] = float(np.max(np.abs(g - g1)) / np.max(np.abs(g1)))
                    row["hmc_ms_per_step_1gpu_all_rows"] = ref[(C, "ms")]
                steps = 20
                dev.run("hmc", betas, C, 3, 0, isteps=2, stepsize=eps, seed=1, store_draws=False, want=())
                barrier()
                raw = dev.run("hmc", betas, C, steps, 0, isteps=2, stepsize=eps, seed=2, store_draws=False, want=())
                barrier()
                ms = _allmax(raw["device_ms"], world)
                row.update(hmc_ms_per_step=ms / steps, leapfrog_us=1e3 * ms / (steps * 2), accept_rate=float(raw["accept_rate"].mean()),
                           leapfrog_chain_steps_per_s=C * steps * 2 / (ms * 1e-3),
                           chains_identical_on_all_ranks_after_hmc=_same_on_all_ranks(raw["final_state"], world))
                if rank == 0:
                    row["speedup_vs_1gpu"] = ref[(C, "ms")] / (ms / steps)
                out["%s_chains%d" % (exch, C)] = row
            finally:
                dev.close()
        barrier()
    out["note"] = ("rows / N per GPU, chain state replicated; per evaluation the (M + 1) x C block of likelihood sums crosses the ranks: "
                   "'p2p' = one kernel over NVLink peer memory (rank-ordered sums, identical on every rank), 'nccl' = in-stream "
                   "ncclAllReduce; ms per HMC step = 2 leapfrogs, CUDA events in the library, max over ranks")
    return out


def _run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from smc_pkg import load_package
    smc = load_package()
    from simplemcmc_jl_b200 import api, capi, parallel

    K, W = args.steps, max(args.warmup, 3)
    C = CHAINS_PER_GPU
    i8_on = os.environ.get("SMC_GLM_I8", "1") != "0"
    X, Y, beta0, hold = make_data(pinned=True)
    mode = posterior_mode(X, Y)
    rng = np.random.default_rng(100 + rank)
    init = mode[None, :] + 1e-3 * rng.standard_normal((C, N_COLS))
    data = dict(X=X, Y=Y)

    # ---------------- device-resident throughput: data bound once, K steps timed with CUDA events in the library
    ctx = capi.Context.default(local)
    low, dev = api._build(MODEL, [("beta", np.zeros(N_COLS))], data, True, C, local, "assign", comm=ctx)
    clocks = ClockSampler(local)
    dev.run("hmc", init, C, W, 0, isteps=ISTEPS, stepsize=STEPSIZE, seed=5, chain_offset=rank * C, store_draws=False,
            want=())                                                                   # warm-up steps

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    clocks.mark()
    t0 = time.time()
    raw = dev.run("hmc", init, C, K, 0, isteps=ISTEPS, stepsize=STEPSIZE, seed=6, chain_offset=rank * C, store_draws=False,
                  want=())
    barrier()
    wall = time.time() - t0
    clk = clocks.summary()
    dev_ms = _allmax(raw["device_ms"], world)
    leap = _allsum(raw["n_leapfrogs"], world)
    value = leap / (dev_ms * 1e-3)

    # ---------------- roofline of the dominant kernel: CUDA events around each launch (on the library's stream), L2 flushed between
    dev.eval(init)
    ms_eval, ms_glm = dev.eval_timed(reps=5, flush_l2=True)
    flops = 4.0 * N_ROWS * N_COLS * C
    peak, peak_src = fp64_peak()
    f64 = flops / (ms_glm * 1e-3) / 1e12
    alg_mb = (8.0 * N_ROWS * N_COLS + 8.0 * N_ROWS) / 1e6
    if i8_on:
        i8 = I8_SLICE_PAIRS * f64
        roof = {"bound": "tensor", "achieved": i8, "peak": I8_PEAK_TOPS, "unit": "TFLOP/s", "frac": i8 / I8_PEAK_TOPS, "traffic": None,
                "kernel": "k_glm_i8 (tcgen05.mma kind::i8, int32 accumulators in TMEM; fused X*B -> Normal logpdf+adjoint -> X'*D from "
                          "seven balanced 8-bit slices per operand, FP64 recombination)",
                "unit_note": "tera int8 multiply-add operations per second (x2), not floating point: %d slice-pair products of the shape of "
                             "each FP64 contraction" % I8_SLICE_PAIRS,
                "ops_per_launch": I8_SLICE_PAIRS * flops, "ms_per_launch": ms_glm, "ms_per_eval": ms_eval,
                "peak_source": "measured: tcgen05.mma kind::i8 issue rate on this pool's B200, every SM issuing back to back "
                               "(tools/i8_umma_probe.cu, profiles/i8_umma_probe_r01.jsonl); MEASURED_PEAKS.json has no int8 figure",
                "fp64_equivalent": {"achieved": f64, "peak": peak, "unit": "TFLOP/s", "frac": f64 / peak, "flops_per_launch": flops,
                                    "peak_source": peak_src,
                                    "note": "F = 4*N*M flops per chain-step (SURVEY.md 8d) against the FP64 tensor (DMMA) peak, the bound of "
                                            "the kernel this path replaces: above 1 because the contraction no longer runs on the FP64 pipe"},
                "note": "algorithmic HBM bytes per launch = 8*N*M + 8*N = %.0f MB (the sliced X is 8 bytes per element like the FP64 matrix)" % alg_mb}
        tr = os.path.join(ROOT, "profiles", "glm_i8_traffic_r02.json")
        try:    # the pipe this kernel is actually bound by (FP64 instructions do not issue while the tensor core works, DESIGN.md
                # section 8): tensor + FP64 cycles of the SM's shared pipe, from the committed ncu capture of the same kernel
            prof = json.load(open(os.path.join(ROOT, "profiles", "glm_i8_v14_ncu_r02.json")))[0]
            roof["shared_pipe_ncu"] = {
                "busy_frac": prof["sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"] / 100.0,
                "imma_frac": prof["sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active"] / 100.0,
                "fp64_frac": prof["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"] / 100.0,
                "source": "profiles/glm_i8_v14_ncu_r02.json (ncu --set full of this kernel at N = 2e5 x 1024 chains), not measured in this run: "
                          "sm__pipe_shared_cycles_active = int8 tensor cycles + FP64 cycles, which exclude each other on this SM"}
        except Exception:
            pass
    else:
        roof = {"bound": "tensor", "achieved": f64, "peak": peak, "unit": "TFLOP/s", "frac": f64 / peak, "traffic": None,
                "kernel": "k_glm (FP64 DMMA mma.sync.m8n8k4, fused X*B -> Normal logpdf+adjoint -> X'*D)",
                "flops_per_launch": flops, "ms_per_launch": ms_glm, "ms_per_eval": ms_eval,
                "peak_source": peak_src, "note": "FP64 tensor (DMMA) roofline: F = 4*N*M flops per chain-step (SURVEY.md 8d); "
                "algorithmic HBM bytes per launch = 8*N*M + 8*N = %.0f MB" % alg_mb}
        tr = os.path.join(ROOT, "profiles", "glm_traffic_r01.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr))["dram_bytes_per_launch"]
            roof["traffic_source"] = "read from the committed `ncu --set full` capture %s (dram__bytes_read.sum + dram__bytes_write.sum of one " \
                                     "launch at this size), not measured in this run" % os.path.relpath(tr, ROOT)
        except Exception:
            pass
    dev.close()

    # ---------------- labelled extra (NOT the headline): the same model and sampler with the opt-in
    # sufficient-statistics rewrite (X'X, X'y, y'y hoisted once; 2*M^2 instead of 4*N*M flops per chain-step)
    low2, dev2 = api._build(MODEL, [("beta", np.zeros(N_COLS))], data, True, C, local, "assign", comm=ctx, sufficient_stats=True)
    dev2.run("hmc", init, C, W, 0, isteps=ISTEPS, stepsize=STEPSIZE, seed=5, chain_offset=rank * C, store_draws=False, want=())
    barrier()
    raw2 = dev2.run("hmc", init, C, 50 * K, 0, isteps=ISTEPS, stepsize=STEPSIZE, seed=6, chain_offset=rank * C, store_draws=False,
                    want=())
    extra = {"sufficient_statistics_rewrite": {
        "value": _allsum(raw2["n_leapfrogs"], world) / (_allmax(raw2["device_ms"], world) * 1e-3), "unit": UNIT, "steps": 50 * K,
        "accept_rate": float(np.mean(raw2["accept_rate"])),
        "note": "opt-in `sufficient_stats=True`: a different algorithm from the reference's generated code (Gram matrix hoisted "
                "into the let-header); reported beside the headline, never instead of it; roofline.achieved stays F = 4*N*M"}}
    dev2.close()

    # ---------------- end to end through the public API with host buffers (pinned): lowering + H2D + K steps + D2H draws
    def e2e_call(steps, seed):
        return smc.simpleHMC(MODEL, steps=steps, burnin=0, isteps=ISTEPS, stepsize=STEPSIZE, chains=C, seed=seed, data=data,
                             device=local, comm=ctx, chain_offset=rank * C, beta=mode)
    e2e_call(2, 3)
    walls, leaps = [], []
    for rep in range(3):      # the host side of a shared box is noisy (0.75 - 2.1 s seen for the same call): three complete calls
        barrier()
        t1 = time.time()
        res = e2e_call(K, 4 + rep)
        if world > 1:   # the one collective of a chain-sharded run: moments for R-hat
            mom = torch.tensor(np.stack([res.misc["chain_mean"].sum(0), (res.misc["chain_mean"] ** 2).sum(0),
                                         res.misc["chain_m2"].sum(0)]), device="cuda")
            dist.all_reduce(mom)
        barrier()
        walls.append(_allmax(time.time() - t1, world))
        leaps.append(_allsum(res.misc["n_leapfrogs"], world))
    mid = sorted(range(3), key=lambda k: walls[k])[1]
    h2d = X.nbytes + Y.nbytes + mode.nbytes
    d2h = sum(v.nbytes for v in res.params.values()) + res.loglik.nbytes + res.accept.nbytes
    e2e = {"value": leaps[mid] / walls[mid], "unit": UNIT, "h2d_bytes_per_step": h2d / K,
           "d2h_bytes_per_step": d2h / K, "wall_s": walls[mid], "wall_s_all_calls": walls,
           "what": "one simpleHMC(model, steps=K, chains=4096, ...) call per rank with host (pinned) X, Y: lowering, tape compile, "
                   "H2D bind, K steps, D2H of all draws; wall clock (max over ranks), median of 3 calls (all in wall_s_all_calls)"}

    # ---------------- labelled extras on the other configs (each one after every headline measurement is complete; an extra
    # that fails records its error string and never takes the line down)
    def guarded(name, fn):
        try:
            extra[name] = fn()
        except Exception as e:
            extra[name] = {"error": ("%s: %s" % (type(e).__name__, e))[:400]}
        try:
            barrier()
        except Exception:
            pass
    guarded("cfg1_2p20_chains", lambda: cfg1_extra(api, ctx, local, rank, world, barrier))
    if world == 1:
        guarded("nuts_cfg4_16k_chains", lambda: nuts_extra(api, ctx, local))
    if world > 1:
        guarded("strong_scaling_cfg2", lambda: strong_scaling_extra(api, ctx, local, rank, world, data, mode, K, W, barrier))
        guarded("row_sharded_cfg3", lambda: row_sharded_extra(api, capi, parallel, local, rank, world, barrier))
    # the FP64 DMMA kernel the int8 path replaced (or vice versa when the run itself was started with SMC_GLM_I8=0), in a
    # child process because the switch is read once per process
    if rank == 0 and world == 1:
        extra["fp64_dmma_glm_path" if i8_on else "int8_tensor_core_glm_path"] = glm_path_extra(K, W, not i8_on)

    if rank == 0:
        cpu, _ = cpu_baseline(X, Y, mode, 24)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "configs[1]: linear regression Y~Normal(X*beta,1), N=1e6 x M=64 (numpy default_rng(2)), "
                                       "simpleHMC isteps=2, 4096 chains per GPU, chains sharded over GPUs, X replicated",
                           "chains_per_gpu": C, "isteps": ISTEPS, "stepsize": STEPSIZE, "accept_rate": float(np.mean(raw["accept_rate"])),
                           "init": "posterior mode + 1e-3*N(0,1) per chain (computed outside the timed region)",
                           "glm_kernel": "k_glm_i8" if i8_on else "k_glm",
                           "l2": "inputs larger than L2: X is 512 MB vs 126 MB L2; roofline launches additionally flush L2 (256 MB memset)",
                           "wall_s_timed_region": wall},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(raw["n_kernel_launches"]), "roofline": roof, "cpu_baseline": cpu,
                "labelled_extras": extra}
    else:
        line = None
    if world > 1:
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
