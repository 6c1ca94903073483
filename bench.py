#!/usr/bin/env python
"""bench.py -- HMC leapfrog chain-steps/s on BASELINE.json configs[1]:
linear regression Y ~ Normal(X*beta, 1), N = 1 000 000 rows, M = 64 columns, 4096 chains per GPU, HMC isteps = 2.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours      : libsimplemcmc_cuda.so through the host mirror of the reference API.
            value = device-resident throughput (CUDA events inside the library, on its launching stream);
            e2e   = one complete simpleHMC call with HOST buffers (tape lowering, H2D of X/Y from pinned memory,
                    K steps, D2H of the draws) timed with the wall clock.
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
        return max(v for k, v in d.items() if k.startswith("dmma_") and isinstance(v, float) and v < 100), \
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
    return {"value": r["n_leapfrogs"] / dt, "unit": UNIT, "cores": cores, "kind": "port", **extra,
            "sample": "%d chains (one per host core) x %d HMC steps x %d leapfrogs at the full N=%d, M=%d (oracle/cpu_ref.cpp, "
                      "statement-by-statement C++ restatement of the generated closure), %.1f s" % (cores, steps, ISTEPS, X.shape[0], X.shape[1], dt)}, dt


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


def int8_extra(K, W):
    import subprocess
    note = ("opt-in SMC_GLM_I8=1: Normal GLM statement from seven balanced 8-bit slices per operand on tcgen05.mma kind::i8 "
            "(int32 accumulators in TMEM); same statement list and the same 1e-12 parity bar as the headline, reported beside it "
            "until the parity suite has run with the switch set")
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "i8_bench_extra.py"), str(K), str(W)],
                             capture_output=True, text=True, timeout=300, env=dict(os.environ, SMC_GLM_I8="1"))
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
    from simplemcmc_jl_b200 import api, capi

    K, W = args.steps, max(args.warmup, 3)
    C = CHAINS_PER_GPU
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
    dev_ms = torch.tensor([raw["device_ms"]], dtype=torch.float64, device="cuda")
    leap = torch.tensor([float(raw["n_leapfrogs"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dev_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(leap, op=dist.ReduceOp.SUM)
    dev_ms, leap = float(dev_ms.item()), float(leap.item())
    value = leap / (dev_ms * 1e-3)

    # ---------------- roofline of the dominant kernel (k_glm): CUDA events around each launch, L2 flushed between
    dev.eval(init)
    ms_eval, ms_glm = dev.eval_timed(reps=5, flush_l2=True)
    flops = 4.0 * N_ROWS * N_COLS * C
    peak, peak_src = fp64_peak()
    achieved = flops / (ms_glm * 1e-3) / 1e12
    roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "kernel": "k_glm (FP64 DMMA mma.sync.m8n8k4, fused X*B -> Normal logpdf+adjoint -> X'*D)",
            "flops_per_launch": flops, "ms_per_launch": ms_glm, "ms_per_eval": ms_eval,
            "peak_source": peak_src, "note": "FP64 tensor (DMMA) roofline: F = 4*N*M flops per chain-step (SURVEY.md 8d); "
            "algorithmic HBM bytes per launch = 8*N*M + 8*N = %.0f MB" % ((8.0 * N_ROWS * N_COLS + 8.0 * N_ROWS) / 1e6)}
    tr = os.path.join(ROOT, "profiles", "glm_traffic_r01.json")
    if os.path.exists(tr):
        try:
            roof["traffic"] = json.load(open(tr))["dram_bytes_per_launch"]
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
    ms2 = torch.tensor([raw2["device_ms"]], dtype=torch.float64, device="cuda")
    leap2 = torch.tensor([float(raw2["n_leapfrogs"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        dist.all_reduce(leap2, op=dist.ReduceOp.SUM)
    extra = {"sufficient_statistics_rewrite": {
        "value": float(leap2.item()) / (float(ms2.item()) * 1e-3), "unit": UNIT, "steps": 50 * K,
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
    for rep in range(3):      # the host side of a shared box is noisy (0.75 - 2.1 s seen for the same call): best of three complete calls
        barrier()
        t1 = time.time()
        res = e2e_call(K, 4 + rep)
        if world > 1:   # the one collective of a chain-sharded run: moments for R-hat
            mom = torch.tensor(np.stack([res.misc["chain_mean"].sum(0), (res.misc["chain_mean"] ** 2).sum(0),
                                         res.misc["chain_m2"].sum(0)]), device="cuda")
            dist.all_reduce(mom)
        barrier()
        e2e_t = torch.tensor([time.time() - t1], dtype=torch.float64, device="cuda")
        e2e_leap = torch.tensor([float(res.misc["n_leapfrogs"])], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(e2e_leap, op=dist.ReduceOp.SUM)
        walls.append(float(e2e_t.item()))
        leaps.append(float(e2e_leap.item()))
    mid = sorted(range(3), key=lambda k: walls[k])[0]
    h2d = X.nbytes + Y.nbytes + mode.nbytes
    d2h = sum(v.nbytes for v in res.params.values()) + res.loglik.nbytes + res.accept.nbytes
    e2e = {"value": leaps[mid] / walls[mid], "unit": UNIT, "h2d_bytes_per_step": h2d / K,
           "d2h_bytes_per_step": d2h / K, "wall_s": walls[mid], "wall_s_all_calls": walls,
           "what": "one simpleHMC(model, steps=K, chains=4096, ...) call per rank with host (pinned) X, Y: lowering, tape compile, "
                   "H2D bind, K steps, D2H of all draws; wall clock (max over ranks), best of 3 calls (all in wall_s_all_calls)"}

    # ---------------- labelled extra (NOT the headline): the same model with the library's opt-in int8 tensor-core GLM path
    # (csrc/glm_i8.cu), in a child process because SMC_GLM_I8 is read once per process; it runs after
    # every measurement above is complete and cannot change them
    if rank == 0 and world == 1:
        extra["int8_tensor_core_glm_path"] = int8_extra(K, W)

    if rank == 0:
        cpu, _ = cpu_baseline(X, Y, mode, 24)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": "configs[1]: linear regression Y~Normal(X*beta,1), N=1e6 x M=64 (numpy default_rng(2)), "
                                       "simpleHMC isteps=2, 4096 chains per GPU, chains sharded over GPUs, X replicated",
                           "chains_per_gpu": C, "isteps": ISTEPS, "stepsize": STEPSIZE, "accept_rate": float(np.mean(raw["accept_rate"])),
                           "init": "posterior mode + 1e-3*N(0,1) per chain (computed outside the timed region)",
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
