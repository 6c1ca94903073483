"""Secondary measurements for the other BASELINE.json configs (device-resident, CUDA-event timed inside the library):
cfg1 linear N=1000 M=10 with many chains, cfg3 logistic (one GPU, all rows), cfg4 NUTS on the sleepstudy model,
cfg5 Ornstein-Uhlenbeck RWM/HMC/NUTS sweep.  Writes one JSON object to stdout."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
from util import HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, ou_data, sleepstudy_data

out = {}
which = sys.argv[1:] or ["cfg1", "cfg2s", "cfg3", "cfg4", "cfg5"]

if "cfg1" in which:
    X, Y, _ = linear_data(1000, 10, 1)
    mode = np.linalg.solve(X.T @ X + np.eye(10), X.T @ Y)
    rows = []
    for C in (1, 1024, 65536, 1 << 20):
        low, dev = api._build(LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y), True, C, 0, "assign")
        steps = 2000 if C == 1 else (200 if C <= 65536 else 120)
        dev.run("hmc", mode, C, 5, 0, isteps=2, stepsize=0.05, seed=1, store_draws=False, want=())
        raw = dev.run("hmc", mode, C, steps, 0, isteps=2, stepsize=0.05, seed=2, store_draws=False, want=())
        dev.eval(np.tile(mode, (C, 1)))
        ms_eval, ms_glm = dev.eval_timed(5, False)
        rows.append(dict(chains=C, steps=steps, leapfrog_chain_steps_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         accept=float(raw["accept_rate"].mean()), ms_eval=ms_eval, ms_glm=ms_glm,
                         glm_tflops_algorithmic=4.0 * 1000 * 10 * C / (ms_glm * 1e-3) / 1e12, launches=raw["n_kernel_launches"]))
        dev.close()
    out["cfg1_linear_N1000_M10_hmc"] = rows

if "cfg2s" in which:
    # configs[1] with the opt-in sufficient-statistics rewrite (labelled extra: not the reference's arithmetic)
    import bench
    X, Y, _, _ = bench.make_data()
    mode = bench.posterior_mode(X, Y)
    rows = []
    for C in (4096, 65536, 1 << 20):
        low, dev = api._build(bench.MODEL, [("beta", np.zeros(64))], dict(X=X, Y=Y), True, C, 0, "assign", sufficient_stats=True)
        init = mode[None, :] + 1e-3 * np.random.default_rng(1).standard_normal((C, 64))
        steps = 400 if C <= 65536 else 40
        dev.run("hmc", init, C, 5, 0, isteps=2, stepsize=bench.STEPSIZE, seed=1, store_draws=False, want=())
        raw = dev.run("hmc", init, C, steps, 0, isteps=2, stepsize=bench.STEPSIZE, seed=2, store_draws=False, want=())
        rows.append(dict(chains=C, steps=steps, leapfrog_chain_steps_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         accept=float(raw["accept_rate"].mean()), launches=raw["n_kernel_launches"]))
        dev.close()
    out["cfg2_linear_N1e6_M64_hmc_sufficient_statistics"] = rows

if "cfg3" in which:
    rng = np.random.default_rng(3)
    N, M = 10_000_000, 10
    X = np.empty((M, N)).T
    X[:, 0] = 1.0
    for j in range(1, M):
        X[:, j] = rng.standard_normal(N)
    b0 = rng.standard_normal(M)
    Y = (rng.random(N) < 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
    rows = []
    for C in (1, 8, 64):
        low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
        init = b0[None, :] + 1e-3 * rng.standard_normal((C, M))
        kw = dict(isteps=2, stepsize=0.1 / np.sqrt(N / 1000), store_draws=False, want=())
        dev.run("hmc", init, C, 4, 0, seed=3, **kw)      # warm-up: buffers sized, evaluation and iteration graphs captured
        raw = dev.run("hmc", init, C, 40 if C <= 8 else 12, 0, seed=1, **kw)
        dev.eval(init)
        ms_eval, ms_glm = dev.eval_timed(5, True)
        bytes_alg = 8.0 * N * M + 8.0 * N
        rows.append(dict(chains=C, leapfrog_chain_steps_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         accept=float(raw["accept_rate"].mean()), ms_eval=ms_eval, ms_glm=ms_glm,
                         glm_gbs_algorithmic=bytes_alg / (ms_glm * 1e-3) / 1e9, glm_gbs_padded=(8.0 * N * 16 + 8.0 * N) / (ms_glm * 1e-3) / 1e9))
        dev.close()
    out["cfg3_logistic_N1e7_M10_hmc_1gpu"] = rows
    # the count variant of SURVEY.md 8(d): Y ~ Binomial(20, prob)
    Yb = rng.binomial(20, 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
    BINOM = "vars ~ Normal(0, 1.0)\nprob = 1 / (1. + exp(X * vars))\nY ~ Binomial(20, prob)"
    rows = []
    for C in (1, 64):
        low, dev = api._build(BINOM, [("vars", np.zeros(M))], dict(X=X, Y=Yb), True, C, 0, "assign")
        init = b0[None, :] + 1e-3 * rng.standard_normal((C, M))
        kw = dict(isteps=2, stepsize=0.02 / np.sqrt(N / 1000), store_draws=False, want=())
        dev.run("hmc", init, C, 4, 0, seed=3, **kw)
        raw = dev.run("hmc", init, C, 40 if C <= 8 else 12, 0, seed=1, **kw)
        dev.eval(init)
        ms_eval, ms_glm = dev.eval_timed(5, True)
        rows.append(dict(chains=C, leapfrog_chain_steps_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         accept=float(raw["accept_rate"].mean()), ms_eval=ms_eval, ms_glm=ms_glm,
                         glm_gbs_algorithmic=(8.0 * N * M + 8.0 * N) / (ms_glm * 1e-3) / 1e9))
        dev.close()
    out["cfg3_binomial_counts_N1e7_M10_hmc_1gpu"] = rows

if "cfg4" in which:
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    rows = []
    for C in (1024, 16384):
        low, dev = api._build(SLEEPSTUDY2, init, data, True, C, 0, "assign")
        steps = 60
        dev.run("nuts", np.array(low.init), C, 2, 0, seed=3, store_draws=False, want=())   # warm-up
        t0 = time.time()
        raw = dev.run("nuts", np.array(low.init), C, steps, 0, seed=1, store_draws=False, want=())
        rows.append(dict(chains=C, steps=steps, leapfrogs_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         iterations_per_s=C * steps / (raw["device_ms"] * 1e-3), evals=raw["n_evals"],
                         active_lane_fraction=raw["n_leapfrogs"] / (max(raw["n_evals"] - 1, 1) * C), wall_s=time.time() - t0,
                         mean_jmax=float(raw["jmax"].mean()), launches=raw["n_kernel_launches"]))
        dev.close()
    out["cfg4_sleepstudy_38params_nuts"] = rows
    # cfg4 (i): examples/hierarchical.jl, 30 parameters, gather adjoint as scatter-add (the corrected gradient, Appendix C Q2)
    data, (N, D, L) = hierarchical_data()
    init = [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))]
    rows = []
    for C in (1024, 16384):
        low, dev = api._build(HIERARCHICAL, init, data, True, C, 0, "add")
        steps = 60
        dev.run("nuts", np.array(low.init), C, 2, 0, seed=3, store_draws=False, want=())   # warm-up
        t0 = time.time()
        raw = dev.run("nuts", np.array(low.init), C, steps, 0, seed=1, store_draws=False, want=())
        rows.append(dict(chains=C, steps=steps, leapfrogs_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3),
                         iterations_per_s=C * steps / (raw["device_ms"] * 1e-3), evals=raw["n_evals"],
                         active_lane_fraction=raw["n_leapfrogs"] / (max(raw["n_evals"] - 1, 1) * C), wall_s=time.time() - t0,
                         mean_jmax=float(raw["jmax"].mean()), launches=raw["n_kernel_launches"]))
        dev.close()
    out["cfg4_hierarchical_30params_nuts"] = rows

if "cfg5" in which:
    x = ou_data(1_000_000, 5)
    init = [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)]
    rows = []
    for C in (1, 64, 4096):
        for kind, kw, steps in (("rwm", {}, 200), ("hmc", dict(isteps=2, stepsize=2e-5), 100), ("nuts", {}, 8)):
            if C == 4096 and kind == "nuts":
                continue
            low, dev = api._build(ORNSTEIN, init, dict(x=x), kind != "rwm", C, 0, "assign")
            dev.run(kind, np.array(low.init), C, 2, 0, seed=3, store_draws=False, want=(), **kw)      # warm-up: buffers sized, evaluation graph captured
            raw = dev.run(kind, np.array(low.init), C, steps, 0, seed=1, store_draws=False, want=(), **kw)
            dev.eval(np.tile(np.array(low.init), (C, 1)), gradient=kind != "rwm")
            ms_eval, ms_k = dev.eval_timed(20, True)       # eager evaluation (every kernel of the tape) and the AFFNORM launch alone, L2 flushed
            rows.append(dict(sampler=kind, chains=C, steps=steps, chain_steps_per_s=C * steps / (raw["device_ms"] * 1e-3),
                             evals_per_s=C * (raw["n_evals"] - 1) / (raw["device_ms"] * 1e-3), accept=float(raw["accept_rate"].mean()),
                             launches=raw["n_kernel_launches"], affnorm=low.n_affnorm, ms_eval_eager_l2_flushed=ms_eval, ms_affnorm_kernel=ms_k,
                             kernel_GBs_algorithmic=8e6 / (ms_k * 1e-3) / 1e9 if ms_k > 0 else None,
                             kernel_TFLOPs_20flop=20e6 * C / (ms_k * 1e-3) / 1e12 if ms_k > 0 else None))
            dev.close()
    out["cfg5_ornstein_T1e6"] = rows
    # labelled extra: the same model with the opt-in sufficient-statistics rewrite (3x3 Gram matrix of [x[t], x[t-1], 1])
    rows = []
    for C in (64, 4096, 65536):
        low, dev = api._build(ORNSTEIN, init, dict(x=x), True, C, 0, "assign", sufficient_stats=True)
        kw = dict(isteps=2, stepsize=1e-4)
        dev.run("hmc", np.array(low.init), C, 3, 0, seed=3, store_draws=False, want=(), **kw)
        raw = dev.run("hmc", np.array(low.init), C, 200, 0, seed=1, store_draws=False, want=(), **kw)
        rows.append(dict(sampler="hmc", chains=C, steps=200, evals_per_s=C * (raw["n_evals"] - 1) / (raw["device_ms"] * 1e-3),
                         accept=float(raw["accept_rate"].mean()), launches=raw["n_kernel_launches"]))
        dev.close()
    out["cfg5_ornstein_T1e6_sufficient_statistics"] = rows

print(json.dumps(out, indent=1))
