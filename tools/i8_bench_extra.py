"""Child process of bench.py (labelled extra, never the headline): configs[1] through one of the two GLM kernels of the
library -- the int8 tensor-core path (csrc/glm_i8.cu, the default) or, with SMC_GLM_I8=0, the FP64 DMMA kernel it replaces
(the switch is read once per process, hence a process of its own).  Prints one JSON line: HMC leapfrog chain-steps/s, the
GLM kernel's event-timed launch, and the parity of logp / gradient of four chains against a long-double NumPy evaluation
of the full data.
    [SMC_GLM_I8=0] python tools/i8_bench_extra.py [steps] [warmup]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import bench
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api, capi

I8 = os.environ.get("SMC_GLM_I8", "1") != "0"
K = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C, M, N = bench.CHAINS_PER_GPU, bench.N_COLS, bench.N_ROWS
X, Y, beta0, hold = bench.make_data()
mode = bench.posterior_mode(X, Y)
init = mode[None, :] + 1e-3 * np.random.default_rng(100).standard_normal((C, M))
ctx = capi.Context.default(0)
low, dev = api._build(bench.MODEL, [("beta", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign", comm=ctx)
lp, g = dev.eval(init)
# parity of four chains against long double (row blocks keep the temporaries small)
nchk = 4
Bl = init[:nchk].T.astype(np.longdouble)                       # M x nchk
L = np.zeros(nchk, np.longdouble); G = np.zeros((M, nchk), np.longdouble)
for r0 in range(0, N, 100_000):
    Xl = X[r0:r0 + 100_000].astype(np.longdouble)
    d = Y[r0:r0 + 100_000].astype(np.longdouble)[:, None] - Xl @ Bl
    L += np.sum(-0.5 * d * d, axis=0)
    G += Xl.T @ d
L += -N * np.longdouble(0.91893853320467274178) + np.sum(-0.5 * Bl * Bl - np.longdouble(0.91893853320467274178), axis=0)
G += -Bl
err_l = float(np.max(np.abs((lp[:nchk] - L) / L)))
err_g = float(np.max(np.abs(g[:nchk].T - G) / np.max(np.abs(G), axis=0)))
ms_eval, ms_glm = dev.eval_timed(reps=5, flush_l2=True)
dev.run("hmc", init, C, W, 0, isteps=bench.ISTEPS, stepsize=bench.STEPSIZE, seed=5, store_draws=False, want=())
raw = dev.run("hmc", init, C, K, 0, isteps=bench.ISTEPS, stepsize=bench.STEPSIZE, seed=6, store_draws=False, want=())
flops = 4.0 * N * M * C
res = {"kernel": "k_glm_i8 (tcgen05.mma kind::i8)" if I8 else "k_glm (FP64 DMMA mma.sync.m8n8k4)",
       "value": raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3), "unit": bench.UNIT, "steps": K,
       "accept_rate": float(np.mean(raw["accept_rate"])), "glm_ms_per_launch": ms_glm, "ms_per_eval": ms_eval,
       "fp64_TFLOPs" if not I8 else "fp64_equivalent_TFLOPs": flops / (ms_glm * 1e-3) / 1e12,
       "relerr_logp_vs_long_double": err_l, "relerr_grad_vs_long_double": err_g, "n_kernel_launches": int(raw["n_kernel_launches"])}
if I8:
    # 28 slice-pair products of the same shape on the int8 tensor cores; peak = the issue rate of the instruction measured by
    # tools/i8_umma_probe.cu on this pool (profiles/i8_umma_probe_r01.jsonl)
    res["int8_roofline"] = {"bound": "tensor", "achieved": bench.I8_SLICE_PAIRS * flops / (ms_glm * 1e-3) / 1e12, "peak": bench.I8_PEAK_TOPS,
                            "unit": "Tera int8 op/s", "frac": bench.I8_SLICE_PAIRS * flops / (ms_glm * 1e-3) / 1e12 / bench.I8_PEAK_TOPS}
    res["beta_slices"] = 8 if os.environ.get("SMC_GLM_I8_BETA8") == "1" else 7
else:
    peak, _ = bench.fp64_peak()
    res["dmma_roofline"] = {"bound": "tensor", "achieved": flops / (ms_glm * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                            "frac": flops / (ms_glm * 1e-3) / 1e12 / peak}
print(json.dumps(res))
dev.close()
