"""BASELINE.json configs[3] at its stated size: simpleNUTS, 16 384 chains, steps = 2000 (nadapt is a literal 1000 in the reference,
src/samplers.jl:214, so the second half runs at the adapted step size).  Reports leapfrogs/s, active-lane fraction, the adapted
step size, tree depths and the pooled R-hat from the device-side moments.
    [SMC_NUTS_ASYNC=1] python tools/nuts_baseline_run.py sleepstudy|hierarchical [chains] [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api, parallel
from util import HIERARCHICAL, SLEEPSTUDY2, hierarchical_data, sleepstudy_data

which = sys.argv[1] if len(sys.argv) > 1 else "sleepstudy"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
burnin = steps // 2
if which == "sleepstudy":
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    low, dev = api._build(SLEEPSTUDY2, init, data, True, C, 0, "assign")
else:
    data, (N, D, L) = hierarchical_data()
    init = [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))]
    low, dev = api._build(HIERARCHICAL, init, data, True, C, 0, "add")
dev.run("nuts", np.array(low.init), C, 4, 0, seed=3, store_draws=False, want=())          # warm-up: buffers, evaluation graphs
t0 = time.time()
raw = dev.run("nuts", np.array(low.init), C, steps, burnin, seed=1, store_draws=False, want=("epsilon", "jmax"))
wall = time.time() - t0
M = raw["mean"].shape[1]
pooled = parallel.rhat_from_partials(parallel.moment_partials(raw["mean"], raw["m2"], steps - burnin), steps - burnin, M)
print(json.dumps(dict(model=which, chains=C, steps=steps, burnin=burnin, mode={"1": "asynchronous from iteration 0", "0": "lockstep"}.get(os.environ.get("SMC_NUTS_ASYNC"), "default: lockstep, asynchronous once a checkpoint finds < 0.36 of the lanes active"),
                      leapfrogs_per_s=raw["n_leapfrogs"] / (raw["device_ms"] * 1e-3), device_s=raw["device_ms"] * 1e-3, wall_s=wall,
                      rounds=int(raw["n_evals"]) - 1, active_lane_fraction=raw["n_leapfrogs"] / (max(raw["n_evals"] - 1, 1) * C),
                      iterations_per_s=C * steps / (raw["device_ms"] * 1e-3), kernel_launches=int(raw["n_kernel_launches"]),
                      epsilon_adapted_mean=float(raw["epsilon"][-1].mean()), jmax_mean_second_half=float(raw["jmax"][burnin:].mean()),
                      accept_rate=float(raw["accept_rate"].mean()), rhat_max=float(np.nanmax(pooled["rhat"])),
                      rhat_median=float(np.nanmedian(pooled["rhat"])))))
dev.close()
