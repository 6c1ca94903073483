"""Short runs of the secondary configs for `ncu --metrics gpu__time_duration.sum` launch lists / `ncu --set full` captures:
  python tools/prof_cfg.py cfg3 [chains] | cfg4 [chains] | cfg5 [chains] [sampler]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
from util import HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, ou_data, sleepstudy_data

which = sys.argv[1]
C = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if which == "cfg1":
    C = C or (1 << 20)
    X, Y, _ = linear_data(1000, 10, 1)
    mode = np.linalg.solve(X.T @ X + np.eye(10), X.T @ Y)
    low, dev = api._build(LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y), True, C, 0, "assign")
    raw = dev.run("hmc", mode, C, 4, 0, isteps=2, stepsize=0.05, seed=1, store_draws=False, want=())
elif which == "cfg3":
    C = C or 1
    rng = np.random.default_rng(3)
    N, M = 10_000_000, 10
    X = np.empty((M, N)).T
    X[:, 0] = 1.0
    for j in range(1, M):
        X[:, j] = rng.standard_normal(N)
    b0 = rng.standard_normal(M)
    Y = (rng.random(N) < 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
    low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
    raw = dev.run("hmc", b0, C, 4, 0, isteps=2, stepsize=0.1 / np.sqrt(N / 1000), seed=1, store_draws=False, want=())
elif which == "cfg4h":
    C = C or 16384
    data, (N, D, L) = hierarchical_data()
    init = [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))]
    low, dev = api._build(HIERARCHICAL, init, data, True, C, 0, "add")
    raw = dev.run("nuts", np.array(low.init), C, int(os.environ.get("NUTS_STEPS", 3)), 0, seed=1, store_draws=False, want=())
elif which == "cfg4":
    C = C or 16384
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    low, dev = api._build(SLEEPSTUDY2, init, data, True, C, 0, "assign")
    raw = dev.run("nuts", np.array(low.init), C, int(os.environ.get("NUTS_STEPS", 3)), 0, seed=1, store_draws=False, want=())
else:
    C = C or 64
    kind = sys.argv[3] if len(sys.argv) > 3 else "hmc"
    x = ou_data(1_000_000, 5)
    init = [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)]
    low, dev = api._build(ORNSTEIN, init, dict(x=x), kind != "rwm", C, 0, "assign")
    raw = dev.run(kind, np.array(low.init), C, 6, 0, seed=1, store_draws=False, want=(), **(dict(isteps=2, stepsize=1e-4) if kind == "hmc" else {}))
print(which, C, "device_ms", raw["device_ms"], "evals", raw["n_evals"], "launches", raw["n_kernel_launches"])
