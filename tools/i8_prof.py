"""A run of the int8 GLM path for ncu (default N = 200 000, M = 64, 1024 chains: about 1 ms per launch; the configs[1] size
for the DRAM traffic of one launch: 1000000 4096).
    python tools/i8_prof.py [rows] [chains]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api, capi
N, M, C = (int(sys.argv[1]) if len(sys.argv) > 1 else 200_000), 64, (int(sys.argv[2]) if len(sys.argv) > 2 else 1024)
rng = np.random.default_rng(0)
X = np.empty((M, N)).T
X[:, 0] = 1.0
for j in range(1, M):
    X[:, j] = rng.standard_normal(N)
Y = X @ (0.1 * rng.standard_normal(M)) + rng.standard_normal(N)
model = "beta ~ Normal(0, 1.0)\nresid = Y - X * beta\nresid ~ Normal(0, 1.0)"
low, dev = api._build(model, [("beta", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
b = 0.05 * rng.standard_normal((C, M))
for _ in range(4):
    lp, g = dev.eval(b)
print("ok", float(lp[0]), dev.eval_timed(reps=3))
dev.close()
