import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, bench
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api, capi
X, Y, b0, hold = bench.make_data()
mode = bench.posterior_mode(X, Y)
C = 512
low, dev = api._build(bench.MODEL, [("beta", np.zeros(64))], dict(X=X, Y=Y), True, C, 0, "assign")
rng = np.random.default_rng(0)
init = mode[None, :] + 1e-3 * rng.standard_normal((C, 64))
for eps in (1e-4, 2e-4, 3e-4, 4e-4, 5e-4, 7e-4, 1e-3):
    raw = dev.run("hmc", init, C, 40, 10, isteps=2, stepsize=eps, seed=1, store_draws=False, want=())
    print("stepsize %.1e accept %.3f  ms/step %.2f" % (eps, raw["accept_rate"].mean(), raw["device_ms"] / 40))
