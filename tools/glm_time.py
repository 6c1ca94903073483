import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, bench
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
X, Y, b0, hold = bench.make_data()
C = int(os.environ.get("CHAINS", 4096))
low, dev = api._build(bench.MODEL, [("beta", np.zeros(64))], dict(X=X, Y=Y), True, C, 0, "assign")
b = 0.01 * np.random.default_rng(0).standard_normal((C, 64))
lp, g = dev.eval(b)
r = Y - X @ b[3]; g0 = -b[3] + X.T @ r
print("grad err", np.abs(g[3] - g0).max() / np.abs(g0).max())
ms_eval, ms_glm = dev.eval_timed(5, True)
print("variant %s C=%d: glm %.3f ms  %.2f TFLOP/s (%.1f%%)" % (os.environ.get("SMC_GLM_VARIANT", "0"), C, ms_glm, 4e6 * 64 * C / ms_glm / 1e9, 100 * 4e6 * 64 * C / ms_glm / 1e9 / 37.156))
