"""Stress of the int8 GLM kernel's pipeline protocol: many tiles per CTA, a short read-back period (SMC_GLM_I8_FLUSH), scale
growth in the middle of a CTA's rows (a block of large residuals), repeated evaluations compared bit for bit.
    SMC_GLM_I8_FLUSH=3 SMC_GLM_I8_CHECK=1 SMC_NO_GRAPHS=1 python tools/i8_stress.py [rows] [chains] [reps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
N = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
M = 64
rng = np.random.default_rng(4)
X = np.empty((M, N)).T
X[:, 0] = 1.0
for j in range(1, M):
    X[:, j] = rng.standard_normal(N)
b0 = 0.1 * rng.standard_normal(M)
Y = X @ b0 + rng.standard_normal(N)
# residual scale grows along the rows: every CTA's sticky scales have to grow several times
Y[N // 3:] += 40.0 * rng.standard_normal(N - N // 3)
Y[2 * N // 3:] += 3000.0 * rng.standard_normal(N - 2 * N // 3)
model = "beta ~ Normal(0, 1.0)\nresid = Y - X * beta\nresid ~ Normal(0, 1.0)"
low, dev = api._build(model, [("beta", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
B = b0[None, :] + 0.05 * rng.standard_normal((C, M))
t0 = time.time()
lp, g = dev.eval(B)
same = True
for r in range(reps):
    lp2, g2 = dev.eval(B)
    same = same and np.array_equal(lp, lp2) and np.array_equal(g, g2)
Xl, Yl = X.astype(np.longdouble), Y.astype(np.longdouble)
errs = []
for c in (0, C // 2, C - 1):
    bl = B[c].astype(np.longdouble)
    d = Yl - Xl @ bl
    L = np.sum(-0.5 * d * d) - N * np.longdouble(0.91893853320467274178) + np.sum(-0.5 * bl * bl) - M * np.longdouble(0.91893853320467274178)
    G = -bl + Xl.T @ d
    errs.append((float(abs((lp[c] - L) / L)), float(np.max(np.abs(g[c] - G)) / np.max(np.abs(G)))))
print(json.dumps(dict(rows=N, chains=C, flush=os.environ.get("SMC_GLM_I8_FLUSH"), reps=reps, bit_identical_repeats=bool(same),
                      relerr_logp=max(e[0] for e in errs), relerr_grad=max(e[1] for e in errs), seconds=round(time.time() - t0, 2))))
dev.close()
