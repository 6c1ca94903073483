"""Smallest possible device check of the opt-in int8 GLM path: SMC_GLM_I8=1 python tools/i8_quick.py"""
import json, os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from smc_pkg import load_package
smc = load_package()
N, M, C, S = 3000, 50, 70, 1.5
rng = np.random.default_rng(21)
X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
b0 = rng.standard_normal(M)
Y = X @ b0 + rng.standard_normal(N)
model = "beta ~ Normal(0, 1.0)\nresid = Y - X * beta\nresid ~ Normal(0, 1.5)"
f, _, _, _ = smc.generateModelFunction(model, gradient=True, chains=C, data=dict(X=X, Y=Y), beta=np.zeros(M))
B = b0[None, :] + 1e-3 * rng.standard_normal((C, M))
lp, g = f(B)
Xl, Yl, Bl = X.astype(np.longdouble), Y.astype(np.longdouble), B.astype(np.longdouble)
d = Yl[None, :] - Bl @ Xl.T
c = np.log(np.sqrt(2 * np.pi))
Lt = np.sum(-0.5 * Bl * Bl - c, axis=1) + np.sum(-0.5 * (d / S) ** 2 - np.log(S) - c, axis=1)
Gt = -Bl + (d @ Xl) / S ** 2
print(json.dumps(dict(i8=os.environ.get("SMC_GLM_I8"), relerr_logp=float(np.max(np.abs((lp - Lt) / Lt))),
                      relerr_grad=float(np.max(np.abs(g - Gt)) / np.max(np.abs(Gt))), seconds=round(time.time() - t0, 2))))
