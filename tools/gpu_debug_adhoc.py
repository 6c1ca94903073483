"""Ad-hoc GPU check used during development (not a test): rule matrix + example models vs the oracle."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # (a development aid that calls the oracle as the checker, like the tests)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
smc = load_package()
from oracle.refmodel import generate_model_function
from rule_cases import all_cases

def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if a.size else 0.0

t0 = time.time()
bad = 0; worst = 0
cases = all_cases()
for k, (ex, data, x0) in enumerate(cases):
    model = "sum(%s)" % ex
    try:
        ref = generate_model_function(model, data=data, gradient=True, x=x0)
        f, n, pars, init = smc.generateModelFunction(model, gradient=True, data=data, x=x0)
        l1, g1 = f(init)
        l0, g0 = ref(init)
        e = max(abs(l1 - l0) / max(abs(l0), 1e-300) if np.isfinite(l0) else 0.0, relerr(g1, g0) if np.max(np.abs(g0)) > 0 else float(np.max(np.abs(g1))))
        worst = max(worst, e)
        if not e < 1e-12:
            bad += 1; print("MISMATCH", ex, {k: np.shape(v) for k, v in data.items()}, np.shape(x0), l1, l0, g1, g0)
    except Exception as ex_:
        bad += 1; print("ERROR", ex, {k: np.shape(v) for k, v in data.items()}, np.shape(x0), type(ex_).__name__, str(ex_)[:200])
print("rule matrix: %d cases, %d bad, worst rel err %.2e, %.1fs" % (len(cases), bad, worst, time.time() - t0))

rng = np.random.default_rng(1)
N, M = 1000, 10
X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))]); b0 = rng.standard_normal(M)
Y = X @ b0 + rng.standard_normal(N)
lin = "beta ~ Normal(0, 1.0)\nresid = Y - X * beta\nresid ~ Normal(0, 1.0)"
data = dict(X=X, Y=Y)
ref = generate_model_function(lin, data=data, gradient=True, beta=np.zeros(M))
from simplemcmc_jl_b200 import api, capi
for fuse in (True, False):
    for C in (1, 5, 64, 200):
        low, dev = api._build(lin, [("beta", np.zeros(M))], data, True, C, 0, "assign", fuse=fuse)
        betas = 0.3 * rng.standard_normal((C, M))
        lp, g = dev.eval(betas)
        el = eg = 0
        for c in range(C):
            l0, g0 = ref(betas[c]); el = max(el, abs(lp[c] - l0) / abs(l0)); eg = max(eg, relerr(g[c], g0))
        print("linear fuse=%s C=%d: logp err %.2e grad err %.2e (fused stmts %d)" % (fuse, C, el, eg, low.n_fused))
Yb = (rng.random(N) < 1 / (1 + np.exp(X @ b0))).astype(float)
logit = "vars ~ Normal(0, 1.0)\nprob = 1 / (1. + exp(X * vars))\nY ~ Bernoulli(prob)"
data2 = dict(X=X, Y=Yb)
ref2 = generate_model_function(logit, data=data2, gradient=True, vars=np.zeros(M))
for fuse in (True, False):
    low, dev = api._build(logit, [("vars", np.zeros(M))], data2, True, 37, 0, "assign", fuse=fuse)
    betas = 0.3 * rng.standard_normal((37, M)); lp, g = dev.eval(betas)
    el = eg = 0
    for c in range(37):
        l0, g0 = ref2(betas[c]); el = max(el, abs(lp[c] - l0) / abs(l0)); eg = max(eg, relerr(g[c], g0))
    print("logistic fuse=%s: logp err %.2e grad err %.2e (fused %d)" % (fuse, el, eg, low.n_fused))
# HMC trajectory with injected randomness
from oracle.samplers import simple_hmc, simple_rwm, ReplayRNG
steps, burnin, C = 30, 5, 3
normals = rng.standard_normal((steps, C, M)); unif = rng.random((steps, C, 1))
low, dev = api._build(lin, [("beta", np.zeros(M))], data, True, C, 0, "assign")
raw = dev.run("hmc", np.zeros(M), C, steps, burnin, isteps=2, stepsize=0.05, inject_normals=normals, inject_uniforms=unif)
for c in range(C):
    o = simple_hmc(ref, np.zeros(M), ReplayRNG(normals[:, c], unif[:, c]), steps, burnin, 2, 0.05)
    print("hmc chain %d: draws err %.2e loglik err %.2e accept equal %s" % (c, relerr(raw["draws"][:, c], o["draws"]), relerr(raw["loglik"][:, c], o["loglik"]), np.array_equal(raw["accept"][:, c], o["accept"])))
ref0 = generate_model_function(lin, data=data, gradient=False, beta=np.zeros(M))
low, dev = api._build(lin, [("beta", np.zeros(M))], data, False, C, 0, "assign")
raw = dev.run("rwm", np.zeros(M), C, steps, burnin, inject_normals=normals, inject_uniforms=unif)
for c in range(C):
    o = simple_rwm(ref0, np.zeros(M), ReplayRNG(normals[:, c], unif[:, c]), steps, burnin)
    print("rwm chain %d: draws err %.2e loglik err %.2e accept equal %s" % (c, relerr(raw["draws"][:, c], o["draws"]), relerr(raw["loglik"][:, c], o["loglik"]), np.array_equal(raw["accept"][:, c], o["accept"])))
# free-running HMC
res = smc.simpleHMC(lin, steps=2000, burnin=500, isteps=2, stepsize=0.05, chains=256, data=data, beta=np.zeros(M))
post_mean = np.linalg.solve(X.T @ X + np.eye(M), X.T @ Y)
print(res); print("posterior mean err", np.abs(res.params["beta"].mean(axis=(1, 2)) - post_mean).max(), "rhat max", res.misc["rhat"].max())
# big config timing
N2, M2, C2 = 1_000_000, 64, 4096
X2 = np.column_stack([np.ones(N2), rng.standard_normal((N2, M2 - 1))]); Y2 = X2 @ rng.standard_normal(M2) + rng.standard_normal(N2)
t = time.time()
low, dev = api._build(lin, [("beta", np.zeros(M2))], dict(X=X2, Y=Y2), True, C2, 0, "assign")
print("bind+finalize %.2fs" % (time.time() - t))
betas = 0.01 * rng.standard_normal((C2, M2))
lp, g = dev.eval(betas)
r = Y2 - X2 @ betas[3]; l0 = (-0.5 * betas[3] @ betas[3] - 0.5 * r @ r - (N2 + M2) * 0.9189385332046727); g0 = -betas[3] + X2.T @ r
print("cfg2 closed-form check: logp err %.2e grad err %.2e" % (abs(lp[3] - l0) / abs(l0), relerr(g[3], g0)))
ms_eval, ms_glm = dev.eval_timed(5, True)
flops = 4.0 * N2 * M2 * C2
print("cfg2 eval %.2f ms, glm kernel %.2f ms -> %.2f TFLOP/s (%.1f%% of 37.15 measured DMMA peak); %.3g leapfrog chain-steps/s"
      % (ms_eval, ms_glm, flops / ms_glm / 1e9, 100 * flops / ms_glm / 1e9 / 37.15, C2 / (ms_eval * 1e-3)))
