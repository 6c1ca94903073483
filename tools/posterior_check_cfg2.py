"""north_star's second correctness leg at the configs[1] size: posterior means and variances of the sampled chains against the
closed form of the conjugate linear model (beta | Y ~ Normal(mode, (X'X + I)^-1)), in Monte-Carlo standard errors.  4096 chains
started from the exact posterior (mode + chol((X'X + I)^-1) z) so that no burn-in bias is in the comparison, simpleHMC with the
bench's settings; moments from the device-side Welford accumulators.   python tools/posterior_check_cfg2.py [steps]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import bench
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 150
C, M, N = bench.CHAINS_PER_GPU, bench.N_COLS, bench.N_ROWS
X, Y, _, _ = bench.make_data()
A = X.T @ X + np.eye(M)
cov = np.linalg.inv(A)
mode = cov @ (X.T @ Y)
rng = np.random.default_rng(42)
init = mode[None, :] + rng.standard_normal((C, M)) @ np.linalg.cholesky(cov).T
low, dev = api._build(bench.MODEL, [("beta", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
t0 = time.time()
raw = dev.run("hmc", init, C, steps, 0, isteps=bench.ISTEPS, stepsize=bench.STEPSIZE, seed=17, store_draws=False, want=())
dt = time.time() - t0
mean_c, m2_c = raw["mean"], raw["m2"]                      # per chain: mean over its samples, sum of squared deviations
# pooled over chains: the chain means are (nearly) independent draws of a quantity with variance <= posterior variance
pooled_mean = mean_c.mean(0)
se_mean = mean_c.std(0, ddof=1) / np.sqrt(C)
z_mean = (pooled_mean - mode) / se_mean
# total variance of all draws = within + between
within = m2_c.sum(0) / (C * steps)
between = ((mean_c - pooled_mean) ** 2).mean(0)
var_hat = within + between
# standard error of the variance estimate from the spread of per-chain second moments about the pooled mean
per_chain_var = m2_c / steps + (mean_c - pooled_mean) ** 2
se_var = per_chain_var.std(0, ddof=1) / np.sqrt(C)
z_var = (var_hat - np.diag(cov)) / se_var
print(json.dumps(dict(model="configs[1]: linear regression N=1e6 x M=64, simpleHMC isteps=2 stepsize=5e-4", chains=C, steps=steps,
                      glm_kernel="k_glm_i8" if os.environ.get("SMC_GLM_I8", "1") != "0" else "k_glm", seconds=round(dt, 2),
                      accept_rate=float(raw["accept_rate"].mean()),
                      max_abs_z_of_posterior_means=float(np.max(np.abs(z_mean))), rms_z_of_posterior_means=float(np.sqrt(np.mean(z_mean ** 2))),
                      max_abs_z_of_posterior_variances=float(np.max(np.abs(z_var))), rms_z_of_posterior_variances=float(np.sqrt(np.mean(z_var ** 2))),
                      max_rel_dev_of_variances=float(np.max(np.abs(var_hat / np.diag(cov) - 1.0))),
                      note="z = (estimate - closed form) / Monte-Carlo standard error over the 4096 chains, 64 components: |z| of a few is "
                           "agreement; chains start from exact posterior draws")))
dev.close()
