import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from smc_pkg import load_package
smc = load_package()
from util import SLEEPSTUDY2, sleepstudy_data
data, nsub = sleepstudy_data()
t0 = time.time()
res = smc.simpleNUTS(SLEEPSTUDY2, steps=1300, burnin=1000, chains=4096, seed=7, data=data, store_draws=True,
                     intercept=250.0, slope=10.0, rand_int=np.zeros(nsub), rand_slope=np.zeros(nsub))
print("wall %.1f s device %.1f s" % (time.time() - t0, res.misc["device_ms"] / 1e3), "accept", res.acceptRate,
      "leapfrogs/s %.3g" % (res.misc["n_leapfrogs"] / (res.misc["device_ms"] * 1e-3)), "evals", res.misc["n_evals"])
print("eps last", float(res.misc["epsilon"][-1].mean()), "jmax mean last 100", float(res.misc["jmax"][-100:].mean()))
print("intercept mean %.3f sd %.3f; slope mean %.3f sd %.3f" % (res.params["intercept"].mean(), res.params["intercept"].std(),
      res.params["slope"].mean(), res.params["slope"].std()))
rh = res.misc["rhat"]; print("rhat max", float(np.nanmax(rh)))
# compare against a long HMC-free reference: ordinary least squares fixed effects of the synthetic data
