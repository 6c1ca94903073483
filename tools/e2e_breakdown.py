"""Where the end-to-end time of one simpleHMC call on configs[1] goes (host wall clock per phase)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import bench
from smc_pkg import load_package
smc = load_package()
from simplemcmc_jl_b200 import api, capi
from simplemcmc_jl_b200.lowering import Lowering

X, Y, beta0, hold = bench.make_data(pinned=True)
mode = bench.posterior_mode(X, Y)
C, K = bench.CHAINS_PER_GPU, int(sys.argv[1]) if len(sys.argv) > 1 else 10
ctx = capi.Context.default(0)
for rep in range(6):
    t = [time.time()]
    low = Lowering(bench.MODEL, [("beta", mode)], dict(X=X, Y=Y)); t.append(time.time())
    tape, bound = low.serialise(); t.append(time.time())
    dev = capi.DeviceModel.__new__(capi.DeviceModel)
    dev.ctx, dev.lib = ctx, ctx.lib
    import ctypes
    dev.h = ctypes.c_void_p()
    capi.check(dev.lib.smc_model_compile(ctx.h, tape, len(tape), ctypes.byref(dev.h))); t.append(time.time())
    dev.h2d_bytes = 0
    for name, arr in bound:
        dev.bind(name, arr)
    t.append(time.time())
    dev.max_chains, dev.row_shard = 0, False
    dev.finalize(C); t.append(time.time())
    dev.nparams = 64
    raw = dev.run("hmc", np.array(low.init), C, K, 0, isteps=2, stepsize=bench.STEPSIZE, seed=1); t.append(time.time())
    dev.close(); t.append(time.time())
    names = ["lowering", "serialise", "compile", "bind", "finalize", "run(K=%d)" % K, "close"]
    print("rep", rep, " ".join("%s=%.1fms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])), "device_ms=%.1f" % raw["device_ms"])
t0 = time.time()
res = smc.simpleHMC(bench.MODEL, steps=K, burnin=0, isteps=2, stepsize=bench.STEPSIZE, chains=C, seed=4, data=dict(X=X, Y=Y), comm=ctx, beta=mode)
print("simpleHMC total %.1f ms (device %.1f ms)" % (1e3 * (time.time() - t0), res.misc["device_ms"]))
