"""Per-iteration NUTS parity, device vs oracle, on the cases of tests/test_gpu_samplers.py::test_nuts_trajectory_parity."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
from oracle.refmodel import generate_model_function
from oracle.samplers import ReplayRNG, simple_nuts
from util import LINEAR, linear_data
X, Y, _ = linear_data(200, 5)
cases = [(LINEAR, [("beta", np.zeros(5))], dict(X=X, Y=Y)), ("x ~ Normal(1, 2)\ny ~ Gamma(2, 1)", [("x", 0.5), ("y", 1.0)], {})]
for model, init, dat in cases:
    ref = generate_model_function(model, data=dat, gradient=True, init=init)
    M = len(ref.init)
    steps, burnin, C, NU = 30, 5, 3, 270
    rng = np.random.default_rng(13)
    normals, unif = rng.standard_normal((steps, C, M)), rng.random((steps, C, NU))
    low, dev = api._build(model, init, dat, True, C, 0, "assign")
    raw = dev.run("nuts", ref.init, C, steps, burnin, inject_normals=normals, inject_uniforms=unif)
    for c in range(C):
        o = simple_nuts(ref, ref.init, ReplayRNG(normals[:, c], unif[:, c]), steps, burnin)
        de = np.abs(raw["epsilon"][:, c] - o["epsilon"]) / np.abs(o["epsilon"])
        dd = np.max(np.abs(raw["draws"][:, c] - o["draws"]), axis=1) / np.max(np.abs(o["draws"]))
        print("case", model[:12].replace("\n", " "), "chain", c, "eps err by iter:", " ".join("%.0e" % v for v in de))
        print("     draws err by sample:", " ".join("%.0e" % v for v in dd))
    dev.close()
