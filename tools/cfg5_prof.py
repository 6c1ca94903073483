"""cfg5 (OU, T = 1e6) evaluations for an ncu launch list: C = 1, 64, 4096 (eager launches, SMC_NO_GRAPHS=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200 import api
from util import ORNSTEIN, ou_data
x = ou_data(1_000_000, 5)
init = [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)]
for C in (1, 64, 4096):
    low, dev = api._build(ORNSTEIN, init, dict(x=x), True, C, 0, "assign")
    b = np.tile(np.array(low.init), (C, 1))
    for _ in range(3):
        lp, g = dev.eval(b)
    print(C, float(lp[0]), dev.eval_timed(5, True))
    dev.close()
