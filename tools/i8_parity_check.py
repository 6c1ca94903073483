"""Parity of the int8 tensor-core GLM path (csrc/glm_i8.cu, the default for 33-64 columns and >= 32 chains) against the
oracle and against the FP64 DMMA path (SMC_GLM_I8=0), through the public API.  Each path runs in its own process (the
switch is read once):
    python tools/i8_parity_check.py [rows] [chains]
Prints one JSON line; exit code 1 when the 1e-12 bar is missed."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np

MODEL = """
    beta ~ Normal(0, 1.0)
    resid = Y - X * beta
    resid ~ Normal(0, 1.5)
"""


def evaluate(N, M, C):
    from smc_pkg import load_package
    smc = load_package()
    rng = np.random.default_rng(21)
    X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
    b0 = rng.standard_normal(M)
    Y = X @ b0 + rng.standard_normal(N)
    data = dict(X=X, Y=Y)
    f, nparams, pars, init = smc.generateModelFunction(MODEL, gradient=True, chains=C, data=data, beta=np.zeros(M))
    betas = b0[None, :] + 1e-3 * rng.standard_normal((C, M))
    betas[C // 2, 3] = np.nan                    # a chain out of support must come back as -Inf / zero gradient
    lp, g = f(betas)
    return data, betas, lp, g


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000         # not a multiple of 128: ragged last tile
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 70           # not a multiple of 32: ragged chain tile
    M = 50                                                      # narrower than the 64 parameter slots
    if "--dump" in sys.argv:                                    # child: evaluate with whatever SMC_GLM_I8 says
        data, betas, lp, g = evaluate(N, M, C)
        np.savez(sys.argv[sys.argv.index("--dump") + 1], lp=lp, g=g)
        sys.exit(0)
    assert os.environ.get("SMC_GLM_I8", "1") != "0", "this process must run the int8 path (unset SMC_GLM_I8)"
    data, betas, lp, g = evaluate(N, M, C)
    from oracle.refmodel import generate_model_function
    ref = generate_model_function(MODEL, data=data, gradient=True, beta=np.zeros(M))
    bad = C // 2
    err_l = err_g = 0.0
    for c in range(C):
        if c == bad:
            continue
        l0, g0 = ref(betas[c])
        err_l = max(err_l, abs(lp[c] - l0) / abs(l0))
        err_g = max(err_g, float(np.max(np.abs(g[c] - g0)) / np.max(np.abs(g0))))
    # the FP64 DMMA path in a child process
    tmp = "/tmp/i8_parity_default.npz"
    env = dict(os.environ, SMC_GLM_I8="0")
    subprocess.check_call([sys.executable, os.path.abspath(__file__), str(N), str(C), "--dump", tmp], env=env)
    dflt = np.load(tmp)
    ok_rows = np.arange(C) != bad
    res = dict(rows=N, M=M, chains=C, relerr_logp_vs_oracle=err_l, relerr_grad_vs_oracle=err_g,
               relerr_logp_vs_dmma=float(np.max(np.abs(lp[ok_rows] - dflt["lp"][ok_rows]) / np.abs(dflt["lp"][ok_rows]))),
               relerr_grad_vs_dmma=float(np.max(np.abs(g[ok_rows] - dflt["g"][ok_rows])) / np.max(np.abs(dflt["g"][ok_rows]))),
               bad_chain_logp=float(lp[bad]), bad_chain_grad_zero=bool(np.all(g[bad] == 0.0)))
    res["meets_1e-12"] = bool(err_l <= 1e-12 and err_g <= 1e-12 and lp[bad] == -np.inf)
    print(json.dumps(res))
    sys.exit(0 if res["meets_1e-12"] else 1)
