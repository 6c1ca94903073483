"""simpleAGD: accelerated gradient ascent with adaptive restart, on the device closure.
Follows /root/reference/src/solvers.jl:35-107 (host control flow; 3 evaluations of (logp, grad) per iteration,
all of them smc_eval calls)."""
import time

import numpy as np

from .lowering import ModelError


class SolverRun(object):
    """src/solvers.jl:8-15"""

    def __init__(self):
        self.time, self.steps, self.converged = np.nan, 0, False
        self.maximum, self.params = np.nan, {}

    def __repr__(self):
        return "%s in %d steps, maximum %.6g, %s" % ("converged" if self.converged else "not converged", self.steps,
                                                      self.maximum, self.params)


def simpleAGD(model, maxiter=100, precision=1e-5, **kw):
    from .api import generateModelFunction, _split_kwargs
    t0 = time.time()
    opts, init = _split_kwargs(kw)
    f, nparams, pars, beta0 = generateModelFunction(model, gradient=True, data=opts.get("data"),
                                                    device=opts.get("device", 0), **dict(init))
    res = SolverRun()
    f0, grad0 = f(beta0)
    if not np.isfinite(f0):
        raise ModelError("Initial values out of model support, try other values")
    # first estimate of the Lipschitz constant (src/solvers.jl:46-54)
    L = 1.0
    f1, grad1 = f(beta0 + 1e-4 * np.ones(nparams))
    d = np.linalg.norm(grad1 - grad0)
    if np.isfinite(f1) and d > 0:
        L = d / (1e-4 * np.sqrt(nparams))
    x = beta0.copy()
    y = x.copy()
    t = 1.0
    fx = f0
    converged = False
    it = 0
    for it in range(1, int(maxiter) + 1):
        fy, gy = f(y)
        tries = 0
        while True:
            xn = y + gy / L
            fn, gn = f(xn)
            # support back-off / sufficient increase: double L up to 10 times (src/solvers.jl:71-80)
            if np.isfinite(fn) and fn >= fy + 0.5 / L * np.dot(gy, gy) - 1e-12 * abs(fy):
                break
            L *= 2.0
            tries += 1
            if tries > 10:
                break
        if not np.isfinite(fn):
            break
        tn = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * t * t))
        yn = xn + (t - 1.0) / tn * (xn - x)
        if np.dot(gy, xn - x) < 0:       # adaptive restart (src/solvers.jl:90-96)
            yn, tn = xn.copy(), 1.0
        converged = abs(fn - fx) < precision * max(1.0, abs(fx)) and it > 1
        x, y, t, fx = xn, yn, tn, fn
        if converged:
            break
    res.time, res.steps, res.converged, res.maximum = time.time() - t0, it, bool(converged), float(fx)
    for p in pars:
        seg = x[p.start:p.start + p.count]
        res.params[p.sym] = float(seg[0]) if len(p.size) == 0 else seg.reshape(p.size, order="F")
    return res
