"""TEST INFRASTRUCTURE (oracle): literal CPU restatement of the reference's two solvers,
/root/reference/src/solvers.jl:35-107 (simpleAGD) and :109-212 (simpleNM), statement by statement, on a
closure `ll_func` with the reference's signature (beta -> (logp, grad) / beta -> logp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this; the product
(simplemcmc.jl_b200/solvers.py) never does.  Parity status: pinned by source only (the reference ships no golden
vectors for the solvers and cannot run here).  Every intermediate the reference computes is recorded in `trace` so
that the device-side implementation can be compared iterate by iterate.
"""
import math

import numpy as np


class SolverRun(object):
    """src/solvers.jl:8-15"""

    def __init__(self, steps, maximum, converged):
        self.steps, self.maximum, self.converged = steps, maximum, converged
        self.params, self.misc = {}, {}


def simple_agd(ll_func, z0, maxiter=100, precision=1e-5):
    """src/solvers.jl:35-107.  Returns (steps, maximum, converged, z, trace)."""
    z0 = np.array(z0, dtype=np.float64)
    trace = []
    # first calc, :42-43
    f0, grad0 = ll_func(z0)
    if not math.isfinite(f0):
        raise AssertionError("Initial values out of model support, try other values")
    # initial Lipschitz constant, :46-54
    L0 = 1.0
    f1, grad1 = ll_func(z0 + grad0 / L0)
    while not math.isfinite(f1):
        L0 = L0 * 2.0
        f1, grad1 = ll_func(z0 + grad0 / L0)
    tmp = grad1 - grad0
    L00 = math.sqrt(np.dot(tmp, tmp) / np.dot(grad0, grad0)) * L0
    L0 = max(L0, L00)
    # loop, :57-98
    zb0 = z0.copy()
    theta0 = 1.0
    converged = False
    i = 0
    while i < maxiter and not converged:
        y0 = (1 - theta0) * z0 + theta0 * zb0                      # :63
        fy1, grady = ll_func(y0)                                   # :64
        zb1 = zb0 + grady / (theta0 * L0)                          # :65
        z1 = (1 - theta0) * z0 + theta0 * zb1                      # :66
        fz1, gradz = ll_func(zb1)                                  # :68
        h = 0
        while not math.isfinite(fz1) and h < 10:                   # :70-78
            L0 = L0 * 2.0
            zb1 = zb0 + grady / (theta0 * L0)
            z1 = (1 - theta0) * z0 + theta0 * zb1
            fz1, gradz = ll_func(zb1)
            h += 1
        if not h < 10:
            raise AssertionError("Convergence failed, exited function support")   # :79
        f1, gradz = ll_func(z1)                                    # :83
        d = y0 - z1
        L1 = L0 * 2.0 if 2 * abs(np.dot(d, grady - gradz)) / np.dot(d, d) > L0 else L0 * 0.9     # :84
        theta1 = 2 / (1 + math.sqrt(1 + 4 * L1 / (L0 * theta0 * theta0)))                        # :86
        converged = abs(f1 - f0) < precision                       # :88
        restart = np.dot(grady, z1 - z0) < 0.0                     # :90
        trace.append(dict(y0=y0.copy(), z1=z1.copy(), zb1=zb1.copy(), f1=f1, fy1=fy1, L0=L0, L1=L1, theta1=theta1,
                          restart=bool(restart), h=h))
        if restart:
            z0, zb0, theta0, L0, f0 = z1, z1, 1.0, L1, f1           # :92
        else:
            z0, zb0, theta0, L0, f0 = z1, zb1, theta1, L1, f1       # :95
        i += 1
    return i, f1, i < maxiter, z0, trace                             # :101 (converged field = i<maxiter)


def simple_nm(func, init, maxiter=100, precision=1e-3):
    """src/solvers.jl:122-212 (nelmin, simplified).  Returns (steps, ylo, converged, p[:, ilo], trace)."""
    if not precision > 0.0:
        raise AssertionError("precision should be > 0.")
    init = np.array(init, dtype=np.float64)
    n = len(init)
    if not n >= 1:
        raise AssertionError("there should be at least one parameter")
    f0 = -func(init)                                                # :131
    if not math.isfinite(f0):
        raise AssertionError("Initial values out of model support, try other values")
    step = np.ones(n)
    ccoeff, ecoeff, rcoeff = 0.5, 2.0, 1.0

    def stop_crit(p):                                               # :141
        return all(p[i, :].max() - p[i, :].min() < precision for i in range(n))

    p = np.array([[init[i] + step[i] * (i == j) for j in range(n + 1)] for i in range(n)], dtype=np.float64)  # :144
    y = np.array([-func(p[:, j]) for j in range(n + 1)])            # :145
    ilo = int(np.argmin(y))
    ylo = y[ilo]
    it = 0
    trace = []
    while it < maxiter and not stop_crit(p):
        ihi = int(np.argmax(y))                                     # :152
        keep = [j for j in range(n + 1) if j != ihi]
        pbar = np.array([sum(p[i, j] for j in keep) for i in range(n)]) / n       # :156 (sequential sum)
        pstar = pbar + rcoeff * (pbar - p[:, ihi])                  # :159
        ystar = -func(pstar)
        move = None
        if ystar < ylo:                                             # :162 extension
            p2star = pbar + ecoeff * (pstar - pbar)
            y2star = -func(p2star)
            if ystar < y2star:
                p[:, ihi], y[ihi] = pstar, ystar
            else:
                p[:, ihi], y[ihi] = p2star, y2star
            move = "extend"
        else:
            l = int(np.sum(ystar < y))                              # :171
            if l > 1:
                p[:, ihi], y[ihi] = pstar, ystar
                move = "reflect"
            elif l == 0:                                            # :177
                p2star = pbar + ccoeff * (p[:, ihi] - pbar)
                y2star = -func(p2star)
                if y[ihi] < y2star:                                 # :181 contract the whole simplex
                    p = np.array([[(p[i, j] + p[i, ilo]) * 0.5 for j in range(n + 1)] for i in range(n)])
                    y = np.array([-func(p[:, j]) for j in range(n + 1)])
                    move = "shrink"
                else:
                    p[:, ihi], y[ihi] = p2star, y2star
                    move = "contract_hi"
            elif l == 1:                                            # :190
                p2star = pbar + ccoeff * (pstar - pbar)
                y2star = -func(p2star)
                if ystar < y2star:
                    p[:, ihi], y[ihi] = pstar, ystar
                else:
                    p[:, ihi], y[ihi] = p2star, y2star
                move = "contract_refl"
        ilo = int(np.argmin(y))
        ylo = y[ilo]
        it += 1
        trace.append(dict(move=move, ylo=float(ylo), best=p[:, ilo].copy()))
    return it, float(ylo), it < maxiter, p[:, ilo].copy(), trace
