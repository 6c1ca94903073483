"""simpleAGD and simpleNM on the device closure.

simpleAGD follows /root/reference/src/solvers.jl:35-107 step for step (Lipschitz search on z0 + grad0/L0 with doubling
until finite, theta interpolation, L adaptation by 2x / 0.9x, |f1-f0| < precision, gradient restart, converged field =
i < maxiter); simpleNM follows :122-212 (nelmin: reflection / extension / two contractions / shrink, L-infinity
simplex criterion).  The control flow is the reference's; what is B200-specific is how the closure is called: points
whose coordinates do not depend on each other's values are evaluated as chains of ONE smc_eval launch --
 * AGD: zb1 and z1 both depend only on grad(y0) (solvers.jl:65-66), so ll(zb1) (:68) and ll(z1) (:83) are one
   two-chain evaluation: 2 launches per iteration instead of 3;
 * NM: the reflected point, the extension and both contractions are affine in (pbar, p[:,ihi]) (:159,:163,:178,:191),
   so all four candidates are one four-chain evaluation; the initial simplex and a shrink are one (n+1)-chain call.
The values consumed are exactly the ones the reference would have computed at those points.
"""
import inspect
import time

import numpy as np

from .lowering import ModelError


class SolverRun(object):
    """src/solvers.jl:8-15"""

    def __init__(self):
        self.time, self.steps, self.converged = np.nan, 0, False
        self.maximum, self.params, self.misc = np.nan, {}, {}

    def __repr__(self):
        # show(), src/solvers.jl:17-24
        head = "  ".join("%s= %s" % (k, v) for k, v in self.params.items())
        tail = "" if self.converged else "convergence precision not reached, "
        return "%s\n%s%d iterations, %.1f sec." % (head, tail, self.steps, self.time)


def _closure(model, kw, gradient, chains):
    from .api import generateModelFunction, _split_kwargs
    opts, init = _split_kwargs(dict(kw))
    caller = inspect.stack()[2].frame.f_globals
    return generateModelFunction(model, gradient=gradient, data=opts.get("data"), device=opts.get("device", 0),
                                 chains=chains, ref_adjoint=opts.get("ref_adjoint", "assign"), _globals=caller,
                                 **dict(init))


def _params(res, pars, z):
    for p in pars:
        seg = z[p.start:p.start + p.count]
        res.params[p.sym] = float(seg[0]) if len(p.size) == 0 else seg.reshape(p.size, order="F").copy()


def simpleAGD(model, maxiter=100, precision=1e-5, **kw):
    """src/solvers.jl:35-107"""
    t0 = time.time()
    ll, nparams, pars, z0 = _closure(model, kw, True, 2)
    res = SolverRun()
    nev = 0

    def ll1(z):
        return ll(z)

    def ll2(za, zb):
        lp, g = ll(np.stack([za, zb]))
        return float(lp[0]), g[0], float(lp[1]), g[1]

    f0, grad0 = ll1(z0)                                               # :42
    nev += 1
    if not np.isfinite(f0):
        raise ModelError("Initial values out of model support, try other values")
    L0 = 1.0                                                          # :46-54
    f1, grad1 = ll1(z0 + grad0 / L0)
    nev += 1
    while not np.isfinite(f1):
        L0 = L0 * 2.0
        f1, grad1 = ll1(z0 + grad0 / L0)
        nev += 1
    tmp = grad1 - grad0
    L00 = np.sqrt(np.dot(tmp, tmp) / np.dot(grad0, grad0)) * L0
    L0 = max(L0, float(L00))
    zb0 = z0.copy()
    theta0 = 1.0
    converged = False
    i = 0
    iterates = []
    while i < maxiter and not converged:
        y0 = (1 - theta0) * z0 + theta0 * zb0                         # :63
        fy1, grady = ll1(y0)                                          # :64
        zb1 = zb0 + grady / (theta0 * L0)                             # :65
        z1 = (1 - theta0) * z0 + theta0 * zb1                         # :66
        fz1, _, f1, gradz = ll2(zb1, z1)                              # :68 and :83 in one launch
        nev += 2
        h = 0
        while not np.isfinite(fz1) and h < 10:                        # :70-78
            L0 = L0 * 2.0
            zb1 = zb0 + grady / (theta0 * L0)
            z1 = (1 - theta0) * z0 + theta0 * zb1
            fz1, _, f1, gradz = ll2(zb1, z1)
            nev += 1
            h += 1
        if not h < 10:
            raise AssertionError("Convergence failed, exited function support")      # :79
        d = y0 - z1
        L1 = L0 * 2.0 if 2 * abs(np.dot(d, grady - gradz)) / np.dot(d, d) > L0 else L0 * 0.9      # :84
        theta1 = 2 / (1 + np.sqrt(1 + 4 * L1 / (L0 * theta0 * theta0)))                           # :86
        converged = abs(f1 - f0) < precision                          # :88
        if np.dot(grady, z1 - z0) < 0.0:                              # :90-96
            z0, zb0, theta0, L0, f0 = z1, z1, 1.0, L1, f1
        else:
            z0, zb0, theta0, L0, f0 = z1, zb1, float(theta1), L1, f1
        iterates.append(z1.copy())
        i += 1
    res.time, res.steps, res.maximum, res.converged = time.time() - t0, i, float(f1), i < maxiter     # :101
    res.misc["iterates"], res.misc["n_launches"], res.misc["L"] = iterates, nev, L0
    _params(res, pars, z0)
    ll.dev.close()
    return res


def simpleNM(model, maxiter=100, precision=1e-3, **kw):
    """src/solvers.jl:122-212"""
    t0 = time.time()
    if not precision > 0.0:
        raise AssertionError("precision should be > 0.")
    from .api import _split_kwargs
    _, init_pairs = _split_kwargs(dict(kw))
    n_guess = int(sum(np.size(v) for _, v in init_pairs))
    func, n, pars, init = _closure(model, kw, False, max(n_guess + 1, 4))
    if not n >= 1:
        raise AssertionError("there should be at least one parameter")
    res = SolverRun()
    nev = 1
    f0 = -func(init)                                                  # :131
    if not np.isfinite(f0):
        raise ModelError("Initial values out of model support, try other values")
    step = np.ones(n)
    ccoeff, ecoeff, rcoeff = 0.5, 2.0, 1.0

    def stop_crit():                                                  # :141
        return bool(np.all(p.max(axis=1) - p.min(axis=1) < precision))

    def neg_all(points):                                              # columns of `points` as chains of one launch
        return -np.asarray(func(np.ascontiguousarray(points.T)), dtype=np.float64)

    p = init[:, None] + np.hstack([np.diag(step), np.zeros((n, 1))])  # :144
    y = neg_all(p)                                                    # :145
    nev += 1
    ilo = int(np.argmin(y))
    ylo = y[ilo]
    it = 0
    moves = []
    while it < maxiter and not stop_crit():
        ihi = int(np.argmax(y))                                       # :152
        pbar = np.zeros(n)
        for j in range(n + 1):                                        # :156, summed in column order
            if j != ihi:
                pbar = pbar + p[:, j]
        pbar = pbar / n
        pstar = pbar + rcoeff * (pbar - p[:, ihi])                    # :159
        cand = np.stack([pstar, pbar + ecoeff * (pstar - pbar),       # :163
                         pbar + ccoeff * (p[:, ihi] - pbar),          # :178
                         pbar + ccoeff * (pstar - pbar)], axis=1)     # :191
        ystar, yext, ychi, ycrf = neg_all(cand)
        nev += 1
        if ystar < ylo:                                               # :162-169
            if ystar < yext:
                p[:, ihi], y[ihi] = pstar, ystar
            else:
                p[:, ihi], y[ihi] = cand[:, 1], yext
            moves.append("extend")
        else:
            l = int(np.sum(ystar < y))                                # :171
            if l > 1:
                p[:, ihi], y[ihi] = pstar, ystar
                moves.append("reflect")
            elif l == 0:
                if y[ihi] < ychi:                                     # :181-183
                    p = (p + p[:, [ilo]]) * 0.5
                    y = neg_all(p)
                    nev += 1
                    moves.append("shrink")
                else:
                    p[:, ihi], y[ihi] = cand[:, 2], ychi
                    moves.append("contract_hi")
            else:                                                     # l == 1, :190-197
                if ystar < ycrf:
                    p[:, ihi], y[ihi] = pstar, ystar
                else:
                    p[:, ihi], y[ihi] = cand[:, 3], ycrf
                moves.append("contract_refl")
        ilo = int(np.argmin(y))
        ylo = y[ilo]
        it += 1
    res.time, res.steps, res.maximum, res.converged = time.time() - t0, it, float(ylo), it < maxiter     # :205
    res.misc["moves"], res.misc["n_launches"] = moves, nev
    _params(res, pars, p[:, ilo])
    func.dev.close()
    return res
