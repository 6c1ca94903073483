"""The reference's public surface, unchanged in names, keyword convention and return structures
(/root/reference/src/SimpleMCMC.jl:9-11):

    generateModelFunction(model; gradient=false, debug=false, init...)      src/parsing.jl:394
    simpleRWM(model; steps=1000, burnin=100, init...)                        src/samplers.jl:70
    simpleHMC(model; steps=1000, burnin=100, isteps=2, stepsize=1e-3, init...)  src/samplers.jl:124
    simpleNUTS(model; steps=1000, burnin=100, init...)                       src/samplers.jl:180
    simpleAGD(model; maxiter=100, precision=1e-5, init...)                   src/solvers.jl:35

Every keyword that is not reserved is a model parameter with its initial value (README.md:50), packed into the
parameter vector in keyword order, matrices column-major (src/parsing.jl:303-334).  External variables are looked
up by name: first in the reserved keyword `data` (a dict), then in the caller's module globals -- the Python
stand-in for "globals of Main" (src/parsing.jl:484-485).  New reserved keywords (B200 build): `chains`, `seed`,
`device`, `data`, `shard`, `ref_adjoint`, `store_draws`, `comm`, `chain_offset`, `sufficient_stats` (opt-in: a Normal
identity-link regression evaluates from X'X, X'y, y'y computed once -- 2*M^2 instead of 4*N*M flops per chain-step; a
different algorithm from the reference's generated code, so it is never on by default and is reported separately).
The loops themselves run inside libsimplemcmc_cuda.so.
"""
import inspect
import time

import numpy as np

from . import capi
from .lowering import Lowering, ModelError

RESERVED = ("steps", "burnin", "isteps", "stepsize", "maxiter", "precision", "gradient", "debug", "chains", "seed",
            "device", "data", "shard", "ref_adjoint", "store_draws", "comm", "chain_offset", "sufficient_stats")


class _DataLookup(dict):
    """name -> value: explicit `data` first, then the caller's globals"""

    def __init__(self, explicit, fallback):
        dict.__init__(self, explicit or {})
        self.fallback = fallback or {}

    def __contains__(self, k):
        return dict.__contains__(self, k) or self._ok(k)

    def _ok(self, k):
        v = self.fallback.get(k, None)
        return isinstance(v, (int, float, np.ndarray, list, np.integer, np.floating))

    def __missing__(self, k):
        if self._ok(k):
            return self.fallback[k]
        raise KeyError(k)


def _caller_globals(depth=2):
    try:
        return inspect.stack()[depth].frame.f_globals
    except Exception:
        return {}


class ModelFunction(object):
    """the closure returned by generateModelFunction: f(beta) -> logp  or  (logp, grad)"""

    def __init__(self, lowered, dev, gradient):
        self.lowered, self.dev, self.gradient = lowered, dev, gradient

    def __call__(self, beta):
        beta = np.asarray(beta, dtype=np.float64)
        single = beta.ndim == 1
        logp, grad = self.dev.eval(beta, gradient=self.gradient)
        if single:
            return (float(logp[0]), grad[0]) if self.gradient else float(logp[0])
        return (logp, grad) if self.gradient else logp


def _build(model, init, data, gradient, chains, device, ref_adjoint, shard=None, comm=None, fuse=True, fuse_statements=True,
           sufficient_stats=False):
    low = Lowering(model, init, data, gradient=gradient, ref_adjoint=ref_adjoint, fuse=fuse, fuse_statements=fuse_statements,
                   sufficient_stats=bool(sufficient_stats) and shard != "rows")
    tape, bound = low.serialise()
    ctx = comm if isinstance(comm, capi.Context) else capi.Context.default(device)
    row_shard = (shard == "rows")
    dev = capi.DeviceModel(ctx, tape, bound, max_chains=chains, row_shard=row_shard)
    return low, dev


def generateModelFunction(model, gradient=False, debug=False, chains=1, device=0, data=None, ref_adjoint="assign",
                          sufficient_stats=False, _globals=None, **init):
    """-> (f, nparams, pars, init) like src/parsing.jl:492; debug=True returns the statement lists instead of
    compiling (the reference returns the generated Expr).  `_globals` (internal): the module globals external variables
    are looked up in when the call comes through a solver rather than from user code."""
    lookup = _DataLookup(data, _globals if _globals is not None else _caller_globals())
    if debug:
        low = Lowering(model, list(init.items()), lookup, gradient=gradient, ref_adjoint=ref_adjoint)
        return low.statements()
    low, dev = _build(model, list(init.items()), lookup, gradient, chains, device, ref_adjoint, sufficient_stats=sufficient_stats)
    f = ModelFunction(low, dev, gradient)
    # the reference's preCalculate rejects initial values outside the support (src/parsing.jl:382-383)
    if gradient:
        lp = f(np.array(low.init))[0]
        if lp == -np.inf:
            raise ModelError("Initial values out of model support, try other values")
    return f, low.bsize, low.pars, np.array(low.init)


class MCMCRun(object):
    """src/samplers.jl:8-20; with chains > 1 every per-sample array gains a trailing chain axis"""

    def __init__(self, steps, burnin):
        self.acceptRate = np.nan
        self.time = np.nan
        self.steps, self.burnin, self.samples = steps, burnin, steps - burnin
        self.ess = (np.nan, np.nan)
        self.essBySec = (np.nan, np.nan)
        self.loglik, self.accept = None, None
        self.params, self.misc = {}, {}

    def __repr__(self):
        head = " ".join("%s%s" % (k, tuple(v.shape)) for k, v in self.params.items())
        return ("%s\nTime %.1f sec., Accept rate %.1f %%, Eff. samples (%d, %d), Eff. samples per sec. (%d, %d)"
                % (head, self.time, 100 * self.acceptRate, round(self.ess[0]), round(self.ess[1]),
                   round(self.essBySec[0]), round(self.essBySec[1])))


def _ess(series):
    """lag-1 effective sample size, src/samplers.jl:356-359"""
    if len(series) < 3:
        return float(len(series))
    a, b = series[1:], series[:-1]
    var = np.var(series, ddof=1)
    if var == 0:
        return 0.0
    cov = np.sum((a - a.mean()) * (b - b.mean())) / (len(a) - 1)
    fac = abs(cov) / var
    return len(series) * max(0.0, 1.0 - fac) / (1.0 + fac)


def ess_from_draws(d):
    """calcStats! (src/samplers.jl:352-374) on stored draws d[samples][C][M] -> ess[C][M]: the lag-1 formula for every
    (chain, component) series at once.  The samplers use the statistics accumulated on the device instead (`_finish`);
    this is the same formula on the host, for checking them."""
    n = d.shape[0]
    a, b = d[1:], d[:-1]
    cov = np.sum((a - a.mean(0)) * (b - b.mean(0)), axis=0) / (n - 2)
    var = np.var(d, axis=0, ddof=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        fac = np.abs(cov) / var
        return np.where(var > 0, n * np.maximum(0.0, 1.0 - fac) / (1.0 + fac), 0.0)


def _finish(res, pars, raw, chains, t0, store_draws):
    res.time = time.time() - t0
    res.acceptRate = float(np.mean(raw["accept_rate"]))
    res.misc["accept_rate_by_chain"] = raw["accept_rate"]
    res.misc["device_ms"] = raw["device_ms"]
    res.misc["n_evals"], res.misc["n_leapfrogs"] = raw["n_evals"], raw["n_leapfrogs"]
    res.misc["n_kernel_launches"] = raw["n_kernel_launches"]
    res.misc["chain_mean"], res.misc["chain_m2"] = raw["mean"], raw["m2"]
    n = res.samples
    if chains > 1 and n > 1:
        # Gelman-Rubin potential scale reduction from the per-chain moments
        W = np.mean(raw["m2"] / (n - 1), axis=0)
        B = n * np.var(raw["mean"], axis=0, ddof=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            res.misc["rhat"] = np.sqrt(((n - 1) / n * W + B / n) / W)
    if n >= 3 and "lag1" in raw:
        # calcStats! (src/samplers.jl:352-374) from the statistics the device accumulated -- no draws needed.  With
        # y_t = x_t - x_0:  a = y[1:], b = y[:-1],  cov = (sum a*b - sum(a)*sum(b)/(n-1)) / (n-2),  var = m2/(n-1)
        sy = n * (raw["mean"] - raw["first"])
        sb = sy - (raw["final_state"] - raw["first"])
        cov = (raw["lag1"] - sy * sb / (n - 1)) / (n - 2)
        var = raw["m2"] / (n - 1)
        with np.errstate(divide="ignore", invalid="ignore"):
            fac = np.abs(cov) / var
            ess = np.where(var > 0, n * np.maximum(0.0, 1.0 - fac) / (1.0 + fac), 0.0)      # [C][M]
        per_comp = ess.sum(axis=0)
        res.misc["ess_by_chain"] = ess
        res.ess = (float(per_comp.min()), float(per_comp.max()))
        res.essBySec = (res.ess[0] / res.time, res.ess[1] / res.time)
    if "loglik" in raw:
        res.loglik = raw["loglik"][:, 0] if chains == 1 else raw["loglik"]
        res.accept = raw["accept"][:, 0] if chains == 1 else raw["accept"]
    if store_draws and "draws" in raw:
        d = raw["draws"]            # [samples][C][M]
        res.misc["draws_scm"] = d   # the ABI layout, as received (res.params holds views of it)
        for p in pars:
            block = d[:, :, p.start:p.start + p.count]             # samples x C x count
            arr = np.transpose(block, (2, 0, 1))                      # count x samples x C
            shape = tuple(p.size) + (n,) + ((chains,) if chains > 1 else ())
            if chains == 1:
                arr = arr[:, :, 0]
            res.params[p.sym] = arr.reshape(shape, order="F") if len(p.size) == 2 else arr.reshape(shape)
        if n >= 3 and "lag1" not in raw:
            res.misc["ess_by_chain"] = ess = ess_from_draws(d)
            per_comp = ess.sum(axis=0)
            res.ess = (float(per_comp.min()), float(per_comp.max()))
            res.essBySec = (res.ess[0] / res.time, res.ess[1] / res.time)
    return res


def _split_kwargs(kw):
    opts = {k: kw.pop(k) for k in list(kw) if k in RESERVED}
    return opts, list(kw.items())


def _check_steps(steps, burnin):
    """checkSteps, src/samplers.jl:324-327"""
    assert burnin >= 0, "Burnin rounds (%d) should be >= 0" % burnin
    assert steps > burnin, "Steps (%d) should be > to burnin (%d)" % (steps, burnin)


def _sample(kind, model, kw, defaults):
    t0 = time.time()
    opts, init = _split_kwargs(kw)
    for k, v in defaults.items():
        opts.setdefault(k, v)
    steps, burnin = int(round(opts["steps"])), int(round(opts["burnin"]))
    _check_steps(steps, burnin)
    chains = int(opts.get("chains", 1))
    store = bool(opts.get("store_draws", True))
    lookup = _DataLookup(opts.get("data"), _caller_globals(3))
    low, dev = _build(model, init, lookup, kind != "rwm", chains, opts.get("device", 0),
                      opts.get("ref_adjoint", "assign"), shard=opts.get("shard"), comm=opts.get("comm"),
                      sufficient_stats=opts.get("sufficient_stats", False))
    try:
        raw = dev.run(kind, np.array(low.init), chains, steps, burnin, isteps=int(round(opts.get("isteps", 1))),
                      stepsize=float(opts.get("stepsize", 0.0)), seed=int(opts.get("seed", 1)), store_draws=store,
                      chain_offset=int(opts.get("chain_offset", 0)))
    except capi.SMCError as e:
        if e.status == -4:
            raise AssertionError(str(e))
        raise
    finally:
        dev.close()
    res = MCMCRun(steps, burnin)
    if kind == "nuts":
        res.misc["epsilon"] = raw["epsilon"][:, 0] if chains == 1 else raw["epsilon"]
        res.misc["jmax"] = raw["jmax"][:, 0] if chains == 1 else raw["jmax"]
    return _finish(res, low.pars, raw, chains, t0, store)


def simpleRWM(model, **kw):
    return _sample("rwm", model, kw, dict(steps=1000, burnin=100))


def simpleHMC(model, **kw):
    return _sample("hmc", model, kw, dict(steps=1000, burnin=100, isteps=2, stepsize=1e-3))


def simpleNUTS(model, **kw):
    return _sample("nuts", model, kw, dict(steps=1000, burnin=100))
