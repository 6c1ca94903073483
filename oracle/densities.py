"""ORACLE (test infrastructure only).

Log-densities of SimpleMCMC.jl, restated for the CPU:

* logpdfBernoulli / logpdfNormal / logpdfUniform follow /root/reference/src/distribs.jl:12-33 literally
  (including which violations are hard errors -- `assert` -- and which abandon the evaluation).
* The other nine call libRmath (`dweibull dbinom dgamma dcauchy dlnorm dbeta dpois dt dexp`, give_log=1,
  variate rotated to the first argument: src/distribs.jl:40-48,59,65-67).  libRmath is a third-party
  dependency that is NOT vendored under /root/reference and is not version-pinned (REQUIRE is empty,
  `dlopen("libRmath")` by bare name).  Its published algorithms are restated in oracle/rmath.py (Loader's
  saddle-point forms for dbinom / dpois / dgamma / dbeta / dt, single expressions for the other four) and
  that is what `logpdfX` evaluates here; the textbook closed forms are kept as `closed_form(name, ...)` --
  they are what the first round's device code computed and tests/test_oracle_rmath.py records where they
  stop agreeing with nmath's algorithm (1e-8 relative at lambda = 1e8).  R's documented edge cases: outside
  support -> -Inf, invalid parameters -> NaN.  Parity for the absolute values is "unpinned" against the
  reference binary and pinned instead against 50-digit mpmath values (tests/test_oracle_densities.py,
  tests/test_oracle_rmath.py): nmath and this restatement are both within a few ulp of the exact value.
* The vectorised methods (src/distribs.jl:85-110) sum element by element, left to right, over the
  length of the first array argument; `-Inf` abandons the whole evaluation ("give up eval",
  src/distribs.jl:69-70), `NaN` is a hard error (:71-73).
"""
import math

import numpy as np
from scipy import special as _sp

DISTRIBUTIONS = ["Weibull", "Binomial", "Gamma", "Cauchy", "logNormal", "Beta", "Poisson", "TDist",
                 "Exponential", "Uniform", "Normal", "Bernoulli"]

LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


class GiveUp(Exception):
    pass


def digamma(x):
    return _sp.digamma(x)


def _arr(v):
    return np.asarray(v, dtype=np.float64)


# ---- elementwise kernels (broadcast numpy arrays in, array out; -inf = out of support, nan = invalid)
def _normal(mu, sigma, x):
    if np.any(sigma <= 0.0):
        raise AssertionError("calling logpdfNormal with negative or null stdev")
    r = (x - mu) / sigma
    return -r * r * 0.5 - np.log(sigma) - LOG_SQRT_2PI     # distribs.jl:23-28


def _uniform(a, b, x):
    if np.any(~(a < b)):
        raise AssertionError("calling logpdfUniform with lower limit >= upper limit")
    with np.errstate(divide="ignore"):
        return np.where((a <= x) & (x <= b), -np.log(b - a), -np.inf)     # distribs.jl:30-33


def _bernoulli(p, x):
    if np.any(~((0.0 <= p) & (p <= 1.0))):
        raise AssertionError("calling Bernoulli with prob > 1. or < 0.")
    if np.any((x != 0.0) & (x != 1.0)):
        raise ValueError("calling Bernoulli with variable other than 0 or 1 (false or true)")
    with np.errstate(divide="ignore"):
        # x==0: log(1-p) (p==1 -> give up); x==1: log(p) (p==0 -> give up)   distribs.jl:12-21
        return np.where(x == 0.0, np.log(1.0 - p), np.log(p))


def _weibull(sh, sc, x):          # dweibull(x, shape, scale, log)
    with np.errstate(all="ignore"):
        t1 = np.power(x / sc, sh - 1.0)
        t2 = t1 * (x / sc)
        res = -t2 + np.log(sh * t1 / sc)
        res = np.where(x < 0.0, -np.inf, res)
        res = np.where(~np.isfinite(x), -np.inf, res)
        res = np.where((x == 0.0) & (sh < 1.0), np.inf, res)
        res = np.where((sh <= 0.0) | (sc <= 0.0), np.nan, res)
    return res


def _binomial(n, p, x):           # dbinom(x, n, p, log)
    with np.errstate(all="ignore"):
        lc = _sp.gammaln(n + 1.0) - _sp.gammaln(x + 1.0) - _sp.gammaln(n - x + 1.0)
        res = lc + _sp.xlogy(x, p) + _sp.xlog1py(n - x, -p)
        res = np.where((x < 0.0) | (x > n) | (x != np.round(x)), -np.inf, res)
        res = np.where((p < 0.0) | (p > 1.0) | (n < 0.0) | (n != np.round(n)), np.nan, res)
    return res


def _gamma(sh, sc, x):            # dgamma(x, shape, scale, log)
    with np.errstate(all="ignore"):
        res = (sh - 1.0) * np.log(x) - x / sc - _sp.gammaln(sh) - sh * np.log(sc)
        at0 = np.where(sh < 1.0, np.inf, np.where(sh > 1.0, -np.inf, -np.log(sc)))
        res = np.where(x == 0.0, at0, res)
        res = np.where(x < 0.0, -np.inf, res)
        res = np.where(sh == 0.0, np.where(x == 0.0, np.inf, -np.inf), res)
        res = np.where((sh < 0.0) | (sc <= 0.0), np.nan, res)
    return res


def _cauchy(mu, sc, x):           # dcauchy(x, location, scale, log)
    with np.errstate(all="ignore"):
        y = (x - mu) / sc
        res = -np.log(math.pi * sc * (1.0 + y * y))
        res = np.where(sc <= 0.0, np.nan, res)
    return res


def _lognormal(lmu, lsc, x):      # dlnorm(x, meanlog, sdlog, log)
    with np.errstate(all="ignore"):
        y = (np.log(x) - lmu) / lsc
        res = -(LOG_SQRT_2PI + 0.5 * y * y + np.log(x * lsc))
        res = np.where(x <= 0.0, -np.inf, res)
        res = np.where(lsc <= 0.0, np.nan, res)
    return res


def _beta(a, b, x):               # dbeta(x, a, b, log)
    with np.errstate(all="ignore"):
        res = _sp.gammaln(a + b) - _sp.gammaln(a) - _sp.gammaln(b) + _sp.xlogy(a - 1.0, x) + _sp.xlog1py(b - 1.0, -x)
        at0 = np.where(a > 1.0, -np.inf, np.where(a < 1.0, np.inf, np.log(b)))
        at1 = np.where(b > 1.0, -np.inf, np.where(b < 1.0, np.inf, np.log(a)))
        res = np.where(x == 0.0, at0, res)
        res = np.where(x == 1.0, at1, res)
        res = np.where((x < 0.0) | (x > 1.0), -np.inf, res)
        res = np.where((a <= 0.0) | (b <= 0.0), np.nan, res)
    return res


def _poisson(lam, x):             # dpois(x, lambda, log)
    with np.errstate(all="ignore"):
        res = _sp.xlogy(x, lam) - lam - _sp.gammaln(x + 1.0)
        res = np.where((x < 0.0) | (x != np.round(x)), -np.inf, res)
        res = np.where(lam < 0.0, np.nan, res)
    return res


def _tdist(df, x):                # dt(x, df, log)
    with np.errstate(all="ignore"):
        res = _sp.gammaln((df + 1.0) / 2.0) - _sp.gammaln(df / 2.0) - 0.5 * np.log(df * math.pi) \
            - (df + 1.0) / 2.0 * np.log1p(x * x / df)
        res = np.where(df <= 0.0, np.nan, res)
    return res


def _exponential(sc, x):          # dexp(x, scale, log)
    with np.errstate(all="ignore"):
        res = -x / sc - np.log(sc)
        res = np.where(x < 0.0, -np.inf, res)
        res = np.where(sc <= 0.0, np.nan, res)
    return res


def _from_rmath(name):
    from . import rmath
    fn = np.frompyfunc(rmath.LOGPDF[name], rmath.LOGPDF[name].__code__.co_argcount, 1)

    def g(*args):
        return np.asarray(fn(*args), dtype=np.float64)
    return g


CLOSED_FORM = {"Binomial": _binomial, "Gamma": _gamma, "Beta": _beta, "Poisson": _poisson, "TDist": _tdist}


def closed_form(name, *args):
    """the textbook closed form of one of the five saddle-point densities (per element, broadcast)"""
    return CLOSED_FORM[name](*[_arr(a) for a in args])


_binomial, _gamma, _beta, _poisson, _tdist = (_from_rmath(n) for n in ("Binomial", "Gamma", "Beta", "Poisson", "TDist"))

_ELEMENTWISE = {"Normal": _normal, "Uniform": _uniform, "Bernoulli": _bernoulli, "Weibull": _weibull,
                "Binomial": _binomial, "Gamma": _gamma, "Cauchy": _cauchy, "logNormal": _lognormal,
                "Beta": _beta, "Poisson": _poisson, "TDist": _tdist, "Exponential": _exponential}


def elementwise(name, *args):
    """per-element log-density, arguments broadcast (no summation, no exceptions for -inf/nan)"""
    return _ELEMENTWISE[name](*[_arr(a) for a in args])


def _make(name):
    fn = _ELEMENTWISE[name]

    def logpdf(*args):
        from .refmodel import GiveUpEval
        arrays = [a for a in args if isinstance(a, np.ndarray)]
        if arrays:
            n = arrays[0].size           # length of the FIRST array argument, distribs.jl:99
            flat = []
            for a in args:
                if isinstance(a, np.ndarray):
                    f = a.flatten(order="F").astype(np.float64)
                    if f.size < n:
                        raise IndexError("logpdf%s: array argument shorter than the first one" % name)
                    flat.append(f[:n])
                else:
                    flat.append(np.float64(a))
            vals = np.atleast_1d(fn(*flat))
            if vals.shape[0] != n:
                vals = np.broadcast_to(vals, (n,))
        else:
            vals = np.atleast_1d(fn(*[np.float64(a) for a in args]))
        bad = np.isnan(vals) | (vals == -np.inf)
        if bad.any():
            first = int(np.argmax(bad))
            if np.isnan(vals[first]):
                raise ValueError("calling logpdf%s%s returned an error" % (name, tuple(args)))
            raise GiveUpEval("give up eval")
        if vals.size == 0:
            return 0.0
        return float(np.cumsum(vals)[-1])      # strictly left-to-right `res +=`, distribs.jl:100-102
    logpdf.__name__ = "logpdf" + name
    return logpdf


for _n in DISTRIBUTIONS:
    globals()["logpdf" + _n] = _make(_n)
