"""ORACLE (test infrastructure only): the nine libRmath densities the reference binds by `ccall`
(/root/reference/src/distribs.jl:40-48 -- dweibull dbinom dgamma dcauchy dlnorm dbeta dpois dt dexp, give_log = 1),
restated from R's published nmath ALGORITHMS rather than from closed forms:

    C. Loader (2000), "Fast and accurate computation of binomial probabilities": the saddle-point form
        log dbinom = stirlerr(n) - stirlerr(x) - stirlerr(n-x) - bd0(x, np) - bd0(n-x, nq) - 0.5 log(2 pi x (n-x)/n)
    with  stirlerr(n) = log n! - log( sqrt(2 pi n) (n/e)^n )   (table for n <= 15 in steps of 1/2, asymptotic series beyond)
    and   bd0(x, M)   = x log(x/M) + M - x                      (series in v = (x-M)/(x+M) when |x-M| < 0.1 (x+M)),
    reused by dpois_raw, and through them by dgamma (= dpois_raw(shape-1, x/scale)/scale), dbeta (= dbinom_raw for
    a, b > 2) and dt.  dweibull, dcauchy, dlnorm and dexp are single closed-form expressions in nmath and are restated
    as such.

libRmath is NOT vendored in /root/reference and no version is pinned (REQUIRE is empty), so this module is an
*alternative evaluator* beside oracle/densities.py (closed forms = what the device computes): tests/test_oracle_rmath.py
diffs the two over ordinary and extreme arguments and states where the closed forms stop agreeing with libRmath's
algorithm to 1e-12.  The stirlerr half-integer table is generated from its definition with 60-digit arithmetic
(mpmath) instead of being typed in.  Scalar code (math.*), one value per call.
"""
import math

M_LN_SQRT_2PI = 0.918938533204672741780329736406
M_LN_2PI = 1.837877066409345483560659472811
M_2PI = 6.283185307179586476925286766559
DBL_MIN = 2.2250738585072014e-308
DBL_EPSILON = 2.220446049250313e-16
NEG_INF = -math.inf


def _halves_table():
    import mpmath as mp
    old = mp.mp.dps
    mp.mp.dps = 60
    try:
        out = [0.0]
        for k in range(1, 31):
            n = mp.mpf(k) / 2
            out.append(float(mp.loggamma(n + 1) - (n + mp.mpf(1) / 2) * mp.log(n) + n - mp.log(mp.sqrt(2 * mp.pi))))
        return out
    finally:
        mp.mp.dps = old


_SFERR_HALVES = _halves_table()
_S0, _S1, _S2, _S3, _S4 = 1 / 12.0, 1 / 360.0, 1 / 1260.0, 1 / 1680.0, 1 / 1188.0


def stirlerr(n):
    """nmath/stirlerr.c: error of Stirling's formula"""
    if n <= 15.0:
        nn = n + n
        if nn == int(nn):
            return _SFERR_HALVES[int(nn)]
        return math.lgamma(n + 1.0) - (n + 0.5) * math.log(n) + n - M_LN_SQRT_2PI
    nn = n * n
    if n > 500:
        return (_S0 - _S1 / nn) / n
    if n > 80:
        return (_S0 - (_S1 - _S2 / nn) / nn) / n
    if n > 35:
        return (_S0 - (_S1 - (_S2 - _S3 / nn) / nn) / nn) / n
    return (_S0 - (_S1 - (_S2 - (_S3 - _S4 / nn) / nn) / nn) / nn) / n


def bd0(x, np_):
    """nmath/bd0.c: x log(x/np) + np - x without cancellation near x = np"""
    if not (math.isfinite(x) and math.isfinite(np_)) or np_ == 0.0:
        return math.nan
    if abs(x - np_) < 0.1 * (x + np_):
        v = (x - np_) / (x + np_)
        s = (x - np_) * v
        if abs(s) < DBL_MIN:
            return s
        ej = 2 * x * v
        v = v * v
        for j in range(1, 1000):
            ej *= v
            s1 = s + ej / ((j << 1) + 1)
            if s1 == s:
                return s1
            s = s1
    return x * math.log(x / np_) + np_ - x


def dpois_raw(x, lam):
    """nmath/dpois.c dpois_raw, log scale"""
    if lam == 0:
        return 0.0 if x == 0 else NEG_INF
    if not math.isfinite(lam):
        return NEG_INF
    if x < 0:
        return NEG_INF
    if x <= lam * DBL_MIN:
        return -lam
    if lam < x * DBL_MIN:
        if not math.isfinite(x):
            return NEG_INF
        return -lam + x * math.log(lam) - math.lgamma(x + 1)
    return -0.5 * math.log(M_2PI * x) + (-stirlerr(x) - bd0(x, lam))


def dbinom_raw(x, n, p, q):
    """nmath/dbinom.c dbinom_raw, log scale"""
    if p == 0:
        return 0.0 if x == 0 else NEG_INF
    if q == 0:
        return 0.0 if x == n else NEG_INF
    if x == 0:
        if n == 0:
            return 0.0
        return -bd0(n, n * q) - n * p if p < 0.1 else n * math.log(q)
    if x == n:
        return -bd0(n, n * p) - n * q if q < 0.1 else n * math.log(p)
    if x < 0 or x > n:
        return NEG_INF
    lc = stirlerr(n) - stirlerr(x) - stirlerr(n - x) - bd0(x, n * p) - bd0(n - x, n * q)
    lf = M_LN_2PI + math.log(x) + math.log1p(-x / n)
    return lc - 0.5 * lf


def _nonint(x):
    return abs(x - round(x)) > 1e-7 * max(1.0, abs(x))


def dbinom(x, n, p):
    if math.isnan(x) or math.isnan(n) or math.isnan(p):
        return x + n + p
    if p < 0 or p > 1 or n < 0 or _nonint(n):
        return math.nan
    if _nonint(x) or x < 0 or not math.isfinite(x):
        return NEG_INF
    return dbinom_raw(float(round(x)), float(round(n)), p, 1 - p)


def dpois(x, lam):
    if math.isnan(x) or math.isnan(lam):
        return x + lam
    if lam < 0:
        return math.nan
    if _nonint(x) or x < 0 or not math.isfinite(x):
        return NEG_INF
    return dpois_raw(float(round(x)), lam)


def dgamma(x, shape, scale):
    if math.isnan(x) or math.isnan(shape) or math.isnan(scale):
        return x + shape + scale
    if shape < 0 or scale <= 0:
        return math.nan
    if x < 0:
        return NEG_INF
    if shape == 0:
        return math.inf if x == 0 else NEG_INF
    if x == 0:
        if shape < 1:
            return math.inf
        if shape > 1:
            return NEG_INF
        return -math.log(scale)
    if shape < 1:
        pr = dpois_raw(shape, x / scale)
        return pr + (math.log(shape / x) if math.isfinite(shape / x) else math.log(shape) - math.log(x))
    return dpois_raw(shape - 1, x / scale) - math.log(scale)


def lbeta(a, b):
    """nmath/lbeta.c: log B(a, b) without cancellation for large arguments.  nmath's lgammacor(x) (the correction
    lgamma(x) - ((x - 1/2) log x - x + log sqrt(2 pi)), x >= 10) is the same function as stirlerr(x)."""
    p, q = min(a, b), max(a, b)
    if p < 0:
        return math.nan
    if p == 0:
        return math.inf
    if not math.isfinite(q):
        return NEG_INF
    if p >= 10:
        corr = stirlerr(p) + stirlerr(q) - stirlerr(p + q)
        return math.log(q) * -0.5 + M_LN_SQRT_2PI + corr + (p - 0.5) * math.log(p / (p + q)) + q * math.log1p(-p / (p + q))
    if q >= 10:
        corr = stirlerr(q) - stirlerr(p + q)
        return math.lgamma(p) + corr + p - p * math.log(p + q) + (q - 0.5) * math.log1p(-p / (p + q))
    return math.log(math.gamma(p) * (math.gamma(q) / math.gamma(p + q)))


def dbeta(x, a, b):
    if math.isnan(x) or math.isnan(a) or math.isnan(b):
        return x + a + b
    if a <= 0 or b <= 0:
        return math.nan
    if x < 0 or x > 1:
        return NEG_INF
    if x == 0:
        if a > 1:
            return NEG_INF
        if a < 1:
            return math.inf
        return math.log(b)
    if x == 1:
        if b > 1:
            return NEG_INF
        if b < 1:
            return math.inf
        return math.log(a)
    if a <= 2 or b <= 2:
        return (a - 1) * math.log(x) + (b - 1) * math.log1p(-x) - lbeta(a, b)
    return math.log(a + b - 1) + dbinom_raw(a - 1, a + b - 2, x, 1 - x)


def dt(x, n):
    if math.isnan(x) or math.isnan(n):
        return x + n
    if n <= 0:
        return math.nan
    if not math.isfinite(x):
        return NEG_INF
    if not math.isfinite(n):
        return -(M_LN_SQRT_2PI + 0.5 * x * x)
    t = -bd0(n / 2.0, (n + 1) / 2.0) + stirlerr((n + 1) / 2.0) - stirlerr(n / 2.0)
    x2n = x * x / n
    if x2n > 1.0 / DBL_EPSILON:
        ax = abs(x)
        l_x2n = math.log(ax) - math.log(n) / 2.0
        u = n * l_x2n
    elif x2n > 0.2:
        l_x2n = math.log(1 + x2n) / 2.0
        u = n * l_x2n
    else:
        l_x2n = math.log1p(x2n) / 2.0
        u = -bd0(n / 2.0, (n + x * x) / 2.0) + x * x / 2.0
    return t - u - (M_LN_SQRT_2PI + l_x2n)


def dweibull(x, shape, scale):
    if shape <= 0 or scale <= 0:
        return math.nan
    if x < 0 or not math.isfinite(x):
        return NEG_INF
    if x == 0 and shape < 1:
        return math.inf
    tmp1 = math.pow(x / scale, shape - 1)
    tmp2 = tmp1 * (x / scale)
    v = shape * tmp1 / scale
    return -tmp2 + (math.log(v) if v > 0 else NEG_INF)


def dcauchy(x, location, scale):
    if scale <= 0:
        return math.nan
    y = (x - location) / scale
    return -math.log(math.pi * scale * (1.0 + y * y))


def dlnorm(x, meanlog, sdlog):
    if sdlog <= 0:
        return math.nan
    if x <= 0:
        return NEG_INF
    y = (math.log(x) - meanlog) / sdlog
    return -(M_LN_SQRT_2PI + 0.5 * y * y + math.log(x * sdlog))


def dexp(x, scale):
    if scale <= 0:
        return math.nan
    if x < 0:
        return NEG_INF
    return (-x / scale) - math.log(scale)


# the reference rotates the variate to libRmath's first argument (src/distribs.jl:59): logpdfX(params..., x)
LOGPDF = {
    "Weibull": lambda sh, sc, x: dweibull(x, sh, sc),
    "Binomial": lambda n, p, x: dbinom(x, n, p),
    "Gamma": lambda sh, sc, x: dgamma(x, sh, sc),
    "Cauchy": lambda mu, sc, x: dcauchy(x, mu, sc),
    "logNormal": lambda lm, ls, x: dlnorm(x, lm, ls),
    "Beta": lambda a, b, x: dbeta(x, a, b),
    "Poisson": lambda lam, x: dpois(x, lam),
    "TDist": lambda df, x: dt(x, df),
    "Exponential": lambda sc, x: dexp(x, sc),
}
