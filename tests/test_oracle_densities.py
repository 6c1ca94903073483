"""Densities of the oracle against exact closed forms evaluated with mpmath at 50 digits.  The reference holds no
known-answer values for its densities (libRmath is un-vendored): these are the substitute pins."""
import mpmath as mp
import numpy as np
import pytest

from oracle import densities as D

mp.mp.dps = 50


def _exact(name, a):
    a = [mp.mpf(float(v)) for v in a]
    if name == "Normal":
        mu, s, x = a
        return -((x - mu) / s) ** 2 / 2 - mp.log(s) - mp.log(mp.sqrt(2 * mp.pi))
    if name == "Uniform":
        lo, hi, x = a
        return -mp.log(hi - lo)
    if name == "Weibull":
        sh, sc, x = a
        return mp.log(sh / sc) + (sh - 1) * mp.log(x / sc) - (x / sc) ** sh
    if name == "Gamma":
        sh, sc, x = a
        return (sh - 1) * mp.log(x) - x / sc - mp.loggamma(sh) - sh * mp.log(sc)
    if name == "Cauchy":
        mu, sc, x = a
        return -mp.log(mp.pi * sc * (1 + ((x - mu) / sc) ** 2))
    if name == "logNormal":
        mu, s, x = a
        return -mp.log(x * s * mp.sqrt(2 * mp.pi)) - (mp.log(x) - mu) ** 2 / (2 * s * s)
    if name == "Beta":
        p, q, x = a
        return mp.loggamma(p + q) - mp.loggamma(p) - mp.loggamma(q) + (p - 1) * mp.log(x) + (q - 1) * mp.log(1 - x)
    if name == "TDist":
        df, x = a
        return mp.loggamma((df + 1) / 2) - mp.loggamma(df / 2) - mp.log(df * mp.pi) / 2 - (df + 1) / 2 * mp.log(1 + x * x / df)
    if name == "Exponential":
        sc, x = a
        return -x / sc - mp.log(sc)
    if name == "Poisson":
        lam, x = a
        return x * mp.log(lam) - lam - mp.loggamma(x + 1)
    if name == "Binomial":
        n, p, x = a
        return mp.loggamma(n + 1) - mp.loggamma(x + 1) - mp.loggamma(n - x + 1) + x * mp.log(p) + (n - x) * mp.log(1 - p)
    if name == "Bernoulli":
        p, x = a
        return mp.log(p) if x == 1 else mp.log(1 - p)
    raise KeyError(name)


POINTS = {
    "Normal": [(0.0, 1.0, 0.3), (2.5, 0.1, 2.4), (-3.0, 7.0, 12.0)],
    "Uniform": [(0.0, 1.0, 0.5), (-10.0, 12.0, 3.0)],
    "Weibull": [(2.0, 1.0, 0.7), (0.5, 3.0, 2.2), (5.0, 0.2, 0.25)],
    "Gamma": [(2.0, 100.0, 20.0), (2.0, 2.0, 0.1), (0.3, 1.5, 4.0), (30.0, 0.5, 11.0)],
    "Cauchy": [(0.0, 1.0, 0.3), (3.0, 0.2, -4.0)],
    "logNormal": [(0.0, 1.0, 0.7), (1.5, 0.3, 9.0)],
    "Beta": [(2.0, 3.0, 0.3), (0.5, 0.5, 0.99), (15.0, 4.0, 0.8)],
    "TDist": [(3.0, 0.2), (0.7, -5.0), (40.0, 1.5)],
    "Exponential": [(1.0, 0.5), (20.0, 3.0)],
    "Poisson": [(3.5, 2.0), (0.2, 0.0), (50.0, 61.0)],
    "Binomial": [(20.0, 0.3, 5.0), (10.0, 0.99, 10.0), (100.0, 0.5, 37.0)],
    "Bernoulli": [(0.3, 1.0), (0.3, 0.0), (0.999, 0.0)],
}


@pytest.mark.parametrize("name", sorted(POINTS))
def test_density_matches_mpmath(name):
    fn = getattr(D, "logpdf" + name)
    for args in POINTS[name]:
        got = fn(*args)
        want = float(_exact(name, args))
        assert abs(got - want) <= 5e-14 * max(1.0, abs(want)), (name, args, got, want)


def test_vectorised_sum_is_sequential_and_uses_first_array_length():
    x = np.random.default_rng(0).standard_normal(1000)
    got = D.logpdfNormal(0.0, 1.0, x)
    acc = 0.0
    for v in x:
        acc += D.logpdfNormal(0.0, 1.0, float(v))
    assert got == acc


def test_support_and_domain_behaviour():
    from oracle.refmodel import GiveUpEval
    with pytest.raises(GiveUpEval):
        D.logpdfUniform(0.0, 1.0, 2.0)
    with pytest.raises(GiveUpEval):
        D.logpdfGamma(2.0, 1.0, -1.0)
    with pytest.raises(GiveUpEval):
        D.logpdfBernoulli(1.0, 0.0)
    with pytest.raises(AssertionError):
        D.logpdfNormal(0.0, -1.0, 0.0)        # assert in src/distribs.jl:25
    with pytest.raises(ValueError):
        D.logpdfGamma(-2.0, 1.0, 1.0)         # NaN from libRmath -> error(), src/distribs.jl:71-73


def test_digamma_matches_mpmath():
    for x in (0.1, 0.5, 1.0, 2.5, 10.0, 123.4):
        assert abs(D.digamma(x) - float(mp.digamma(x))) < 1e-13 * max(1.0, abs(float(mp.digamma(x))))
