"""oracle/rmath.py (nmath's published algorithms: stirlerr, bd0, dpois_raw, dbinom_raw and the densities built on
them) against 60-digit mpmath values, and against the textbook closed forms the first round's device code used.
VERDICT r01 item 7: where do the closed forms stop agreeing with libRmath's algorithm?  Answer recorded below:
ordinary arguments 1e-13; large counts / shapes lose up to 1e-8 (cancellation between x log(lambda), lambda and
lgamma(x+1), each ~1e9 at lambda = 1e8, for a result ~ -10), which is why the device now follows nmath's algorithm."""
import math

import mpmath as mp
import numpy as np
import pytest

from oracle import densities as D
from oracle import rmath as R
from test_oracle_densities import _exact

mp.mp.dps = 60

GRID = {
    "Poisson": [(l, x) for l in (0.01, 1.0, 30.0, 1e3, 1e5, 1e8)
                for x in (0.0, 1.0, round(l), round(l + 3 * math.sqrt(l)) + 1, round(2 * l) + 5)],
    "Binomial": [(n, p, x) for n in (1.0, 20.0, 1e3, 1e6) for p in (1e-6, 0.01, 0.3, 0.5, 0.97)
                 for x in (0.0, 1.0, round(n * p), round(n / 2), n - 1, n)],
    "Gamma": [(sh, sc, x) for sh in (0.01, 0.5, 1.0, 2.0, 30.0, 1e4, 1e7) for sc in (1e-3, 1.0, 100.0)
              for x in (1e-10 * sc, 0.1 * sc, sh * sc, 3 * sh * sc + 1)],
    "Beta": [(a, b, x) for a in (0.5, 1.0, 2.0, 2.5, 15.0, 1e4) for b in (0.5, 1.5, 3.0, 400.0)
             for x in (1e-9, 0.01, 0.3, a / (a + b), 0.999)],
    "TDist": [(df, x) for df in (0.1, 0.7, 1.0, 3.0, 40.0, 1e4, 1e8) for x in (0.0, 1e-8, 0.2, 1.5, -30.0, 1e6)],
    "Weibull": [(sh, sc, x) for sh in (0.5, 2.0, 7.0) for sc in (0.2, 3.0) for x in (1e-6, 0.5, 4.0)],
    "Cauchy": [(mu, sc, x) for mu in (-3.0, 0.0) for sc in (0.01, 5.0) for x in (-1e6, 0.1, 40.0)],
    "logNormal": [(m, s, x) for m in (-2.0, 1.5) for s in (0.05, 3.0) for x in (1e-8, 1.0, 1e5)],
    "Exponential": [(sc, x) for sc in (1e-3, 1.0, 1e4) for x in (0.0, 0.5, 1e3)],
}
ORDINARY = {   # the ranges the example models live in: every term O(1e3) at most
    "Poisson": lambda a: a[0] <= 1e3, "Binomial": lambda a: a[0] <= 1e3, "Gamma": lambda a: a[0] <= 30.0,
    "Beta": lambda a: a[0] <= 15.0 and a[1] <= 3.0, "TDist": lambda a: a[0] <= 40.0 and abs(a[1]) <= 30.0,
}


def _rel(got, want):
    return abs(got - want) / max(1.0, abs(want))


@pytest.mark.parametrize("name", sorted(GRID))
def test_rmath_algorithms_match_exact_values(name):
    worst = 0.0
    for a in GRID[name]:
        got, want = R.LOGPDF[name](*a), float(_exact(name, a))
        worst = max(worst, _rel(got, want))
        # Beta: n p = (a + b - 2) x is rounded before bd0 sees it (as in nmath): a few 1e-14 at a = 1e4
        tol = 5e-14 if name == "Beta" else 4e-15
        assert _rel(got, want) <= tol, (name, a, got, want)
    assert worst <= 5e-14


@pytest.mark.parametrize("name", sorted(D.CLOSED_FORM))
def test_closed_forms_agree_in_the_ordinary_range_and_drift_outside(name):
    inside = outside = 0.0
    for a in GRID[name]:
        r, c = R.LOGPDF[name](*a), float(D.closed_form(name, *a))
        if ORDINARY[name](a):
            inside = max(inside, _rel(c, r))
        else:
            outside = max(outside, _rel(c, r))
    assert inside <= 1e-12, (name, inside)
    # the finding that made the device follow nmath's algorithm: outside the ordinary range the textbook forms are
    # worse than the 1e-12 parity bar for four of the five
    if name != "Beta":
        assert outside > 1e-12, (name, outside)
    assert outside < 1e-7


def test_stirlerr_and_bd0_building_blocks():
    for n in (0.5, 1.0, 7.5, 15.0, 15.5, 20.0, 36.0, 81.0, 501.0, 1e6):
        m = mp.mpf(n)
        want = float(mp.loggamma(m + 1) - (m + mp.mpf(1) / 2) * mp.log(m) + m - mp.log(mp.sqrt(2 * mp.pi)))
        assert abs(R.stirlerr(n) - want) <= 3e-16 * max(1.0, abs(want) / 1e-3) + 1e-17, (n, R.stirlerr(n), want)
    assert abs(R.stirlerr(3.3) - float(mp.loggamma(4.3) - 3.8 * mp.log(3.3) + 3.3 - mp.log(mp.sqrt(2 * mp.pi)))) < 1e-14
    for x, M in ((10.0, 10.5), (1e8, 1e8 + 3e4), (5.0, 50.0), (1e-3, 1.1e-3), (7.0, 7.0)):
        want = float(mp.mpf(x) * mp.log(mp.mpf(x) / mp.mpf(M)) + mp.mpf(M) - mp.mpf(x))
        assert abs(R.bd0(x, M) - want) <= 4e-16 * abs(want) + 1e-300, (x, M)


def test_edge_cases_follow_r():
    inf = math.inf
    assert R.dbinom(0.0, 10.0, 0.0) == 0.0 and R.dbinom(3.0, 10.0, 0.0) == -inf and R.dbinom(10.0, 10.0, 1.0) == 0.0
    assert R.dbinom(11.0, 10.0, 0.5) == -inf and R.dbinom(2.5, 10.0, 0.5) == -inf and math.isnan(R.dbinom(1.0, 10.0, 1.5))
    assert math.isnan(R.dbinom(1.0, 2.5, 0.5))
    assert R.dpois(0.0, 0.0) == 0.0 and R.dpois(1.0, 0.0) == -inf and R.dpois(-1.0, 2.0) == -inf and math.isnan(R.dpois(1.0, -2.0))
    assert R.dgamma(0.0, 0.5, 1.0) == inf and R.dgamma(0.0, 2.0, 1.0) == -inf and R.dgamma(0.0, 1.0, 4.0) == -math.log(4.0)
    assert R.dgamma(-1.0, 2.0, 1.0) == -inf and math.isnan(R.dgamma(1.0, -2.0, 1.0)) and math.isnan(R.dgamma(1.0, 2.0, 0.0))
    assert R.dbeta(0.0, 2.0, 3.0) == -inf and R.dbeta(0.0, 1.0, 3.0) == math.log(3.0) and R.dbeta(1.0, 3.0, 0.5) == inf
    assert R.dbeta(1.5, 2.0, 2.0) == -inf and math.isnan(R.dbeta(0.5, 0.0, 2.0))
    assert math.isnan(R.dt(0.0, 0.0)) and R.dt(inf, 3.0) == -inf
    assert R.dt(0.7, inf) == -(R.M_LN_SQRT_2PI + 0.5 * 0.49)
    assert R.dweibull(-1.0, 2.0, 1.0) == -inf and R.dweibull(0.0, 0.5, 1.0) == inf and math.isnan(R.dweibull(1.0, 0.0, 1.0))
    assert R.dlnorm(0.0, 0.0, 1.0) == -inf and R.dexp(-1.0, 2.0) == -inf and math.isnan(R.dcauchy(0.0, 0.0, -1.0))


def test_vectorised_oracle_uses_rmath():
    x = np.array([0.0, 3.0, 7.0])
    got = D.elementwise("Poisson", 2.5, x)
    assert [float(v) for v in got] == [R.dpois(float(v), 2.5) for v in x]
    assert D.logpdfPoisson(2.5, x) == float(np.cumsum(got)[-1])
