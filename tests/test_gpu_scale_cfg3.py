"""BASELINE.json configs[2] at full size on one GPU: logistic / binomial GLM (examples/binomial.jl:13-17), N = 1e7 rows,
M = 10 columns.  Bernoulli and Binomial(20) responses; C in {1, 4, 64} so that the bulk-copy stream kernel (C <= 4) and the
DMMA kernel (C >= 5) both run.  Checked against NumPy closed forms in long double row blocks (the sums of 1e7 terms) and
through size-independent properties: chain-permutation invariance, row-permutation invariance, additivity over row halves."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from util import LOGISTIC, relerr  # noqa: E402

N, M = 10_000_000, 10
BINOMIAL = """
    vars ~ Normal(0, 1.0)
    prob = 1 / (1. + exp(X * vars))
    Y ~ Binomial(20, prob)
"""


def _closed_form(X, Y, b, ntrials):
    """logp and gradient of one chain in long double, row blocks of 1e6 (reference statement list, SURVEY App. A.2:
    prob = 1/(1+exp(eta)); l = y log p + (n - y) log(1 - p) [+ lchoose]; d l / d eta = -(y - n p) ... with this link sign)"""
    from scipy.special import gammaln
    bl = b.astype(np.longdouble)
    L = np.longdouble(0.0)
    G = np.zeros(M, np.longdouble)
    for r0 in range(0, N, 1_000_000):
        Xb = X[r0:r0 + 1_000_000].astype(np.longdouble)
        y = Y[r0:r0 + 1_000_000].astype(np.longdouble)
        eta = Xb @ bl
        p = 1 / (1 + np.exp(eta))
        L += np.sum(y * np.log(p) + (ntrials - y) * np.log1p(-p))
        if ntrials > 1:
            yy = Y[r0:r0 + 1_000_000]
            L += np.sum((gammaln(ntrials + 1.0) - gammaln(yy + 1.0) - gammaln(ntrials - yy + 1.0)).astype(np.longdouble))
        # d/d eta of [y log p + (n - y) log(1 - p)] with p = 1/(1+e^eta): dp/deta = -p (1 - p)  =>  -(y (1 - p) - (n - y) p) = n p - y
        G += Xb.T @ (ntrials * p - y)
    L += np.sum(-0.5 * bl * bl) - M * np.longdouble(0.91893853320467274178)
    G += -bl
    return float(L), G.astype(np.float64)


@pytest.fixture(scope="module")
def data():
    rng = np.random.default_rng(3)
    X = np.empty((M, N)).T
    X[:, 0] = 1.0
    for j in range(1, M):
        X[:, j] = rng.standard_normal(N)
    b0 = 0.3 * rng.standard_normal(M)
    p = 1.0 / (1.0 + np.exp(X @ b0))
    Yb = (rng.random(N) < p).astype(np.float64)
    Yc = rng.binomial(20, p).astype(np.float64)
    return X, Yb, Yc, b0


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import api
    return api


@pytest.mark.parametrize("family", ["bernoulli", "binomial"])
def test_cfg3_full_size(smc, data, family):
    X, Yb, Yc, b0 = data
    model, Y, ntr = (LOGISTIC, Yb, 1) if family == "bernoulli" else (BINOMIAL, Yc, 20)
    rng = np.random.default_rng(5)
    low, dev = smc._build(model, [("vars", np.zeros(M))], dict(X=X, Y=Y), True, 64, 0, "assign")
    assert low.n_fused == 1
    try:
        b = b0[None, :] + 1e-3 * rng.standard_normal((64, M))
        ref = {c: _closed_form(X, Y, b[c], ntr) for c in (0, 3, 63)}
        results = {}
        for C in (1, 4, 64):                                   # stream-bulk kernel (1, 4) and DMMA kernel (64)
            lp, g = dev.eval(b[:C])
            results[C] = (lp, g)
            for c in [c for c in ref if c < C]:
                lp0, g0 = ref[c]
                assert abs(lp[c] - lp0) <= 1e-12 * abs(lp0), (family, C, c, lp[c], lp0)
                assert relerr(g[c], g0) <= 1e-12, (family, C, c, relerr(g[c], g0))
        # the kernels agree with each other on the shared chains (different summation geometries)
        assert np.max(np.abs(results[1][0] - results[64][0][:1]) / np.abs(results[1][0])) <= 1e-13
        assert relerr(results[4][1], results[64][1][:4]) <= 1e-12
        # chains are independent: permuting them permutes the result bit for bit
        perm = rng.permutation(64)
        lp2, g2 = dev.eval(b[perm])
        assert np.array_equal(lp2, results[64][0][perm]) and np.array_equal(g2, results[64][1][perm])
        # a short HMC run from the truth accepts and moves (cfg3's step size; a count of 20 trials carries 20 times
        # the Fisher information of one Bernoulli draw, so the stable step is sqrt(20) times shorter)
        raw = dev.run("hmc", b[:4], 4, 6, 0, isteps=2, stepsize=0.1 / np.sqrt(N / 1000) / np.sqrt(ntr), seed=3, store_draws=False,
                      want=())
        assert raw["n_leapfrogs"] == 4 * 6 * 2 and raw["accept_rate"].mean() > 0.3
    finally:
        dev.close()


def test_cfg3_additivity_over_row_halves(smc, data):
    """the likelihood is a sum over rows: logp(all rows) - prior == [logp(first half) - prior] + [logp(second half) - prior]
    (what row sharding relies on), at 1e-13 because the three launches sum in different orders"""
    X, Yb, Yc, b0 = data
    b = b0[None, :] + 1e-3 * np.random.default_rng(6).standard_normal((4, M))
    prior = np.sum(-0.5 * b * b, axis=1) - M * 0.91893853320467274178
    parts = []
    for sl in (slice(0, N), slice(0, N // 2), slice(N // 2, N)):
        low, dev = smc._build(LOGISTIC, [("vars", np.zeros(M))], dict(X=np.asfortranarray(X[sl]), Y=Yb[sl]), True, 4, 0, "assign")
        lp, g = dev.eval(b)
        parts.append((lp - prior, g + b))
        dev.close()
    assert np.max(np.abs(parts[0][0] - (parts[1][0] + parts[2][0])) / np.abs(parts[0][0])) <= 1e-13
    assert relerr(parts[0][1], parts[1][1] + parts[2][1]) <= 1e-12
