import re

import numpy as np


def normalise_statements(lines):
    """rename gensym'd names (##tmp#12, ##__acc#4) to t1.. / a1.. in order of first appearance"""
    names = {}
    counts = {"t": 0, "a": 0}

    def sub(m):
        full = m.group(0)
        kind = "a" if "__acc" in full else "t"
        if full not in names:
            counts[kind] += 1
            names[full] = "%s%d" % (kind, counts[kind])
        return names[full]
    return [re.sub(r"##[A-Za-z_]+#\d+", sub, ln) for ln in lines]


def read_golden(path):
    sections, cur = {}, None
    for ln in open(path):
        ln = ln.rstrip("\n")
        if not ln or ln.startswith("#"):
            continue
        if ln.startswith("["):
            cur = ln.strip("[]")
            sections[cur] = []
        else:
            sections[cur].append(ln)
    return sections


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)


def linear_data(N=1000, M=10, seed=1):
    """examples/linear.jl:6-11 with NumPy default_rng(seed) (Julia's srand(1) stream cannot be reproduced)"""
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
    beta0 = rng.standard_normal(M)
    Y = X @ beta0 + rng.standard_normal(N)
    return X, Y, beta0


def logistic_data(N=1000, M=10, seed=3):
    """examples/binomial.jl:6-10 -- note P(Y=1) = 1/(1+exp(X*beta0))"""
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
    beta0 = rng.standard_normal(M)
    Y = (rng.random(N) < 1.0 / (1.0 + np.exp(X @ beta0))).astype(np.float64)
    return X, Y, beta0


def ou_data(T=1000, seed=5):
    """examples/ornstein.jl:7-16"""
    rng = np.random.default_rng(seed)
    mu0, tau0, sigma0 = 10.0, 20.0, 0.1
    x = np.empty(T)
    x[0] = 1.0
    f = np.exp(-1.0 / tau0)
    for i in range(1, T):
        x[i] = x[i - 1] * f + mu0 * (1 - f) + sigma0 * rng.standard_normal()
    return x


def hierarchical_data(seed=4):
    """examples/hierarchical.jl:7-22"""
    rng = np.random.default_rng(seed)
    N, D, L = 50, 4, 5
    mu0 = rng.standard_normal((1, L))
    sigma0 = rng.random((1, L))
    beta0 = rng.standard_normal((D, L)) * sigma0 + mu0
    ll = rng.integers(1, D + 1, N)
    X = rng.standard_normal((N, L))
    Y = (rng.random(N) < 1.0 / (1.0 + np.exp(-(beta0[ll - 1, :] * X).sum(axis=1)))).astype(np.float64)
    return dict(oneD=np.ones(D), oneL=np.ones(L), ll=ll, X=X, Y=Y), (N, D, L)


def sleepstudy_data(seed=6):
    """lme4 sleepstudy-shaped synthetic data (examples/sleepstudy.jl:25-31; RDatasets is unavailable)"""
    rng = np.random.default_rng(seed)
    nsub, ndays = 18, 10
    days = np.tile(np.arange(ndays, dtype=np.float64), nsub)
    sub = np.repeat(np.arange(nsub), ndays)
    submatrix = np.zeros((nsub * ndays, nsub))
    submatrix[np.arange(nsub * ndays), sub] = 1.0
    ri, rs = 25 * rng.standard_normal(nsub), 6 * rng.standard_normal(nsub)
    reaction = 250 + 10 * days + ri[sub] + rs[sub] * days + 25 * rng.standard_normal(nsub * ndays)
    return dict(days=days, reaction=reaction, submatrix=submatrix), nsub


LINEAR = """
    beta ~ Normal(0, 1.0)
    resid = Y - X * beta
    resid ~ Normal(0, 1.0)
"""
LOGISTIC = """
    vars ~ Normal(0, 1.0)
    prob = 1 / (1. + exp(X * vars))
    Y ~ Bernoulli(prob)
"""
ORNSTEIN = """
    tau ~ Gamma(2, 100)
    sigma ~ Gamma(2, 2)
    mu ~ Normal(0, 100)
    fac = exp(- 1 / tau)
    resid = x[2:end] - x[1:end-1] * fac - mu * (1. - fac)
    resid ~ Normal(0, sigma)
"""
HIERARCHICAL = """
    mu ~ Normal(0, 1)
    sigma ~ Weibull(2, 1)
    beta ~ Normal(oneD * mu, oneD * sigma)
    effect = (beta[ll,:] .* X) * oneL
    prob = 1. / ( 1. + exp(- effect) )
    Y ~ Bernoulli(prob)
"""
SLEEPSTUDY2 = """
    r_bar = (intercept + submatrix * rand_int) + (slope + submatrix * rand_slope) .* days
    rand_int ~ Normal(0,100)
    rand_slope ~ Normal(0,100)
    resid = r_bar - reaction
    resid ~ Normal(0,10)
"""
