"""ORACLE (test infrastructure only).  Restatement of /root/reference/src/samplers.jl with the randomness
supplied by the caller (so that device trajectories can be compared draw for draw):

  Sample / calc! / update! / uturn / leapFrog   src/samplers.jl:37-59
  simpleRWM main loop + R.A.M. adaptation       src/samplers.jl:89-113
  simpleHMC main loop                           src/samplers.jl:146-167
  simpleNUTS (buildTree, dual averaging)        src/samplers.jl:199-310
  calcStats! lag-1 ESS                          src/samplers.jl:352-374

`rng` objects expose randn(n) and rand(); `ReplayRNG` replays recorded arrays, `NumpyRNG` draws fresh ones.
"""
import copy
import math

import numpy as np


class NumpyRNG(object):
    def __init__(self, seed):
        self.g = np.random.default_rng(seed)

    def randn(self, n):
        return self.g.standard_normal(n)

    def rand(self):
        return float(self.g.random())


class ReplayRNG(object):
    """normals[step] is the randn(M) of that step; uniforms[step] the successive rand() of that step"""

    def __init__(self, normals, uniforms):
        self.normals, self.uniforms = normals, uniforms
        self.step, self.k = -1, 0

    def begin_step(self, i):
        self.step, self.k = i, 0

    def randn(self, n):
        return np.array(self.normals[self.step][:n], dtype=np.float64)

    def rand(self):
        u = float(self.uniforms[self.step][self.k])
        self.k += 1
        return u


class Sample(object):
    """src/samplers.jl:37-44"""

    def __init__(self, beta):
        self.beta = np.array(beta, dtype=np.float64)
        self.grad = np.zeros(0)
        self.v = np.zeros(0)
        self.llik = math.nan
        self.H = math.nan


def calc(s, ll):
    s.llik, s.grad = ll(s.beta)


def update(s):
    s.H = s.llik - np.dot(s.v, s.v) / 2


def uturn(s1, s2):
    return np.dot(s2.beta - s1.beta, s1.v) < 0.0 or np.dot(s2.beta - s1.beta, s2.v) < 0.0


def leap_frog(s, ve, ll):
    n = copy.deepcopy(s)
    n.v = n.v + n.grad * ve / 2.0
    n.beta = n.beta + ve * n.v
    calc(n, ll)
    n.v = n.v + n.grad * ve / 2.0
    update(n)
    return n


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


def simple_hmc(ll_func, beta, rng, steps=1000, burnin=100, isteps=2, stepsize=1e-3):
    """src/samplers.jl:124-171 -> dict(draws [samples x M], loglik, accept)"""
    state0 = Sample(beta)
    calc(state0, ll_func)
    assert math.isfinite(state0.llik), "Initial values out of model support, try other values"
    draws, loglik, accept = [], [], []
    nleap = 0
    for i in range(steps):
        if hasattr(rng, "begin_step"):
            rng.begin_step(i)
        state0.v = rng.randn(len(state0.beta))
        update(state0)
        state = state0
        j = 1
        while j <= isteps and math.isfinite(state.llik):
            state = leap_frog(state, stepsize, ll_func)
            nleap += 1
            j += 1
        dH = state.H - state0.H
        if rng.rand() < (_exp(dH) if not math.isnan(dH) else math.nan):
            state0 = state
        if i >= burnin:
            draws.append(state0.beta.copy())
            loglik.append(state0.llik)
            accept.append(state0 is state)
    return dict(draws=np.array(draws), loglik=np.array(loglik), accept=np.array(accept, dtype=float), nleap=nleap)


def simple_rwm(ll_func, beta, rng, steps=1000, burnin=100):
    """src/samplers.jl:70-118 (ll_func returns the log-density only)"""
    target_accept = 0.234
    beta = np.array(beta, dtype=np.float64)
    n = len(beta)
    lp = ll_func(beta)
    assert lp != -math.inf, "Initial values out of model support, try other values"
    S = np.eye(n)
    draws, loglik, accept = [], [], []
    for i in range(1, steps + 1):
        if hasattr(rng, "begin_step"):
            rng.begin_step(i - 1)
        jump = rng.randn(n)
        oldbeta = beta.copy()
        beta = beta + S @ jump
        old_lp, lp = lp, ll_func(beta)
        alpha = min(1.0, _exp(lp - old_lp))
        if rng.rand() > alpha:
            lp, beta = old_lp, oldbeta
        if i > burnin:
            draws.append(beta.copy())
            loglik.append(lp)
            accept.append(old_lp != lp)
        eta = min(1.0, n * i ** (-2.0 / 3.0))
        SS = np.outer(jump, jump) / np.dot(jump, jump) * eta * (alpha - target_accept)
        SS = S @ (np.eye(n) + SS) @ S.T
        S = np.linalg.cholesky(SS)          # chol(SS)' : lower factor
    return dict(draws=np.array(draws), loglik=np.array(loglik), accept=np.array(accept, dtype=float), S=S)


def simple_nuts(ll_func, beta, rng, steps=1000, burnin=100, max_depth=8):
    """src/samplers.jl:180-315, including the quirk that the initial-epsilon search never runs
    (state0.H is NaN at :204, so epsilon0 = 1.0 -- SURVEY.md Appendix C, Q1)."""
    state0 = Sample(beta)
    calc(state0, ll_func)
    assert math.isfinite(state0.llik), "Initial values out of model support, try other values"
    n = len(state0.beta)
    if hasattr(rng, "begin_step"):
        rng.begin_step(-1)
    epsilon = 1.0
    # (the reference draws randn(nparams) here for the dead epsilon search, :201)
    delta, nadapt, gam, kappa, t0 = 0.7, 1000, 0.05, 0.75, 10
    hbar, mu, lebar = 0.0, math.log(10 * epsilon), 0.0
    deltamax = 100
    ctx = {"eps": epsilon, "u": 0.0, "H0": 0.0, "nleap": 0}

    def build_tree(state, dirn, j):
        if j == 0:
            s1 = leap_frog(state, dirn * ctx["eps"], ll_func)
            ctx["nleap"] += 1
            n1 = 1 if ctx["u"] <= s1.H else 0
            ok = ctx["u"] < (deltamax + s1.H)
            dH = s1.H - ctx["H0"]
            a = min(1.0, _exp(dH)) if not math.isnan(dH) else math.nan
            return s1, s1, s1, n1, ok, a, 1
        sm, sp, s1, n1, ok1, a1, na1 = build_tree(state, dirn, j - 1)
        if ok1:
            if dirn == -1:
                sm, _, s2, n2, ok2, a2, na2 = build_tree(sm, dirn, j - 1)
            else:
                _, sp, s2, n2, ok2, a2, na2 = build_tree(sp, dirn, j - 1)
            if n2 + n1 > 0:
                if rng.rand() <= n2 / (n2 + n1):
                    s1 = s2
            else:
                rng.rand()      # 0/0 = NaN in the reference: the draw happens, the comparison is false
            a1 += a2
            na1 += na2
            ok1 = ok2 and not uturn(sm, sp)
            n1 += n2
        return sm, sp, s1, n1, ok1, a1, na1

    draws, loglik, accept, eps_hist, jmax = [], [], [], [], []
    for i in range(1, steps + 1):
        if hasattr(rng, "begin_step"):
            rng.begin_step(i - 1)
        state0.v = rng.randn(n)
        update(state0)
        ctx["u"] = math.log(rng.rand()) + state0.H
        ctx["H0"] = state0.H
        state = sm = sp = state0
        j, nn, s = 0, 1, True
        alpha, nalpha = 0.0, 1
        while s and j < max_depth:
            dirn = (1.0 if rng.rand() > 0.5 else 0.0) * 2.0 - 1.0
            if dirn == -1:
                sm, _, s1, n1, ok1, alpha, nalpha = build_tree(sm, dirn, j)
            else:
                _, sp, s1, n1, ok1, alpha, nalpha = build_tree(sp, dirn, j)
            if ok1 and rng.rand() < n1 / nn:
                state = s1
            nn += n1
            j += 1
            s = ok1 and not uturn(sm, sp)
        if i <= nadapt:
            hbar = hbar * (1 - 1 / (i + t0)) + (delta - alpha / nalpha) / (i + t0)
            le = mu - math.sqrt(i) / gam * hbar
            lebar = i ** (-kappa) * le + (1 - i ** (-kappa)) * lebar
            ctx["eps"] = math.exp(le)
        else:
            ctx["eps"] = math.exp(lebar)
        if i > burnin:
            draws.append(state.beta.copy())
            loglik.append(state.llik)
            accept.append(state is not state0)
        eps_hist.append(ctx["eps"])
        jmax.append(j)
        state0 = state
    return dict(draws=np.array(draws), loglik=np.array(loglik), accept=np.array(accept, dtype=float),
                epsilon=np.array(eps_hist), jmax=np.array(jmax), nleap=ctx["nleap"])


def ess_lag1(series):
    """calcStats!, src/samplers.jl:356-359"""
    a, b = series[1:], series[:-1]
    cov = np.sum((a - a.mean()) * (b - b.mean())) / (len(a) - 1)
    fac = abs(cov) / np.var(series, ddof=1)
    return len(series) * max(0.0, 1.0 - fac) / (1.0 + fac)
