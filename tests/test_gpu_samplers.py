"""Sampler parity: first-K-step trajectories against the oracle when both consume the same normals/uniforms
(BASELINE.json north_star), the device Philox stream against its NumPy restatement, KS tests on the 17 targets of
test/test2.jl, posterior moments within Monte-Carlo error."""
import math

import numpy as np
import pytest
from scipy import stats

pytestmark = pytest.mark.gpu

from oracle import philox as ophilox  # noqa: E402
from oracle.refmodel import generate_model_function  # noqa: E402
from oracle.samplers import ReplayRNG, simple_hmc, simple_nuts, simple_rwm  # noqa: E402
from util import LINEAR, LOGISTIC, ORNSTEIN, linear_data, logistic_data, ou_data, relerr, sleepstudy_data, SLEEPSTUDY2  # noqa: E402


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    return load_package()


def test_device_philox_matches_oracle(smc):
    from simplemcmc_jl_b200 import capi
    for seed, chain, step in [(1, 0, 0), (12345678901234, 77, 3), (2 ** 63 + 5, 2 ** 33 + 1, 2 ** 32 + 7)]:
        u_dev = capi.philox_uniforms(seed, chain, step, 33)
        u_ref = ophilox.uniforms(seed, chain, step, 33)
        assert np.array_equal(u_dev, u_ref)                       # integer -> double conversion is exact
        z_dev = capi.philox_normals(seed, chain, step, 33)
        z_ref = ophilox.normals(seed, chain, step, 33)
        assert np.allclose(z_dev, z_ref, rtol=1e-13, atol=1e-15)  # log / sincos may differ in the last ulp


def _build(model, init, data, C, gradient=True):
    from simplemcmc_jl_b200 import api
    return api._build(model, init, data, gradient, C, 0, "assign")


def test_hmc_trajectory_parity_linear(smc):
    X, Y, _ = linear_data()
    data = dict(X=X, Y=Y)
    ref = generate_model_function(LINEAR, data=data, gradient=True, beta=np.zeros(10))
    steps, burnin, C = 40, 5, 4
    rng = np.random.default_rng(11)
    normals, unif = rng.standard_normal((steps, C, 10)), rng.random((steps, C, 1))
    low, dev = _build(LINEAR, [("beta", np.zeros(10))], data, C)
    raw = dev.run("hmc", np.zeros(10), C, steps, burnin, isteps=2, stepsize=0.05, inject_normals=normals, inject_uniforms=unif)
    for c in range(C):
        o = simple_hmc(ref, np.zeros(10), ReplayRNG(normals[:, c], unif[:, c]), steps, burnin, 2, 0.05)
        assert relerr(raw["draws"][:, c], o["draws"]) < 1e-12
        assert relerr(raw["loglik"][:, c], o["loglik"]) < 1e-12
        assert np.array_equal(raw["accept"][:, c], o["accept"])
    assert raw["n_leapfrogs"] == C * steps * 2


def test_hmc_stops_trajectory_outside_support(smc):
    """src/samplers.jl:156: the inner loop ends at a non-finite log-density and the proposal is rejected"""
    model = "x ~ Uniform(0, 1)"
    ref = generate_model_function(model, gradient=True, x=0.5)
    steps, C = 25, 3
    rng = np.random.default_rng(5)
    normals, unif = rng.standard_normal((steps, C, 1)), rng.random((steps, C, 1))
    low, dev = _build(model, [("x", 0.5)], {}, C)
    raw = dev.run("hmc", np.array([0.5]), C, steps, 0, isteps=3, stepsize=0.3, inject_normals=normals, inject_uniforms=unif)
    tot = 0
    for c in range(C):
        o = simple_hmc(ref, np.array([0.5]), ReplayRNG(normals[:, c], unif[:, c]), steps, 0, 3, 0.3)
        assert relerr(raw["draws"][:, c], o["draws"]) < 1e-12
        assert np.array_equal(raw["accept"][:, c], o["accept"])
        tot += o["nleap"]
    assert raw["n_leapfrogs"] == tot and tot < C * steps * 3
    assert np.all((raw["draws"] > 0) & (raw["draws"] < 1))


def test_rwm_trajectory_parity_logistic(smc):
    X, Y, _ = logistic_data(500, 6)
    data = dict(X=X, Y=Y)
    ref = generate_model_function(LOGISTIC, data=data, gradient=False, vars=np.zeros(6))
    steps, burnin, C = 60, 10, 3
    rng = np.random.default_rng(12)
    normals, unif = rng.standard_normal((steps, C, 6)), rng.random((steps, C, 1))
    low, dev = _build(LOGISTIC, [("vars", np.zeros(6))], data, C, gradient=False)
    raw = dev.run("rwm", np.zeros(6), C, steps, burnin, inject_normals=normals, inject_uniforms=unif)
    for c in range(C):
        o = simple_rwm(ref, np.zeros(6), ReplayRNG(normals[:, c], unif[:, c]), steps, burnin)
        # the device updates the Cholesky factor by a rank-one update instead of refactorising: equal up to rounding
        assert relerr(raw["draws"][:, c], o["draws"]) < 1e-9
        assert np.array_equal(raw["accept"][:, c], o["accept"])


def test_nuts_trajectory_parity(smc):
    data, nsub = sleepstudy_data()
    X, Y, _ = linear_data(200, 5)
    cases = [(LINEAR, [("beta", np.zeros(5))], dict(X=X, Y=Y)),
             ("x ~ Normal(1, 2)\ny ~ Gamma(2, 1)", [("x", 0.5), ("y", 1.0)], {})]
    for model, init, dat in cases:
        ref = generate_model_function(model, data=dat, gradient=True, init=init)
        M = len(ref.init)
        steps, burnin, C, NU = 30, 5, 3, 270
        rng = np.random.default_rng(13)
        normals, unif = rng.standard_normal((steps, C, M)), rng.random((steps, C, NU))
        low, dev = _build(model, init, dat, C)
        raw = dev.run("nuts", ref.init, C, steps, burnin, inject_normals=normals, inject_uniforms=unif)
        tot = 0
        for c in range(C):
            o = simple_nuts(ref, ref.init, ReplayRNG(normals[:, c], unif[:, c]), steps, burnin)
            assert np.array_equal(raw["jmax"][:, c], o["jmax"]), (raw["jmax"][:, c], o["jmax"])
            # first-K-step parity is tight; later steps only loosely: the reference starts at epsilon = 1 and its dual
            # averaging (no knob to switch it off or to start adapted: nadapt = 1000 is a literal, src/samplers.jl:192)
            # repeatedly overshoots the stability limit of the leapfrog map, which amplifies last-bit differences.
            # tests/test_oracle_nuts_chaos.py shows the same growth between the oracle and a copy of ITSELF whose
            # gradient is perturbed in the last bit: the loose late tolerance is the reference algorithm's own
            # sensitivity, not a device discrepancy.
            K = 12
            assert relerr(raw["epsilon"][:K, c], o["epsilon"][:K]) < 1e-10
            assert relerr(raw["draws"][:K - burnin, c], o["draws"][:K - burnin]) < 1e-10
            assert relerr(raw["epsilon"][:, c], o["epsilon"]) < 1e-4
            assert relerr(raw["draws"][:, c], o["draws"]) < 1e-4
            assert np.array_equal(raw["accept"][:, c], o["accept"])
            tot += o["nleap"]
        assert raw["n_leapfrogs"] == tot


def test_nuts_active_chain_compaction_is_invisible(smc):
    """late rounds of a NUTS iteration evaluate only the chains still building a tree, packed into dense columns.  For a
    model whose evaluation has no cross-element reduction a chain's result cannot depend on its column, so draws, step
    sizes, tree depths and leapfrog counts must be bit-identical to evaluating every chain every round; for a model with
    reductions (their tree depends on the evaluated chain count) the first iterations agree to rounding."""
    import os
    def both(model, init, dat, C, steps):
        out = []
        os.environ["SMC_COMPACT"] = "1"           # (packing is normally reserved for >= 4096 chains)
        for flag in ("1", "0"):
            os.environ["SMC_NO_COMPACT"] = flag
            low, dev = _build(model, init, dat, C)
            out.append(dev.run("nuts", np.array(low.init), C, steps, 0, seed=5))
            dev.close()
        os.environ.pop("SMC_NO_COMPACT")
        os.environ.pop("SMC_COMPACT")
        return out
    full, packed = both("x ~ Normal(1, 2)\ny ~ Gamma(2, 1)", [("x", 0.5), ("y", 1.0)], {}, 700, 40)
    assert full["n_evals"] > 40 * 4                                     # trees of several doublings: packing was exercised
    for k in ("draws", "epsilon", "jmax", "accept", "loglik", "mean", "m2"):
        assert np.array_equal(full[k], packed[k]), k
    assert full["n_leapfrogs"] == packed["n_leapfrogs"] and full["n_evals"] == packed["n_evals"]
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    full, packed = both(SLEEPSTUDY2, init, data, 700, 6)
    assert np.array_equal(full["jmax"][:3], packed["jmax"][:3])
    assert relerr(packed["draws"][:3], full["draws"][:3]) < 1e-8


def test_nuts_device_side_round_loop_is_invisible(smc):
    """SMC_NUTS_DEVICE_LOOP=1 (opt-in, below the packing threshold): an iteration of simpleNUTS is one graph launch whose WHILE
    node repeats the round (pre, evaluation, post) until a kernel finds no chain building its tree (src/samplers.jl:263-311 in
    lockstep): same kernels on the same data in the same order as the host-driven loop, so everything is bit-identical; the
    host loop may only have issued speculative rounds that found no active chain"""
    import os
    data, nsub = sleepstudy_data()
    cases = [("x ~ Normal(1, 2)\ny ~ Gamma(2, 1)", [("x", 0.5), ("y", 1.0)], {}, 200, 40),
             (SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))], data, 64, 12)]
    for model, init, dat, C, steps in cases:
        out = []
        for flag in ("0", "1"):
            os.environ["SMC_NUTS_DEVICE_LOOP"] = flag
            low, dev = _build(model, init, dat, C)
            out.append(dev.run("nuts", np.array(low.init), C, steps, 2, seed=7))
            dev.close()
        os.environ.pop("SMC_NUTS_DEVICE_LOOP")
        host, devl = out
        for k in ("draws", "epsilon", "jmax", "accept", "loglik", "mean", "m2"):
            assert np.array_equal(host[k], devl[k]), k
        assert host["n_leapfrogs"] == devl["n_leapfrogs"] and host["n_evals"] >= devl["n_evals"] > steps
        assert devl["n_kernel_launches"] > 0


def test_nuts_asynchronous_iterations_are_invisible(smc):
    """asynchronous iterations (from 32 chains on): every chain runs through its iterations at its own pace -- a chain whose tree
    is complete starts its next iteration in the next round instead of idling until the longest tree of the iteration is.  RNG
    streams, recorder, step-size adaptation and output positions are keyed by the chain's own iteration number, so everything
    is bit-identical to the iteration-by-iteration loop (SMC_NUTS_ASYNC=0), whether the whole run is asynchronous
    (SMC_NUTS_ASYNC=1), with or without active-chain packing, or the library switches after its probe of the first iterations
    (the default: it does when lockstep leaves most lanes idle) -- in fewer rounds"""
    import os
    data, nsub = sleepstudy_data()
    cases = [("x ~ Normal(1, 2)\ny ~ Gamma(2, 1)", [("x", 0.5), ("y", 1.0)], {}, 700, 60, 5),
             (SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))], data, 96, 10, 0)]
    for model, init, dat, C, steps, burnin in cases:
        out = {}
        for name, env in (("lockstep", dict(SMC_NUTS_ASYNC="0")), ("async", dict(SMC_NUTS_ASYNC="1")), ("auto", {}),
                          ("async_packed", dict(SMC_NUTS_ASYNC="1", SMC_COMPACT="1"))):
            os.environ.update(env)
            try:
                low, dev = _build(model, init, dat, C)
                out[name] = dev.run("nuts", np.array(low.init), C, steps, burnin, seed=11)
                dev.close()
            finally:
                for k in env:
                    os.environ.pop(k)
        ref = out["lockstep"]
        for name in ("async", "auto", "async_packed"):
            if name == "async_packed" and model is SLEEPSTUDY2:
                continue      # (reductions over the data: a packed evaluation rounds differently, see the packing test)
            for k in ("draws", "epsilon", "jmax", "accept", "loglik", "mean", "m2", "accept_rate"):
                assert np.array_equal(ref[k], out[name][k]), (name, k)
            assert ref["n_leapfrogs"] == out[name]["n_leapfrogs"]
        assert out["async"]["n_evals"] < ref["n_evals"] and out["auto"]["n_evals"] <= ref["n_evals"]


def test_device_side_lag1_statistics_give_the_same_ess(smc):
    """calcStats! (src/samplers.jl:352-374) from the moments and lag-1 sums the device accumulates == the same formula on
    the stored draws; with store_draws=False the ESS is still there"""
    X, Y, _ = linear_data(500, 6)
    kw = dict(steps=260, burnin=60, isteps=3, stepsize=0.03, chains=24, seed=9, data=dict(X=X, Y=Y), beta=np.zeros(6))
    r = smc.simpleHMC(LINEAR, **kw)
    from simplemcmc_jl_b200.api import ess_from_draws
    assert relerr(r.misc["ess_by_chain"], ess_from_draws(r.misc["draws_scm"])) < 1e-8
    r2 = smc.simpleHMC(LINEAR, store_draws=False, **kw)
    assert not r2.params and np.allclose(r2.ess, r.ess, rtol=1e-8) and r2.ess[0] > 0
    r3 = smc.simpleNUTS("x ~ Normal(1, 2)", steps=200, burnin=50, chains=5, seed=2, x=0.5)
    assert relerr(r3.misc["ess_by_chain"], ess_from_draws(r3.misc["draws_scm"])) < 1e-8


def test_posterior_moments_linear(smc):
    """posterior mean / variance of the conjugate model within Monte-Carlo standard error"""
    X, Y, _ = linear_data()
    M = 10
    P = np.linalg.inv(X.T @ X + np.eye(M))
    mean, var = P @ X.T @ Y, np.diag(P)
    for kind, kw in (("hmc", dict(isteps=4, stepsize=0.03)), ("nuts", {}), ("rwm", {})):
        fn = dict(hmc=smc.simpleHMC, nuts=smc.simpleNUTS, rwm=smc.simpleRWM)[kind]
        steps, burnin = (3000, 1000) if kind != "rwm" else (6000, 3000)
        res = fn(LINEAR, steps=steps, burnin=burnin, chains=128, seed=3, data=dict(X=X, Y=Y), beta=np.zeros(M), **kw)
        d = res.params["beta"]                       # M x samples x chains
        ess = max(res.ess[0], 200.0)
        assert np.all(np.abs(d.mean(axis=(1, 2)) - mean) < 6 * np.sqrt(var / ess)), kind
        assert np.all(np.abs(d.var(axis=(1, 2)) / var - 1) < 0.15), kind
        assert np.all(res.misc["rhat"] < 1.1), kind


# test/test2.jl:61-85: 17 univariate targets; KS statistic sqrt(n) * max|ECDF - CDF| below the reference's threshold 5
TARGETS = [("Normal(1, 1)", stats.norm(1, 1), 0.0), ("Normal(3, 12)", stats.norm(3, 12), 3.0),
           ("Weibull(1, 1)", stats.weibull_min(1, scale=1), 1.0), ("Weibull(3, 1)", stats.weibull_min(3, scale=1), 1.0),
           ("Uniform(0, 2)", stats.uniform(0, 2), 1.0), ("TDist(2.2)", stats.t(2.2), 0.0), ("TDist(4)", stats.t(4), 0.0),
           ("Beta(1, 2)", stats.beta(1, 2), 0.5), ("Beta(3, 2)", stats.beta(3, 2), 0.5),
           ("Gamma(1, 2)", stats.gamma(1, scale=2), 1.0), ("Gamma(3, 0.2)", stats.gamma(3, scale=0.2), 1.0),
           ("Cauchy(0, 1)", stats.cauchy(0, 1), 0.0), ("Cauchy(-1, 0.2)", stats.cauchy(-1, 0.2), -1.0),
           ("Exponential(3)", stats.expon(scale=3), 1.0), ("Exponential(0.2)", stats.expon(scale=0.2), 1.0),
           ("logNormal(-1, 1)", stats.lognorm(1, scale=math.exp(-1)), 1.0), ("logNormal(2, 0.1)", stats.lognorm(0.1, scale=math.exp(2)), 7.0)]


def _ks(sample, dist):
    x = np.sort(sample)
    n = len(x)
    cdf = dist.cdf(x)
    return math.sqrt(n) * max(np.max(np.abs(np.arange(1, n + 1) / n - cdf)), np.max(np.abs(np.arange(0, n) / n - cdf)))


@pytest.mark.parametrize("sampler", ["rwm", "hmc", "nuts"])
def test_ks_univariate_targets(smc, sampler):
    """test/test2.jl:25-55.  The reference accepts when sqrt(n) * sup|ECDF - CDF| < 5 for its single chain of
    n = 9000 draws, i.e. sup-distance < 0.0527; here the ECDF is pooled over 64 chains and held to the same
    sup-distance (initial value = exact mean, HMC isteps=2, stepsize=std/5 as in test2.jl:31-44)."""
    n_ref = 9000
    for spec, dist, _ in TARGETS:
        model = "x ~ %s" % spec
        mean = float(dist.mean()) if np.isfinite(dist.mean()) else 1.0
        std = float(dist.std()) if np.isfinite(dist.std()) else 1.0
        if sampler == "hmc":
            res = smc.simpleHMC(model, steps=10000, burnin=1000, isteps=2, stepsize=std / 5, chains=64, seed=1, x=mean)
        elif sampler == "nuts":
            res = smc.simpleNUTS(model, steps=4000, burnin=1000, chains=64, seed=1, x=mean)
        else:
            res = smc.simpleRWM(model, steps=10000, burnin=1000, chains=64, seed=1, x=mean)
        d = res.params["x"]                  # samples x chains
        ks = _ks(d.reshape(-1), dist) / math.sqrt(d.size) * math.sqrt(n_ref)
        assert ks < 5.0, (sampler, spec, ks)
