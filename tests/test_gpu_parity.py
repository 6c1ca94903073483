"""Parity of the CUDA path (through the C ABI) with the oracle: log-density and gradient to 1e-12 relative in FP64
on identical inputs (BASELINE.json north_star).  Gradients are compared relative to max|g| (components near a mode
cancel to ~0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.refmodel import generate_model_function  # noqa: E402
from rule_cases import all_cases  # noqa: E402
from util import (HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, logistic_data,  # noqa: E402
                  ou_data, relerr, sleepstudy_data)

TOL = 1e-12


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    return load_package()


def _check(lp, g, lp0, g0, tol=TOL):
    assert abs(lp - lp0) <= tol * max(abs(lp0), 1e-300), (lp, lp0)
    if np.max(np.abs(g0)) > 0:
        assert relerr(g, g0) <= tol, (g, g0)
    else:
        assert np.max(np.abs(g)) == 0


def test_rule_matrix(smc):
    """the op x shape-pattern matrix of test/test.jl:142-224, device vs oracle"""
    worst = 0.0
    for ex, data, x0 in all_cases():
        model = "sum(%s)" % ex
        ref = generate_model_function(model, data=data, gradient=True, x=x0)
        f, n, pars, init = smc.generateModelFunction(model, gradient=True, data=data, x=x0)
        lp, g = f(init)
        lp0, g0 = ref(init)
        _check(lp, g, lp0, g0)
        worst = max(worst, relerr(g, g0) if np.max(np.abs(g0)) > 0 else 0.0)
    assert worst <= TOL


def test_rule_matrix_many_chains(smc):
    """every 7th case of the matrix again with 37 chains in lockstep (chain lanes, element lanes and the per-chain
    scalar staging of the register-program VM all in play): every chain must match the oracle at its own point"""
    C = 37
    rng = np.random.default_rng(11)
    for k, (ex, data, x0) in enumerate(all_cases()):
        if k % 7:
            continue
        model = "sum(%s)" % ex
        ref = generate_model_function(model, data=data, gradient=True, x=x0)
        f, n, pars, init = smc.generateModelFunction(model, gradient=True, chains=C, data=data, x=x0)
        pts = init[None, :] * (1.0 + 1e-3 * rng.standard_normal((C, init.size)))   # stay inside the supports
        lp, g = f(pts)
        for c in range(0, C, 6):
            lp0, g0 = ref(pts[c])
            if np.isfinite(lp0):
                _check(lp[c], g[c], lp0, g0)
            else:
                assert lp[c] == lp0 or (np.isnan(lp0) and not np.isfinite(lp[c]))


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("C", [1, 7, 64, 200])
def test_linear_cfg1(smc, fuse, C):
    from simplemcmc_jl_b200 import api
    X, Y, _ = linear_data()
    data = dict(X=X, Y=Y)
    ref = generate_model_function(LINEAR, data=data, gradient=True, beta=np.zeros(10))
    low, dev = api._build(LINEAR, [("beta", np.zeros(10))], data, True, C, 0, "assign", fuse=fuse)
    assert low.n_fused == (1 if fuse else 0)
    betas = 0.3 * np.random.default_rng(C).standard_normal((C, 10))
    lp, g = dev.eval(betas)
    for c in range(C):
        _check(lp[c], g[c], *ref(betas[c]))
    lp2, _ = dev.eval(betas, gradient=False)
    # (since tape version 4 the gradient evaluation sums the densities inside its adjoint programs: same terms, another order)
    assert np.max(np.abs(lp - lp2) / np.abs(lp)) < 1e-13


@pytest.mark.parametrize("family", ["normal", "bernoulli"])
def test_narrow_glm_many_chains_wide_tile(smc, family):
    """2048 chains and more on a model of at most 16 columns run the 256-chain tile (four m-tiles per warp): a ragged
    chain count that leaves the last tile partly empty, rows not a multiple of the row tile"""
    from simplemcmc_jl_b200 import api
    C, N, M = 2500, 1003, 10
    if family == "normal":
        X, Y, _ = linear_data(N, M, 7)
        model, name = LINEAR, "beta"
    else:
        X, Y, _ = logistic_data(N, M)
        model, name = LOGISTIC, "vars"
    data = dict(X=X, Y=Y)
    ref = generate_model_function(model, data=data, gradient=True, **{name: np.zeros(M)})
    low, dev = api._build(model, [(name, np.zeros(M))], data, True, C, 0, "assign")
    betas = 0.3 * np.random.default_rng(5).standard_normal((C, M))
    lp, g = dev.eval(betas)
    for c in list(range(0, C, 97)) + [C - 1, 2047, 2048, 2303, 2304]:
        _check(lp[c], g[c], *ref(betas[c]))
    lp2, _ = dev.eval(betas, gradient=False)
    # (since tape version 4 the gradient evaluation sums the densities inside its adjoint programs: same terms, another order)
    assert np.max(np.abs(lp - lp2) / np.abs(lp)) < 1e-13
    dev.close()


@pytest.mark.parametrize("fuse", [True, False])
@pytest.mark.parametrize("N,M", [(1000, 10), (777, 3), (4099, 33), (300, 64)])
def test_logistic_cfg3_shapes(smc, fuse, N, M):
    """ragged sizes: rows not a multiple of the row tile, widths that need padding to 16/32/64"""
    from simplemcmc_jl_b200 import api
    X, Y, _ = logistic_data(N, M)
    data = dict(X=X, Y=Y)
    ref = generate_model_function(LOGISTIC, data=data, gradient=True, vars=np.zeros(M))
    C = 19
    low, dev = api._build(LOGISTIC, [("vars", np.zeros(M))], data, True, C, 0, "assign", fuse=fuse)
    betas = 0.2 * np.random.default_rng(N).standard_normal((C, M))
    lp, g = dev.eval(betas)
    for c in range(C):
        _check(lp[c], g[c], *ref(betas[c]))


@pytest.mark.parametrize("C", [1, 2, 3, 4])
@pytest.mark.parametrize("N,M", [(1000, 10), (777, 3), (4099, 16), (515, 29)])
def test_glm_few_chains_stream_path(smc, N, M, C):
    """1-4 chains take the HBM-stream kernel (column-major X, thread per row) for M <= 16 (one chain: M <= 32); both
    families, both link signs, plus the same model evaluated with 19 chains (DMMA kernel) must agree with each other"""
    from simplemcmc_jl_b200 import api
    X, Y, _ = logistic_data(N, M)
    Xl, Yl, _ = linear_data(N, M)
    rng = np.random.default_rng(N + C)
    betas = 0.2 * rng.standard_normal((C, M))
    neg = LOGISTIC.replace("exp(X * vars)", "exp(-(X * vars))")
    for model, data, par in ((LOGISTIC, dict(X=X, Y=Y), "vars"), (neg, dict(X=X, Y=Y), "vars"),
                             (LINEAR, dict(X=Xl, Y=Yl), "beta")):
        ref = generate_model_function(model, data=data, gradient=True, **{par: np.zeros(M)})
        low, dev = api._build(model, [(par, np.zeros(M))], data, True, C, 0, "assign")
        assert low.n_fused == 1
        lp, g = dev.eval(betas)
        for c in range(C):
            _check(lp[c], g[c], *ref(betas[c]))
        low2, dev2 = api._build(model, [(par, np.zeros(M))], data, True, 19, 0, "assign")
        lp19, g19 = dev2.eval(np.vstack([betas, np.zeros((19 - C, M))]))
        assert np.allclose(lp19[:C], lp, rtol=1e-13, atol=0) and relerr(g19[:C], g) <= 1e-12
        dev.close(); dev2.close()


@pytest.mark.parametrize("C", [1, 3, 19])
@pytest.mark.parametrize("N,M", [(1000, 10), (4099, 33)])
def test_binomial_count_glm(smc, N, M, C):
    """cfg3's count variant: Y ~ Binomial(20, 1/(1+exp(+-X*vars))) runs on the fused GLM kernels (stream kernel for few
    chains, DMMA kernel otherwise); lchoose(n, y) is summed once per model, not per chain"""
    from simplemcmc_jl_b200 import api
    rng = np.random.default_rng(N + C)
    X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
    b0 = 0.3 * rng.standard_normal(M)
    Y = rng.binomial(20, 1.0 / (1.0 + np.exp(X @ b0))).astype(np.float64)
    data = dict(X=X, Y=Y)
    betas = b0[None, :] + 0.1 * rng.standard_normal((C, M))
    for link in ("exp(X * vars)", "exp(-(X * vars))"):
        model = "vars ~ Normal(0, 1.0)\nprob = 1 / (1. + %s)\nY ~ Binomial(20, prob)" % link
        ref = generate_model_function(model, data=data, gradient=True, vars=np.zeros(M))
        low, dev = api._build(model, [("vars", np.zeros(M))], data, True, C, 0, "assign")
        assert low.n_fused == 1
        lp, g = dev.eval(betas)
        for c in range(C):
            _check(lp[c], g[c], *ref(betas[c]))
        dev.close()


def test_sufficient_statistics_variant(smc):
    """opt-in `sufficient_stats=True`: a Normal identity-link regression evaluated from X'X, X'y, y'y (2*M^2 flops per chain
    instead of 4*N*M).  Not the reference's arithmetic -- the tolerance is 1e-10 (cancellation in beta'X'X beta - 2 beta'X'y + y'y),
    the default path stays the 1e-12 one -- but the same log-density, gradient and HMC trajectory."""
    from simplemcmc_jl_b200 import api
    for N, M, C in ((1000, 10, 7), (4099, 33, 40), (300, 64, 3)):
        X, Y, _ = linear_data(N, M)
        data = dict(X=X, Y=Y)
        b = 0.2 * np.random.default_rng(N).standard_normal((C, M))
        for lik in ("resid = Y - X * beta\nresid ~ Normal(0, 2.0)", "Y ~ Normal(X * beta, 0.5)", "eta = X * beta\neta ~ Normal(Y, 1.5)"):
            model = "beta ~ Normal(0, 1.0)\n" + lik
            ref = generate_model_function(model, data=data, gradient=True, beta=np.zeros(M))
            low, dev = api._build(model, [("beta", np.zeros(M))], data, True, C, 0, "assign", sufficient_stats=True)
            assert low.n_fused == 1
            lp, g = dev.eval(b)
            for c in range(C):
                _check(lp[c], g[c], *ref(b[c]), tol=1e-10)
            dev.close()
    # a Normal density whose variate is affine in data arrays with per-chain scalar coefficients (the Ornstein-Uhlenbeck
    # residual x[t] - x[t-1]*fac - mu*(1-fac)): the whole likelihood becomes a quadratic form in a 3x3 Gram matrix
    x = ou_data(5000)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, tau=20.0, mu=10.0, sigma=0.1)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=9, data=dict(x=x), sufficient_stats=True,
                                                 tau=20.0, mu=10.0, sigma=0.1)
    assert not any(st.op == 32 and st.aux1 > 1 for st in f.lowered.fwd)          # no length-T density statement is left
    pts = init[None, :] * (1 + 0.05 * np.random.default_rng(2).standard_normal((9, 3)))
    lp, g = f(pts)
    for c in range(9):
        lp0, g0 = ref(pts[c])
        assert abs(lp[c] - lp0) <= 1e-9 * abs(lp0) and relerr(g[c], g0) <= 1e-7, (lp[c], lp0, g[c], g0)
    X, Y, _ = linear_data(2000, 10)
    kw = dict(steps=40, burnin=0, isteps=3, stepsize=0.02, chains=16, seed=3, data=dict(X=X, Y=Y), beta=np.zeros(10))
    r0 = smc.simpleHMC(LINEAR, **kw)
    r1 = smc.simpleHMC(LINEAR, sufficient_stats=True, **kw)
    assert np.array_equal(r0.accept, r1.accept)
    assert relerr(r1.params["beta"], r0.params["beta"]) <= 1e-9


def test_normal_glm_variants(smc):
    """the GLM matcher must keep signs straight for every way of writing the same likelihood"""
    X, Y, _ = linear_data(500, 6)
    data = dict(X=X, Y=Y)
    b = 0.2 * np.random.default_rng(3).standard_normal((5, 6))
    for lik in ("resid = Y - X * beta\nresid ~ Normal(0, 2.0)", "resid = X * beta - Y\nresid ~ Normal(0, 2.0)",
                "Y ~ Normal(X * beta, 0.5)", "eta = X * beta\neta ~ Normal(Y, 1.5)"):
        model = "beta ~ Normal(0, 1.0)\n" + lik
        ref = generate_model_function(model, data=data, gradient=True, beta=np.zeros(6))
        f, *_ = smc.generateModelFunction(model, gradient=True, chains=5, data=data, beta=np.zeros(6))
        assert f.lowered.n_fused == 1
        lp, g = f(b)
        for c in range(5):
            _check(lp[c], g[c], *ref(b[c]))


def test_ornstein_cfg5(smc):
    x = ou_data(1000)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, tau=20.0, mu=10.0, sigma=0.1)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=16, data=dict(x=x), tau=20.0, mu=10.0, sigma=0.1)
    assert [p.sym for p in pars] == ["tau", "mu", "sigma"] and n == 3
    rng = np.random.default_rng(0)
    b = init[None, :] * (1 + 0.05 * rng.standard_normal((16, 3)))
    lp, g = f(b)
    for c in range(16):
        _check(lp[c], g[c], *ref(b[c]), tol=1e-12)   # (the AFFNORM kernel; the register-program interpreter it replaced for this
                                                     # shape needed 1e-11: tests/test_gpu_affnorm.py::test_agrees_with_the_element_wise_tape)


def test_sleepstudy_cfg4(smc):
    data, nsub = sleepstudy_data()
    init = dict(intercept=250.0, slope=10.0, rand_int=np.zeros(nsub), rand_slope=np.zeros(nsub))
    ref = generate_model_function(SLEEPSTUDY2, data=data, gradient=True, init=list(init.items()))
    f, n, pars, b0 = smc.generateModelFunction(SLEEPSTUDY2, gradient=True, chains=8, data=data, **init)
    assert n == 38
    b = b0[None, :] + np.random.default_rng(1).standard_normal((8, 38))
    lp, g = f(b)
    for c in range(8):
        _check(lp[c], g[c], *ref(b[c]))


def test_hierarchical_cfg4_both_gather_adjoints(smc):
    """`assign` reproduces the reference's overwrite (parsing.jl:277-281), `add` is the corrected scatter-add"""
    data, (N, D, L) = hierarchical_data()
    init = dict(mu=np.zeros((1, L)), sigma=np.ones((1, L)), beta=np.zeros((D, L)))
    ref = generate_model_function(HIERARCHICAL, data=data, gradient=True, init=list(init.items()))
    rng = np.random.default_rng(2)
    b = ref.init + 0.05 * rng.standard_normal(len(ref.init))
    b[L:2 * L] = np.abs(b[L:2 * L]) + 0.5
    f, n, pars, _ = smc.generateModelFunction(HIERARCHICAL, gradient=True, data=data, ref_adjoint="assign", **init)
    assert n == 30 and [tuple(p.size) for p in pars] == [(1, L), (1, L), (D, L)]
    lp, g = f(b)
    _check(lp, g, *ref(b))
    f2, *_ = smc.generateModelFunction(HIERARCHICAL, gradient=True, data=data, ref_adjoint="add", **init)
    lp2, g2 = f2(b)
    gn = np.zeros_like(b)
    for i in range(len(b)):
        bp, bm = b.copy(), b.copy()
        bp[i] += 1e-6
        bm[i] -= 1e-6
        gn[i] = (ref(bp)[0] - ref(bm)[0]) / 2e-6
    assert lp2 == lp
    assert np.allclose(g2, gn, rtol=1e-5, atol=1e-6)


def test_out_of_support_chain_gets_minus_inf_and_zero_gradient(smc):
    """src/parsing.jl:465-467 per chain; the other chains are unaffected"""
    model = "x ~ Uniform(0, 1)\ny = x * 2\ny ~ Normal(0, 1)"
    f, *_ = smc.generateModelFunction(model, gradient=True, chains=4, x=0.5)
    lp, g = f(np.array([[0.25], [1.5], [0.75], [-0.1]]))
    assert lp[1] == -np.inf and lp[3] == -np.inf and np.all(g[[1, 3]] == 0)
    ref = generate_model_function(model, gradient=True, x=0.5)
    _check(lp[0], g[0], *ref(np.array([0.25])))
    _check(lp[2], g[2], *ref(np.array([0.75])))


def test_bernoulli_glm_certain_contradiction_gives_up(smc):
    X, Y, _ = logistic_data(256, 4)
    f, *_ = smc.generateModelFunction(LOGISTIC, gradient=True, chains=2, data=dict(X=X, Y=Y), vars=np.zeros(4))
    lp, g = f(np.array([[0.1, 0.0, 0.0, 0.0], [800.0, 0.0, 0.0, 0.0]]))   # exp overflows: prob == 0 with some Y == 1
    assert np.isfinite(lp[0]) and lp[1] == -np.inf and np.all(g[1] == 0)


def test_error_behaviour(smc):
    from simplemcmc_jl_b200 import ModelError
    for ex in ("logpdfBernoulli(1, x)", "logpdfPoisson(1, x)", "logpdfBinomial(3, 0.5, x)", "logpdfBinomial(x, 0.5, 2)"):
        with pytest.raises(ModelError):        # test/test.jl:199-208
            smc.generateModelFunction("sum(%s)" % ex, gradient=True, x=np.array([0.0]))
    with pytest.raises(ModelError):
        smc.generateModelFunction("x ~ Uniform(0, 1)", gradient=True, x=1.5)   # initial values out of support
    with pytest.raises(ModelError):
        smc.generateModelFunction("y = 2 * x\nz ~ Normal(0, 1)", gradient=True, data=dict(z=np.ones(3)), x=1.0)
    with pytest.raises(AssertionError):
        smc.simpleHMC("x ~ Normal(0,1)", steps=10, burnin=20, x=0.0)          # checkSteps, src/samplers.jl:324-327


def test_row_sharding_refuses_tapes_it_cannot_sum_across_ranks(smc):
    """shard="rows": only fused GLM statements sum over the rows across the ranks.  A tape that reduces the sharded data in any
    other statement (Poisson regression, a Normal likelihood whose sigma is a parameter, an unfused lowering) would give every
    rank a different logp / gradient without an error: smc_model_finalize refuses it.  The supported families still build."""
    from simplemcmc_jl_b200 import api
    from simplemcmc_jl_b200.capi import SMCError
    X, Y, _ = logistic_data(512, 4)
    low, dev = api._build(LOGISTIC, [("vars", np.zeros(4))], dict(X=X, Y=Y), True, 2, 0, "assign", shard="rows")
    lp, g = dev.eval(np.zeros((2, 4)))          # (one rank, no communicator: the shard is the whole data)
    assert np.all(np.isfinite(lp))
    dev.close()
    counts = np.random.default_rng(0).poisson(2.0, 512).astype(np.float64)
    refused = [("vars ~ Normal(0, 1.0)\nlam = exp(X * vars)\nY ~ Poisson(lam)", [("vars", np.zeros(4))], dict(X=X, Y=counts), {}),
               ("vars ~ Normal(0, 1.0)\ns ~ Gamma(2, 2)\nresid = Y - X * vars\nresid ~ Normal(0, s)", [("vars", np.zeros(4)), ("s", 1.0)],
                dict(X=X, Y=Y), {}),
               (LOGISTIC, [("vars", np.zeros(4))], dict(X=X, Y=Y), dict(fuse=False))]
    for model, init, data, kw in refused:
        with pytest.raises(SMCError, match="row-sharded"):
            api._build(model, init, data, True, 2, 0, "assign", shard="rows", **kw)


def test_no_gradient_closure_and_kwarg_order(smc):
    f, n, pars, init = smc.generateModelFunction("a ~ Normal(0,1)\nb ~ Normal(a, 2.0)", gradient=False, b=np.array([1.0, 2.0]), a=0.5)
    assert [p.sym for p in pars] == ["b", "a"] and list(init) == [1.0, 2.0, 0.5]    # kwarg order, parsing.jl:303-334
    ref = generate_model_function("a ~ Normal(0,1)\nb ~ Normal(a, 2.0)", gradient=False, init=[("b", np.array([1.0, 2.0])), ("a", 0.5)])
    assert abs(f(init) - ref(init)) < 1e-13
