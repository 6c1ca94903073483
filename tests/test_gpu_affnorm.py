"""SMC_OP_AFFNORM (csrc/affnorm.cu): the affine-residual-into-Normal statement of examples/ornstein.jl:24-26 and models of
that shape, device vs oracle at 1e-12 (logp relative; gradient relative to max |g|), every chain-count regime of the two
kernels, and the BASELINE size T = 1e6 (configs[4])."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.refmodel import generate_model_function  # noqa: E402
from oracle.samplers import ReplayRNG, simple_hmc  # noqa: E402
from util import ORNSTEIN, ou_data, relerr  # noqa: E402

TOL = 1e-12
INIT = dict(tau=20.0, mu=10.0, sigma=0.1)


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    return load_package()


def _check(lp, g, lp0, g0, tol=TOL):
    assert abs(lp - lp0) <= tol * max(abs(lp0), 1e-300), (lp, lp0, abs(lp - lp0) / abs(lp0))
    assert relerr(g, g0) <= tol, (g, g0, relerr(g, g0))


def _points(init, C, seed, spread=0.05):
    rng = np.random.default_rng(seed)
    return init[None, :] * (1 + spread * rng.standard_normal((C, init.size)))


@pytest.mark.parametrize("C", [1, 2, 3, 8, 20, 37, 64, 300])
def test_ornstein_every_chain_regime(smc, C):
    """C = 1, 2, 3 (4 lanes), 8: stream kernel with the chains in registers; 20: the same with chain groups;
    37, 64, 300: thread-per-chain kernel with 64 / 64 / 128 threads per block"""
    x = ou_data(5000)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, **INIT)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=C, data=dict(x=x), **INIT)
    assert f.lowered.n_affnorm == 1
    b = _points(init, C, C)
    lp, g = f(b)
    for c in range(0, C, max(1, C // 12)):
        _check(lp[c], g[c], *ref(b[c]))
    lp2, g2 = f(b)                                           # second call = the captured graph: same bits
    assert np.array_equal(lp, lp2) and np.array_equal(g, g2)
    # no-gradient closure
    f0, *_ = smc.generateModelFunction(ORNSTEIN, gradient=False, chains=C, data=dict(x=x), **INIT)
    assert np.max(np.abs(f0(b) - lp) / np.abs(lp)) <= 1e-15


def test_agrees_with_the_element_wise_tape(smc):
    """the register-program interpreter evaluates the same statements element by element: both paths against each other"""
    x = ou_data(4000)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=16, data=dict(x=x), **INIT)
    os.environ["SMC_NO_AFFNORM"] = "1"
    try:
        fv, *_ = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=16, data=dict(x=x), **INIT)
    finally:
        del os.environ["SMC_NO_AFFNORM"]
    assert f.lowered.n_affnorm == 1 and fv.lowered.n_affnorm == 0
    b = _points(init, 16, 3)
    lp, g = f(b)
    lpv, gv = fv(b)
    assert np.max(np.abs(lp - lpv) / np.abs(lp)) <= 1e-12
    assert max(relerr(g[c], gv[c]) for c in range(16)) <= 1e-11      # the interpreter's own bar (test_ornstein_cfg5)


MODELS = [
    # (model, init, which data) -- term orders, signs, a scalar first, mu a parameter, sigma a literal, three arrays
    ("r = a * s - b\nr ~ Normal(m, 2.0)", dict(s=1.5, m=0.5)),
    ("r = s - (a - b * s)\nr ~ Normal(0, sg)", dict(s=0.7, sg=1.3)),
    ("r = -(a + s) + c * t\nr ~ Normal(m, sg)", dict(s=0.2, t=-0.4, m=0.1, sg=0.9)),
    ("r = a[2:end] - a[1:end-1] * s - b[3:end] * t + u\nr ~ Normal(0, sg)", dict(s=0.5, t=0.25, u=0.3, sg=1.1)),
    ("s ~ Normal(0, 3)\nr = b * s + a\nr ~ Normal(s, 1.5)", dict(s=0.4)),             # the scalar is both a coefficient and the mean
]


@pytest.mark.parametrize("k", range(len(MODELS)))
@pytest.mark.parametrize("C", [1, 5, 40])
def test_other_models_of_the_same_shape(smc, k, C):
    model, init = MODELS[k]
    rng = np.random.default_rng(10 + k)
    n = 4099
    data = dict(a=rng.standard_normal(n), b=rng.standard_normal(n + 1)[: n] + 2.0, c=rng.standard_normal(n))
    if "[" in model:
        data = dict(a=rng.standard_normal(n), b=rng.standard_normal(n + 1))
    ref = generate_model_function(model, data=data, gradient=True, init=list(init.items()))
    f, npar, pars, b0 = smc.generateModelFunction(model, gradient=True, chains=C, data=data, **init)
    assert f.lowered.n_affnorm == 1
    pts = b0[None, :] * (1 + 0.1 * rng.standard_normal((C, b0.size)))
    lp, g = f(pts)
    for c in range(0, C, max(1, C // 5)):
        _check(lp[c], g[c], *ref(pts[c]))


def test_sigma_out_of_support_is_minus_inf_for_that_chain_only(smc):
    x = ou_data(3000)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, **INIT)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=4, data=dict(x=x), **INIT)
    b = np.tile(init, (4, 1))
    b[1, 2] = -0.1            # sigma < 0: Gamma prior and the Normal scale both out of support
    b[3, 0] = -5.0            # tau < 0
    lp, g = f(b)
    assert lp[1] == -np.inf and lp[3] == -np.inf and np.all(g[[1, 3]] == 0)
    _check(lp[0], g[0], *ref(b[0]))
    _check(lp[2], g[2], *ref(b[2]))


def test_hmc_trajectory_matches_the_oracle(smc):
    """first-K-step trajectory parity on the OU model through the new statement (injected normals / uniforms)"""
    from simplemcmc_jl_b200 import api
    x = ou_data(3000)
    data = dict(x=x)
    ref = generate_model_function(ORNSTEIN, data=data, gradient=True, **INIT)
    steps, isteps, eps, C = 12, 3, 2e-4, 2
    rng = np.random.default_rng(4)
    normals, unif = rng.standard_normal((steps, C, 3)), rng.random((steps, C, 1))
    low, dev = api._build(ORNSTEIN, list(INIT.items()), data, True, C, 0, "assign")
    assert low.n_affnorm == 1
    raw = dev.run("hmc", np.array(low.init), C, steps, 0, isteps=isteps, stepsize=eps, inject_normals=normals, inject_uniforms=unif)
    dev.close()
    for c in range(C):
        o = simple_hmc(ref, np.array(ref.init), ReplayRNG(normals[:, c], unif[:, c]), steps, 0, isteps, eps)
        assert np.array_equal(raw["accept"][:, c], o["accept"])
        assert relerr(raw["draws"][:, c], o["draws"]) < 1e-10
        assert relerr(raw["loglik"][:, c], o["loglik"]) < 1e-12


@pytest.mark.parametrize("C", [1, 3, 64])
def test_cfg5_at_baseline_size(smc, C):
    """BASELINE configs[4]: T = 1e6.  Three chains against the oracle (NumPy, sequential sums)."""
    x = ou_data(1_000_000)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, **INIT)
    f, n, pars, init = smc.generateModelFunction(ORNSTEIN, gradient=True, chains=C, data=dict(x=x), **INIT)
    assert f.lowered.n_affnorm == 1
    b = _points(init, C, 7, spread=0.02)
    lp, g = f(b)
    for c in sorted({0, C // 2, C - 1}):
        _check(lp[c], g[c], *ref(b[c]))
    res = smc.simpleHMC(ORNSTEIN, steps=30, burnin=10, isteps=2, stepsize=1e-5, chains=C, data=dict(x=x), **INIT)
    assert 0.0 < res.acceptRate <= 1.0 and np.all(np.isfinite(res.loglik))
