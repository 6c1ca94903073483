"""Random multi-statement models (tests/fuzz_models.py): device log-density and gradient against the oracle.

rule_cases.py differentiates one call at a time, as test/test.jl:142-224 does; here whole graphs go through the lowering,
the statement fusion, the register-program planner (copy propagation, alias materialisation, dead-write elimination) and
both executors, with the reference's quirks (adjoint of a slice / gather is an assignment, ties of max/min) on both
sides.  Bar: 1e-12 relative on logp, 1e-12 of max|g| on the gradient -- the same as test_gpu_parity.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from fuzz_models import make_data, make_params, random_model  # noqa: E402
from oracle.refmodel import ModelError, generate_model_function  # noqa: E402
from util import relerr  # noqa: E402

TOL = 1e-12
SCALE = int(os.environ.get("SMC_FUZZ_SCALE", "1"))     # SMC_FUZZ_SCALE=k: k times as many seeds (occasional deep runs)


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    return load_package()


def _check(lp, g, lp0, g0, what):
    assert abs(lp - lp0) <= TOL * max(abs(lp0), 1e-300), (what, lp, lp0)
    if np.max(np.abs(g0)) > 0:
        assert relerr(g, g0) <= TOL, (what, g, g0)
    else:
        assert np.max(np.abs(g)) == 0, what


def _run(seed, n, m, C, fuse):
    from simplemcmc_jl_b200 import api
    rng = np.random.default_rng(seed)
    data, pars = make_data(rng, n, m), make_params(rng, n, m)
    model = random_model(seed)
    try:
        ref = generate_model_function(model, data=data, gradient=True, init=pars)
    except ModelError as e:
        # sums over long vectors can push a later density out of its support at the initial point: the reference refuses
        # the model (src/parsing.jl:379-382); the device must see the same point as outside the support
        assert "out of model support" in str(e), (seed, model)
        low, dev = api._build(model, pars, data, True, C, 0, "assign", fuse=fuse)
        try:
            lp, g = dev.eval(np.array(low.init)[None, :])
            assert lp[0] == -np.inf and np.all(g == 0), (seed, model)     # src/parsing.jl:465-467
        finally:
            dev.close()
        return False
    low, dev = api._build(model, pars, data, True, C, 0, "assign", fuse=fuse)
    try:
        x0 = np.array(low.init)
        pts = x0[None, :] * (1.0 + 1e-3 * rng.standard_normal((C, x0.size)))
        pts[0] = x0
        lp, g = dev.eval(pts)
        for c in sorted(set([0, C - 1] + list(range(0, C, 9)))):
            lp0, g0 = ref(pts[c])
            assert np.isfinite(lp0)
            _check(lp[c], g[c], lp0, g0, (seed, c, model))
        lp2, _ = dev.eval(pts, gradient=False)
        # (the gradient evaluation sums the densities inside its adjoint programs: same terms, another order of summation)
        with np.errstate(invalid="ignore"):
            assert np.max(np.where(lp == lp2, 0.0, np.abs(lp - lp2) / np.maximum(np.abs(lp), 1e-300))) < 1e-13, (seed, model)
    finally:
        dev.close()
    return True


@pytest.mark.parametrize("fuse", [True, False])
def test_random_models_one_chain(smc, fuse):
    for seed in list(range(150)) + list(range(1000, 1000 + 150 * (SCALE - 1))):
        _run(seed, 23, 5, 1, fuse)


def test_random_models_many_chains(smc):
    for seed in list(range(150, 250)) + list(range(3000, 3000 + 100 * (SCALE - 1))):
        _run(seed, 23, 5, 37, True)


def test_random_models_long_vectors(smc):
    """n = 1501: several element chunks per block and partial reductions across blocks in the register-program kernels"""
    assert sum(_run(seed, 1501, 7, 5, True) for seed in range(250, 290)) >= 25


def test_random_models_hmc_trajectories(smc):
    """simpleHMC (src/samplers.jl:146-167) on random graphs with injected randomness: same accept decisions, same draws.
    The scale parameter `s` can leave its support along a trajectory (Gamma prior, Normal scale): both sides must stop
    the trajectory and reject (src/samplers.jl:156)."""
    from oracle.samplers import ReplayRNG, simple_hmc
    from simplemcmc_jl_b200 import api
    steps, C = 12, 3
    for seed in range(300, 320):
        rng = np.random.default_rng(seed)
        data, pars = make_data(rng), make_params(rng)
        model = random_model(seed)
        ref = generate_model_function(model, data=data, gradient=True, init=pars)
        low, dev = api._build(model, pars, data, True, C, 0, "assign")
        try:
            x0 = np.array(low.init)
            normals, unif = rng.standard_normal((steps, C, x0.size)), rng.random((steps, C, 1))
            raw = dev.run("hmc", x0, C, steps, 0, isteps=3, stepsize=0.02, inject_normals=normals, inject_uniforms=unif)
            for c in range(C):
                o = simple_hmc(ref, x0, ReplayRNG(normals[:, c], unif[:, c]), steps, 0, 3, 0.02)
                assert np.array_equal(raw["accept"][:, c], o["accept"]), (seed, model)
                assert relerr(raw["draws"][:, c], o["draws"]) < 1e-10, (seed, model)
        finally:
            dev.close()
