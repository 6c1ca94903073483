"""simpleAGD / simpleNM on the device closure against the oracle's literal restatement of src/solvers.jl: identical
iterate sequence (1e-12 relative), identical steps / converged / maximum."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.refmodel import generate_model_function  # noqa: E402
from oracle.solvers import simple_agd, simple_nm  # noqa: E402
from util import LINEAR, ORNSTEIN, linear_data, ou_data, relerr  # noqa: E402


@pytest.fixture(scope="module")
def smc():
    from smc_pkg import load_package
    return load_package()


def _agd_case(smc, model, data, init, maxiter, precision):
    ref = generate_model_function(model, data=data, gradient=True, init=list(init.items()))
    steps, fmax, conv, z, trace = simple_agd(ref, np.array(ref.init), maxiter=maxiter, precision=precision)
    res = smc.simpleAGD(model, maxiter=maxiter, precision=precision, data=data, **init)
    assert res.steps == steps and res.converged == conv
    assert abs(res.maximum - fmax) <= 1e-12 * abs(fmax)
    its = res.misc["iterates"]
    assert len(its) == len(trace)
    for k, (a, t) in enumerate(zip(its, trace)):
        assert relerr(a, t["z1"]) <= 1e-12, (k, a, t["z1"])
    return res, z


def test_agd_linear_iterates(smc):
    X, Y, _ = linear_data(1000, 10)
    res, z = _agd_case(smc, LINEAR, dict(X=X, Y=Y), dict(beta=np.zeros(10)), 100, 1e-5)
    assert relerr(res.params["beta"], z) <= 1e-12
    exact = np.linalg.solve(X.T @ X + np.eye(10), X.T @ Y)
    assert res.converged and np.max(np.abs(res.params["beta"] - exact)) < 1e-2


def test_agd_ornstein_iterates(smc):
    """three scalar parameters with positivity constraints: the Lipschitz search and the support back-off both run;
    a short horizon keeps the comparison before round-off has been amplified by the 2x / 0.9x branch on L"""
    x = ou_data(1000)
    init = dict(tau=20.0, mu=10.0, sigma=0.1)
    res, z = _agd_case(smc, ORNSTEIN, dict(x=x), init, 25, 1e-5)
    assert set(res.params) == {"tau", "mu", "sigma"} and isinstance(res.params["tau"], float)
    assert res.misc["n_launches"] <= 2 + 10 + 2 * res.steps + 10   # two launches per iteration (+ searches)


def test_agd_maxiter_zero_and_errors(smc):
    from simplemcmc_jl_b200 import ModelError
    X, Y, _ = linear_data(64, 3)
    res = smc.simpleAGD(LINEAR, maxiter=0, data=dict(X=X, Y=Y), beta=np.zeros(3))
    assert res.steps == 0 and res.converged is False       # i < maxiter is 0 < 0 (solvers.jl:101)
    with pytest.raises(ModelError):
        smc.simpleAGD("x ~ Uniform(0, 1)", x=2.0)


def test_agd_finds_data_in_caller_globals(smc):
    """external variables come from the caller's module when `data` is not given (parsing.jl:484-485)"""
    global X, Y
    X, Y, _ = linear_data(128, 3)
    try:
        res = smc.simpleAGD(LINEAR, maxiter=5, beta=np.zeros(3))
        assert res.steps == 5
    finally:
        del X, Y


def test_nm_matches_oracle(smc):
    X, Y, _ = linear_data(300, 3)
    data = dict(X=X, Y=Y)
    ref = generate_model_function(LINEAR, data=data, gradient=False, beta=np.zeros(3))
    it, ylo, conv, best, trace = simple_nm(ref, np.zeros(3), maxiter=150, precision=1e-3)
    res = smc.simpleNM(LINEAR, maxiter=150, precision=1e-3, data=data, beta=np.zeros(3))
    assert res.steps == it and res.converged == conv
    assert res.misc["moves"] == [t["move"] for t in trace]
    assert abs(res.maximum - ylo) <= 1e-12 * abs(ylo)
    assert relerr(res.params["beta"], best) <= 1e-12
    assert res.misc["n_launches"] <= 2 + res.steps + res.misc["moves"].count("shrink")   # one launch per iteration


def test_nm_ornstein_and_errors(smc):
    x = ou_data(500)
    init = dict(tau=20.0, mu=10.0, sigma=0.1)
    ref = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=False, init=list(init.items()))
    it, ylo, conv, best, trace = simple_nm(ref, np.array(ref.init), maxiter=40, precision=1e-3)
    res = smc.simpleNM(ORNSTEIN, maxiter=40, precision=1e-3, data=dict(x=x), **init)
    assert res.steps == it and res.misc["moves"] == [t["move"] for t in trace]
    assert abs(res.maximum - ylo) <= 1e-11 * abs(ylo)
    with pytest.raises(AssertionError):
        smc.simpleNM(LINEAR, precision=0.0, data=dict(X=np.ones((4, 1)), Y=np.ones(4)), beta=np.zeros(1))
