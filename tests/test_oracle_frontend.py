"""The oracle against what the reference pins for this path: the statement lists of SURVEY.md Appendix A (hand-derived
from the reference), the gradient-vs-finite-difference matrix of test/test.jl, and its error cases."""
import os

import numpy as np
import pytest

from oracle.refmodel import ModelError, generate_model_function
from rule_cases import all_cases
from util import (HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, logistic_data,
                  normalise_statements, ou_data, read_golden, sleepstudy_data)

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_linear_statement_list_matches_appendix_a1():
    X, Y, _ = linear_data()
    f = generate_model_function(LINEAR, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(10))
    fw, bw = f.statements()
    got = normalise_statements(fw + bw)
    gold = read_golden(os.path.join(GOLD, "linear_statements.txt"))
    assert got[:len(fw)] == gold["forward"]
    assert got[len(fw):] == gold["backward"]


def test_binomial_forward_statement_list_matches_appendix_a2():
    X, Y, _ = logistic_data()
    f = generate_model_function(LOGISTIC, data=dict(X=X, Y=Y), gradient=True, vars=np.zeros(10))
    fw, _ = f.statements()
    assert normalise_statements(fw) == read_golden(os.path.join(GOLD, "binomial_statements.txt"))["forward"]


def test_linear_closed_form():
    X, Y, _ = linear_data()
    f = generate_model_function(LINEAR, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(10))
    b = 0.1 * np.random.default_rng(0).standard_normal(10)
    lp, g = f(b)
    phi = lambda r: -r * r / 2 - np.log(np.sqrt(2 * np.pi))
    assert abs(lp - (phi(b).sum() + phi(Y - X @ b).sum())) < 1e-12 * abs(lp)
    assert np.allclose(g, -b + X.T @ (Y - X @ b), rtol=1e-12, atol=1e-9)


def test_logistic_closed_form():
    X, Y, _ = logistic_data()
    f = generate_model_function(LOGISTIC, data=dict(X=X, Y=Y), gradient=True, vars=np.zeros(10))
    b = 0.1 * np.random.default_rng(0).standard_normal(10)
    lp, g = f(b)
    p = 1 / (1 + np.exp(X @ b))
    assert np.allclose(g, -b + X.T @ (p - Y), rtol=1e-10, atol=1e-9)   # SURVEY.md Appendix A.2 closed form


CASES = all_cases()


def _fd(f, b0, h=1e-6):
    g = np.zeros_like(b0)
    for i in range(len(b0)):
        bp, bm = b0.copy(), b0.copy()
        bp[i] += h
        bm[i] -= h
        g[i] = (f(bp)[0] - f(bm)[0]) / (2 * h)
    return g


@pytest.mark.parametrize("chunk", range(8))
def test_rule_matrix_against_finite_differences(chunk):
    """test/test.jl:142-224: every rule x shape pattern; reference tolerance 2e-2 with forward differences
    (test.jl:17-21), here central differences and 1e-4"""
    for ex, data, x0 in CASES[chunk::8]:
        f = generate_model_function("sum(%s)" % ex, data=data, gradient=True, x=x0)
        b0 = np.atleast_1d(np.asarray(x0, dtype=np.float64)).flatten(order="F")
        _, g = f(b0)
        gn = _fd(f, b0)
        err = np.abs(g - gn) / np.maximum(2e-2, np.abs(gn))
        assert np.all(err < 1e-4), (ex, g, gn)


def test_discrete_arguments_are_not_derivable():
    """test/test.jl:199-208"""
    for ex in ("logpdfBernoulli(1, x)", "logpdfPoisson(1, x)", "logpdfBinomial(3, 0.5, x)", "logpdfBinomial(x, 0.5, 2)"):
        with pytest.raises(ModelError):
            generate_model_function("sum(%s)" % ex, gradient=True, x=np.array([0.0]))


def test_parameter_must_influence_outcome():
    with pytest.raises(ModelError):
        generate_model_function("y = 2 * x\nz ~ Normal(0, 1)", data=dict(z=np.ones(3)), gradient=True, x=1.0)


def test_out_of_support_gives_minus_inf_and_zero_gradient():
    """src/parsing.jl:465-467"""
    f = generate_model_function("x ~ Uniform(0, 1)", gradient=True, x=0.5)
    lp, g = f(np.array([1.5]))
    assert lp == -np.inf and np.all(g == 0)


def test_initial_value_out_of_support_is_an_error():
    with pytest.raises(ModelError):
        generate_model_function("x ~ Uniform(0, 1)", gradient=True, x=1.5)


def test_example_models_evaluate_and_match_finite_differences():
    x = ou_data(300)
    f = generate_model_function(ORNSTEIN, data=dict(x=x), gradient=True, tau=20.0, mu=10.0, sigma=0.1)
    b = np.array([20.0, 10.0, 0.1])
    assert np.allclose(f(b)[1], _fd(f, b, 1e-6), rtol=1e-5)
    data, nsub = sleepstudy_data()
    f = generate_model_function(SLEEPSTUDY2, data=data, gradient=True,
                                init=[("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))])
    b = f.init + 0.01
    assert np.allclose(f(b)[1], _fd(f, b, 1e-4), rtol=1e-5, atol=1e-6)


def test_hierarchical_gather_adjoint_is_the_reference_overwrite():
    """`dbeta[ll,:] = dt` overwrites, so with repeated group indices the reference gradient
    differs from the true one; the oracle reproduces the reference."""
    data, (N, D, L) = hierarchical_data()
    init = [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))]
    f = generate_model_function(HIERARCHICAL, data=data, gradient=True, init=init)
    b = f.init + 0.05 * np.random.default_rng(1).standard_normal(len(f.init))
    b[L:2 * L] = np.abs(b[L:2 * L]) + 0.5
    lp, g = f(b)
    gn = _fd(f, b)
    assert np.isfinite(lp)
    assert np.allclose(g[:2 * L], gn[:2 * L], rtol=1e-4, atol=1e-6)      # mu, sigma adjoints are right
    assert not np.allclose(g[2 * L:], gn[2 * L:], rtol=1e-3, atol=1e-6)  # beta adjoint is the known-wrong overwrite
