"""The oracle's literal restatement of src/solvers.jl (simpleAGD :35-107, simpleNM :122-212) against properties that
follow from the reference's statements themselves: on a concave quadratic the AGD iterates are a known closed form for
the first step, the Lipschitz estimate equals the curvature, and both solvers reach the analytic maximum."""
import math

import numpy as np
import pytest

from oracle.solvers import simple_agd, simple_nm
from oracle.refmodel import generate_model_function
from util import LINEAR, linear_data


def _quad(A, b):
    def ll(z):
        return float(-0.5 * z @ A @ z + b @ z), -(A @ z) + b
    return ll


def test_agd_first_iteration_closed_form():
    """isotropic quadratic, curvature a: L0 = max(1, a); theta0 = 1 => y0 = z0, zb1 = z1 = z0 + grad/L0 (:63-66)"""
    a = 4.0
    ll = _quad(a * np.eye(3), np.array([1.0, -2.0, 0.5]))
    z0 = np.array([0.3, 0.1, -0.2])
    steps, fmax, conv, z, trace = simple_agd(ll, z0, maxiter=1)
    g0 = ll(z0)[1]
    assert steps == 1 and conv is False
    t = trace[0]
    assert np.array_equal(t["y0"], z0)
    assert np.allclose(t["z1"], z0 + g0 / a, rtol=0, atol=1e-15)
    assert np.allclose(t["z1"], np.array([1.0, -2.0, 0.5]) / a, atol=1e-15)    # one exact Newton step
    # y0 - z1 = -g/L, grady - gradz = g  =>  2|d.dg|/d.d = 2 a > L0  =>  L1 = 2 L0 (:84)
    assert t["L0"] == a and t["L1"] == 2 * a
    assert t["theta1"] == 2 / (1 + math.sqrt(1 + 4 * 2.0))


def test_agd_lipschitz_search_leaves_support():
    """ll = log(x) - x has support x > 0: from z0 = 0.05, grad0 = 19 and z0 + grad0/L0 is inside the support already,
    but a -Inf region on the other side forces doublings: ll defined on (0, 1.5)"""
    def ll(z):
        x = z[0]
        if not 0 < x < 1.5:
            return -math.inf, np.zeros(1)
        return math.log(x) - x, np.array([1 / x - 1])
    steps, fmax, conv, z, trace = simple_agd(ll, np.array([0.05]), maxiter=200, precision=1e-10)
    assert conv and abs(z[0] - 1.0) < 1e-4 and abs(fmax + 1.0) < 1e-8
    assert trace[0]["L0"] >= 16.0          # 0.05 + 19/L < 1.5 needs L >= 16


def test_agd_reaches_regression_maximum():
    X, Y, _ = linear_data(200, 4)
    ref = generate_model_function(LINEAR, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(4))
    steps, fmax, conv, z, trace = simple_agd(ref, np.zeros(4), maxiter=500, precision=1e-9)
    exact = np.linalg.solve(X.T @ X + np.eye(4), X.T @ Y)
    assert conv and np.max(np.abs(z - exact)) < 1e-4
    assert any(t["restart"] for t in trace) or steps < 500
    assert abs(fmax - ref(exact)[0]) < 1e-6


def test_agd_errors():
    with pytest.raises(AssertionError):
        simple_agd(lambda z: (-math.inf, np.zeros(1)), np.zeros(1))


def test_nm_quadratic_and_moves():
    A = np.diag([1.0, 3.0])
    b = np.array([1.0, -1.5])
    f = lambda z: _quad(A, b)(np.asarray(z))[0]
    it, ylo, conv, best, trace = simple_nm(f, np.array([2.0, 2.0]), maxiter=400, precision=1e-6)
    assert conv and np.max(np.abs(best - b / np.diag(A))) < 1e-5
    assert abs(-ylo - f(b / np.diag(A))) < 1e-10
    kinds = {t["move"] for t in trace}
    assert {"extend", "reflect", "contract_hi", "contract_refl"} <= kinds
    # ylo never increases (a minimiser of -f)
    ys = [t["ylo"] for t in trace]
    assert all(b2 <= a2 + 0 for a2, b2 in zip(ys, ys[1:]))


def test_nm_first_step_by_hand():
    """f = -(x^2 + y^2) at init (1, 1): simplex columns (2,1), (1,2), (1,1); y = (5, 5, 2); ihi = 0 (first maximum);
    pbar = ((1,2)+(1,1))/2 = (1, 1.5); pstar = 2 pbar - (2,1) = (0, 2), ystar = 4; ystar >= ylo = 2 and l = #{ystar < y}
    = 2 > 1  =>  plain reflection (:173-175)"""
    f = lambda z: -(z[0] ** 2 + z[1] ** 2)
    it, ylo, conv, best, trace = simple_nm(f, np.array([1.0, 1.0]), maxiter=1)
    assert it == 1 and conv is False and trace[0]["move"] == "reflect"
    assert ylo == 2.0 and np.array_equal(best, [1.0, 1.0])


def test_nm_errors():
    with pytest.raises(AssertionError):
        simple_nm(lambda z: 0.0, np.zeros(1), precision=0.0)
    with pytest.raises(AssertionError):
        simple_nm(lambda z: math.inf, np.zeros(1))
