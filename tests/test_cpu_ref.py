"""The C++ CPU baseline (oracle/cpu_ref.cpp) against the NumPy oracle: it must compute the same closure before its
timings mean anything."""
import numpy as np

from oracle import cpu_ref
from oracle.refmodel import generate_model_function
from util import LINEAR, LOGISTIC, linear_data, logistic_data, relerr


def test_linear_and_logistic_closures_match_the_oracle():
    rng = np.random.default_rng(0)
    X, Y, _ = linear_data()
    ref = generate_model_function(LINEAR, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(10))
    Xb, Yb, _ = logistic_data()
    ref2 = generate_model_function(LOGISTIC, data=dict(X=Xb, Y=Yb), gradient=True, vars=np.zeros(10))
    for _ in range(5):
        b = 0.3 * rng.standard_normal(10)
        lp, g = cpu_ref.evaluate("linear", X, Y, b)
        lp0, g0 = ref(b)
        assert abs(lp - lp0) <= 1e-13 * abs(lp0) and relerr(g, g0) < 1e-13
        lp, g = cpu_ref.evaluate("logistic", Xb, Yb, b)
        lp0, g0 = ref2(b)
        assert abs(lp - lp0) <= 1e-13 * abs(lp0) and relerr(g, g0) < 1e-13


def test_hmc_port_samples_the_conjugate_posterior():
    X, Y, _ = linear_data()
    P = np.linalg.inv(X.T @ X + np.eye(10))
    mean = P @ X.T @ Y
    r = cpu_ref.hmc("linear", X, Y, np.tile(mean, (4, 1)), 1500, 4, 0.03, seed=3, threads=4)
    assert r["n_leapfrogs"] == 4 * 1500 * 4
    assert np.all(r["accept_rate"] > 0.5)
    assert np.all(np.abs(r["final"] - mean) < 8 * np.sqrt(np.diag(P)))
