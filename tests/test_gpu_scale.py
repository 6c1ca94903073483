"""Full-size (BASELINE.json configs[1]) checks through size-independent properties: closed form of the conjugate
linear model on sampled chains, linearity of the gradient in beta, chain-permutation invariance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from util import LINEAR, relerr  # noqa: E402


@pytest.fixture(scope="module")
def big():
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200 import api
    rng = np.random.default_rng(2)
    N, M, C = 1_000_000, 64, 4096
    X = np.empty((M, N)).T
    X[:, 0] = 1.0
    for j in range(1, M):
        X[:, j] = rng.standard_normal(N)
    Y = X @ rng.standard_normal(M) + rng.standard_normal(N)
    low, dev = api._build(LINEAR, [("beta", np.zeros(M))], dict(X=X, Y=Y), True, C, 0, "assign")
    return X, Y, dev, rng


def test_cfg2_closed_form_and_linearity(big):
    X, Y, dev, rng = big
    N, M = X.shape
    C = 4096
    b = 0.01 * rng.standard_normal((C, M))
    lp, g = dev.eval(b)
    for c in (0, 1, 777, 4095):
        r = Y - X @ b[c]
        lp0 = -0.5 * b[c] @ b[c] - 0.5 * r @ r - (N + M) * 0.9189385332046727
        g0 = -b[c] + X.T @ r
        assert abs(lp[c] - lp0) <= 1e-12 * abs(lp0)
        assert relerr(g[c], g0) <= 1e-12
    # the gradient is affine in beta: g(a) + g(b) - g(0) == g(a + b)
    b2 = b.copy()
    b2[0] = 0.0
    b2[1] = b[2] + b[3]
    _, g2 = dev.eval(b2)
    assert relerr(g[2] + g[3] - g2[0], g2[1]) <= 1e-11
    # chains are independent: permuting them permutes the result bit for bit
    perm = rng.permutation(C)
    lp3, g3 = dev.eval(b[perm])
    assert np.array_equal(lp3, lp[perm]) and np.array_equal(g3, g[perm])


def test_cfg2_hmc_runs_and_accepts(big):
    X, Y, dev, rng = big
    M = X.shape[1]
    mode = np.linalg.solve(X.T @ X + np.eye(M), X.T @ Y)
    init = mode[None, :] + 1e-3 * rng.standard_normal((4096, M))
    raw = dev.run("hmc", init, 4096, 6, 0, isteps=2, stepsize=5e-4, seed=9, store_draws=False, want=())
    assert raw["n_leapfrogs"] == 4096 * 6 * 2
    assert 0.6 < raw["accept_rate"].mean() < 0.95
    assert np.all(np.abs(raw["mean"] - mode) < 0.02)


def test_cfg2_posterior_moments_within_monte_carlo_error(big):
    """north_star's second correctness leg at full size: chains started from exact posterior draws of the conjugate model
    (Normal(mode, (X'X + I)^-1)) stay distributed that way under simpleHMC -- pooled means and variances of 4096 chains x 60
    steps against the closed form, in Monte-Carlo standard errors over the chains (64 components: |z| of a few is agreement;
    tools/posterior_check_cfg2.py, profiles/posterior_cfg2_r02.json: rms z 1.02 / 1.04 at 150 steps)"""
    X, Y, dev, rng = big
    M, C, steps = X.shape[1], 4096, 60
    cov = np.linalg.inv(X.T @ X + np.eye(M))
    mode = cov @ (X.T @ Y)
    init = mode[None, :] + np.random.default_rng(42).standard_normal((C, M)) @ np.linalg.cholesky(cov).T
    raw = dev.run("hmc", init, C, steps, 0, isteps=2, stepsize=5e-4, seed=17, store_draws=False, want=())
    mean_c, m2_c = raw["mean"], raw["m2"]
    pooled = mean_c.mean(0)
    z_mean = (pooled - mode) / (mean_c.std(0, ddof=1) / np.sqrt(C))
    per_chain_var = m2_c / steps + (mean_c - pooled) ** 2
    z_var = (per_chain_var.mean(0) - np.diag(cov)) / (per_chain_var.std(0, ddof=1) / np.sqrt(C))
    assert np.max(np.abs(z_mean)) < 4.5 and np.sqrt(np.mean(z_mean ** 2)) < 1.5, z_mean
    assert np.max(np.abs(z_var)) < 4.5 and np.sqrt(np.mean(z_var ** 2)) < 1.5, z_var
    assert np.max(np.abs(per_chain_var.mean(0) / np.diag(cov) - 1.0)) < 0.03
