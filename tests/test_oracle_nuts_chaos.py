"""How tight can NUTS trajectory parity be held over a whole run?  The reference's NUTS (src/samplers.jl:180-315) starts at
epsilon = 1 and adapts by dual averaging, overshooting the leapfrog stability limit again and again; rounding differences
are amplified along the way.  Here the ORACLE runs against a copy of itself whose gradient is perturbed by one unit in the
last place at every evaluation: the amplification over 30 iterations is about three orders of magnitude (1e-16 -> 4e-13).
The device's elementary functions differ from glibc's by an ulp or two in the same way (and its reductions sum in another
order), so tests/test_gpu_samplers.py::test_nuts_trajectory_parity holds the first 12 iterations at 1e-10 and identical
tree depths and accept decisions over the whole run; its whole-run bar on the values (1e-4) is a loose safety net for the
second, scalar Normal x Gamma case, whose growth is larger than the linear model's measured here -- it is NOT a measured
device discrepancy, and it has not been tightened to this test's 1e-11 because that needs a device run to confirm."""
import numpy as np

from oracle.refmodel import generate_model_function
from oracle.samplers import ReplayRNG, simple_nuts
from util import LINEAR, linear_data, relerr


def test_oracle_nuts_amplifies_last_bit_perturbations_by_about_1e3():
    X, Y, _ = linear_data(200, 5)
    ref = generate_model_function(LINEAR, data=dict(X=X, Y=Y), gradient=True, beta=np.zeros(5))
    flip = np.random.default_rng(1)

    def perturbed(beta):
        lp, g = ref(beta)
        return lp, g * (1.0 + 2.0 ** -52 * flip.integers(-1, 2, size=g.shape))     # +- 1 ulp, or unchanged

    steps, burnin = 30, 5
    rng = np.random.default_rng(13)
    normals, unif = rng.standard_normal((steps, 5)), rng.random((steps, 270))
    a = simple_nuts(ref, ref.init, ReplayRNG(normals, unif), steps, burnin)
    b = simple_nuts(perturbed, ref.init, ReplayRNG(normals, unif), steps, burnin)
    assert np.array_equal(a["jmax"], b["jmax"]) and np.array_equal(a["accept"], b["accept"])
    early = relerr(b["draws"][:3], a["draws"][:3])
    late = relerr(b["draws"], a["draws"])
    assert early < 1e-15
    assert 1e-14 < late < 1e-11                       # amplified, but far from chaotic over this horizon
    assert relerr(b["epsilon"], a["epsilon"]) < 1e-11
    # the step size does overshoot the stability limit 2 / sqrt(lambda_max(X'X + I)) while adapting
    lim = 2.0 / np.sqrt(np.linalg.eigvalsh(X.T @ X + np.eye(5)).max())
    assert (a["epsilon"] > lim).sum() >= 3
