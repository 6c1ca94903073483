"""CPU half of the random-model parity run (tests/fuzz_models.py): the oracle derives every generated graph, its gradient
agrees with central finite differences wherever the reference's own rules are exact (no slice / gather, whose adjoint
the reference assigns instead of accumulating -- SURVEY.md Appendix C Q2 -- and no max/min of a value with itself), and
the host lowering serialises a tape for it.  The device comparison is tests/test_gpu_fuzz.py."""
import numpy as np

from fuzz_models import make_data, make_params, random_model
from oracle.refmodel import generate_model_function


def _case(seed, n=23, m=5):
    rng = np.random.default_rng(seed)
    return random_model(seed), make_data(rng, n, m), make_params(rng, n, m)


def test_oracle_gradient_of_random_graphs_matches_finite_differences():
    checked = 0
    for seed in range(290):
        model, data, pars = _case(seed)
        ref = generate_model_function(model, data=data, gradient=True, init=pars)
        x0 = np.concatenate([np.atleast_1d(v).ravel() for _, v in pars])
        lp, g = ref(x0)
        assert np.isfinite(lp), (seed, model)
        if "[" in model or "max(" in model or "min(" in model:
            continue
        f = generate_model_function(model, data=data, gradient=False, init=pars)
        fd = np.zeros_like(x0)
        for i in range(x0.size):
            h = 1e-6 * max(1.0, abs(x0[i]))
            xp, xm = x0.copy(), x0.copy()
            xp[i] += h
            xm[i] -= h
            fd[i] = (f(xp) - f(xm)) / (2 * h)
        assert np.max(np.abs(fd - g)) <= 1e-6 * max(1.0, np.max(np.abs(g))), (seed, model)
        checked += 1
    assert checked >= 60


def test_lowering_serialises_every_random_graph():
    from smc_pkg import load_package
    load_package()
    from simplemcmc_jl_b200.lowering import Lowering
    for seed in range(290):
        model, data, pars = _case(seed, *((1501, 7) if seed >= 250 else (23, 5)))
        tape, bound = Lowering(model, pars, data).serialise()
        assert len(tape) > 0
