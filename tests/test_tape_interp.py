"""The tape the host front end emits, executed in NumPy (tests/tape_interp.py) against the oracle's model function
(oracle/refmodel.py): statement order, register programs, SMC_F_ACC and SMC_F_FWD_ONLY mean what include/simplemcmc_tape.h
says -- checked without a GPU for the reference's example models and the random graphs of tests/fuzz_models.py.  The device
comparison of the same tapes is tests/test_gpu_parity.py / tests/test_gpu_fuzz.py."""
import numpy as np
import pytest

from smc_pkg import load_package

load_package()
from fuzz_models import make_data, make_params, random_model  # noqa: E402
from oracle.refmodel import generate_model_function  # noqa: E402
from simplemcmc_jl_b200 import fusion  # noqa: E402
from simplemcmc_jl_b200 import lowering as L  # noqa: E402
from simplemcmc_jl_b200.lowering import Lowering  # noqa: E402
from tape_interp import evaluate  # noqa: E402
from util import (HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, logistic_data, ou_data,  # noqa: E402
                  relerr, sleepstudy_data)


def _examples():
    X, Y, _ = linear_data(200, 6)
    yield "linear", LINEAR, [("beta", 0.1 * np.arange(6.0))], dict(X=X, Y=Y)
    X, Y, _ = logistic_data(200, 6)
    yield "binomial", LOGISTIC, [("vars", 0.05 * np.arange(6.0) - 0.1)], dict(X=X, Y=Y)
    yield "ornstein", ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=ou_data(300))
    data, (N, D, Lq) = hierarchical_data()
    rng = np.random.default_rng(2)
    yield "hierarchical", HIERARCHICAL, [("mu", 0.1 * rng.standard_normal((1, Lq))), ("sigma", 1.0 + 0.1 * rng.random((1, Lq))),
                                         ("beta", 0.2 * rng.standard_normal((D, Lq)))], data
    data, nsub = sleepstudy_data()
    yield "sleepstudy", SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", 3.0 * rng.standard_normal(nsub)),
                                      ("rand_slope", rng.standard_normal(nsub))], data


def _check(model, init, data, tol=1e-11, **kw):
    ref = generate_model_function(model, data=data, gradient=True, init=init)
    x0 = np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)).flatten(order="F") for _, v in init])
    lp, g = ref(x0)
    out = {}
    for fs, patterns in ((False, False), (True, False), (True, True)):       # plain statements; + programs; + GLM / AFFNORM (the default)
        low = Lowering(model, init, data, fuse=patterns, fuse_statements=fs, **kw)
        low.serialise()
        l1, g1, n1 = evaluate(low, x0, True)
        l0, _, n0 = evaluate(low, x0, False)
        assert relerr(l1, lp) < tol and relerr(l0, lp) < tol, (fs, patterns, l1, l0, lp)
        assert relerr(g1, g) < tol, (fs, patterns, g1, g)
        if not patterns:
            out[fs] = (low, n1, n0)
    return out


@pytest.mark.parametrize("case", list(_examples()), ids=[c[0] for c in _examples()])
def test_emitted_tape_evaluates_to_the_oracle(case):
    _, model, init, data = case
    _check(model, init, data)


def test_golden_tapes_evaluate_to_the_oracle():
    """the tapes committed under tests/golden/ (tools/make_golden_tapes.py: the front end's DEFAULT lowering, with the fused
    GLM and AFFNORM statements) through the NumPy executor, at a point away from the initial values"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import make_golden_tapes as G
    seen = set()
    for name, model, init, data, kw in G.cases():
        if kw.get("gradient") is False:
            continue
        rng = np.random.default_rng(len(name))
        x0 = np.concatenate([np.atleast_1d(np.asarray(v, dtype=np.float64)).flatten(order="F") for _, v in init])
        x = x0 + 0.05 * (1.0 + np.abs(x0)) * rng.random(x0.size)
        low = Lowering(model, init, data, **kw)
        low.serialise()
        seen |= {st.op for st in low.tape_fwd + low.tape_bwd}
        ref = generate_model_function(model, data=data, gradient=True, init=init)
        lp, g = ref(x)
        if kw.get("ref_adjoint") == "add":                 # the corrected gather adjoint (TODO.md:11): the reference's differs by design
            continue
        l1, g1, _ = evaluate(low, x, True)
        l0, _, _ = evaluate(low, x, False)
        assert relerr(l1, lp) < 1e-11 and relerr(l0, lp) < 1e-11, (name, l1, l0, lp)
        assert relerr(g1, g) < 1e-10, (name, g1, g)
    assert {L.OP["GLM"], L.OP["AFFNORM"], fusion.OP_FUSED, L.OP["MATMUL"], L.OP["GATHER"]} <= seen


def test_gradient_evaluation_walks_each_index_space_once():
    """sleepstudy: the two priors share one program and one reduction slot; both forward programs (the priors over the 18
    subjects, the likelihood over the 180 rows) are forward-only: with a gradient, the program that computes the adjoints of
    an index space recomputes the forward values AND takes the `acc += sum(logpdf)` over -- six statements per evaluation
    with gradient (two matrix products each way, rows, subjects) where there were ten"""
    data, nsub = sleepstudy_data()
    init = [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))]
    low, n_grad, n_plain = _check(SLEEPSTUDY2, init, data)[True]
    assert (n_grad, n_plain) == (6, 4)
    fwd_only = [st for st in low.tape_fwd if st.flags & L.F_FWD_ONLY]
    assert sorted(low.programs[st.aux0].n for st in fwd_only) == [nsub, 180] and all(st.op == fusion.OP_FUSED for st in fwd_only)
    res = low.tape_result.vid
    for n in (nsub, 180):          # one program per index space adds to the accumulator when the gradient is evaluated ...
        taking_over = [p for p in low.programs if p.n == n and (res, 1) in p.reductions and
                       any(ins[0] == L.OP["DLOGPDF"] for ins in p.body)]
        assert len(taking_over) == 1, n
    priors = [p for p in low.programs if p.n == nsub and p.reductions == [(res, 1)] and      # ... and one when it is not
              not any(ins[0] == L.OP["DLOGPDF"] for ins in p.body)]
    assert len(priors) == 1 and sum(1 for ins in priors[0].body if ins[0] == fusion.P_REDUCE) == 2
    assert not any(st.flags & L.F_FWD_ONLY for st in low.tape_bwd)
    # every writer of the accumulator adds to it: the library keeps the register zero between evaluations
    assert all(acc == 1 for p in low.programs for (vid, acc) in p.reductions if vid == res)


def test_random_graphs_through_the_emitted_tape():
    """every seed tests/test_fuzz_cpu.py and tests/test_gpu_fuzz.py use, the long-vector ones (n = 1501) included"""
    from oracle.refmodel import ModelError
    checked = 0
    for seed in range(290):
        model = random_model(seed)
        rng = np.random.default_rng(seed)
        n, m = (1501, 7) if seed >= 250 else (23, 5)
        data, pars = make_data(rng, n, m), make_params(rng, n, m)
        try:
            _check(model, pars, data, tol=1e-9)
        except (NotImplementedError, ModelError):       # (an operation the executor lacks; a draw outside the model's support)
            continue
        checked += 1
    assert checked >= 280
