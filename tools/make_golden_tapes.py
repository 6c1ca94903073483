"""Writes tests/golden/<model>.tape: the device op tape (include/simplemcmc_tape.h) of the five example models of the
reference, as emitted by simplemcmc.jl_b200/lowering.py for the data of tests/util.py.  They are the language-neutral
conformance artefact of the drop-in boundary: a Julia emitter replacing src/parsing.jl:488-492 must
produce the same bytes for the same model, initial values and data shapes.  tests/test_golden_tapes.py checks that the
Python front end still does.   python tools/make_golden_tapes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from smc_pkg import load_package
load_package()
from simplemcmc_jl_b200.lowering import Lowering
from util import (HIERARCHICAL, LINEAR, LOGISTIC, ORNSTEIN, SLEEPSTUDY2, hierarchical_data, linear_data, logistic_data, ou_data,
                  sleepstudy_data)


def cases():
    X, Y, _ = linear_data(1000, 10)
    yield "linear", LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y), {}
    X, Y, _ = logistic_data(1000, 10)
    yield "binomial", LOGISTIC, [("vars", np.zeros(10))], dict(X=X, Y=Y), {}
    yield "ornstein_T1000", ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=ou_data(1000)), {}
    yield "ornstein_T5000_affnorm", ORNSTEIN, [("tau", 20.0), ("mu", 10.0), ("sigma", 0.1)], dict(x=ou_data(5000)), {}
    data, (N, D, L) = hierarchical_data()
    yield "hierarchical", HIERARCHICAL, [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))], data, {}
    yield "hierarchical_scatter_add", HIERARCHICAL, [("mu", np.zeros((1, L))), ("sigma", np.ones((1, L))), ("beta", np.zeros((D, L)))], data, dict(ref_adjoint="add")
    data, nsub = sleepstudy_data()
    yield "sleepstudy", SLEEPSTUDY2, [("intercept", 250.0), ("slope", 10.0), ("rand_int", np.zeros(nsub)), ("rand_slope", np.zeros(nsub))], data, {}
    X, Y, _ = linear_data(1000, 10)
    yield "linear_no_gradient", LINEAR, [("beta", np.zeros(10))], dict(X=X, Y=Y), dict(gradient=False)


def build(model, init, data, kw):
    return Lowering(model, init, data, affnorm=True, **kw).serialise()


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    for name, model, init, data, kw in cases():
        tape, bound = build(model, init, data, kw)
        with open(os.path.join(out, name + ".tape"), "wb") as f:
            f.write(tape)
        print("%-28s %6d bytes, binds %s" % (name, len(tape), [b[0] for b in bound]))
