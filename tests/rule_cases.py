"""The op x argument-shape matrix of /root/reference/test/test.jl:142-224, as data.

Each case is (expression text, {name: value}, name of the differentiated parameter).  Mirrors the
`@mtest pattern f(args) rules...` macro (test.jl:125-135): fixtures v0ref/v1ref/v2ref (test.jl:12-14),
per-argument domain fix-ups, +0.001 shift of the differentiated argument (test.jl:119).
"""
import itertools

import numpy as np

V0 = 2.0
V1 = np.array([2.0, 3.0, 0.1, 0.0, -5.0])
V2 = np.array([[-1.0, 3.0, 0.0], [0.0, 5.0, -2.0]])
REFS = [V0, V1, V2]


def _patterns(kind, arity):
    full = [list(reversed(p)) for p in itertools.product([0, 1], repeat=arity)]
    full = sorted(full, key=lambda p: sum(b << j for j, b in enumerate(p)))
    p1 = full + [[2 * b for b in p] for p in full[1:]]
    if kind == 1:
        return p1
    if kind == 2:
        return [[0] * arity]
    if kind == 3:
        return [[1] * arity]
    if kind == 4:
        return [[2] * arity]
    if kind == 5:
        return [p for p in p1 if sum(p) <= 1]
    raise ValueError(kind)


def _apply(fix, v):
    if isinstance(v, np.ndarray):
        return np.vectorize(fix, otypes=[np.float64])(v)
    return float(fix(v))


def mtest(kind, fname, argnames, fixes=None, skip=(), infix=False):
    fixes = fixes or {}
    cases = []
    for pat in _patterns(kind, len(argnames)):
        vals = {}
        for nm, c in zip(argnames, pat):
            v = REFS[c].copy() if c else REFS[c]
            if nm in fixes:
                v = _apply(fixes[nm], v)
            vals[nm] = v
        for p, nm in enumerate(argnames):
            if nm in skip:
                continue
            names = ["arg%d" % i for i in range(len(argnames))]
            data = {names[i]: vals[a] for i, a in enumerate(argnames) if i != p}
            names[p] = "x"
            if infix:
                ex = (" %s " % fname).join(names)
            else:
                ex = "%s(%s)" % (fname, ", ".join(names))
            x0 = vals[nm] + 0.001
            cases.append((ex, data, x0))
    return cases


def all_cases():
    c = []
    nz = lambda y: 0.1 if y == 0 else y
    pos = lambda x: 0.1 if x <= 0 else x
    c += mtest(1, "+", ["x", "y"], infix=True)
    c += mtest(1, "+", ["x", "y", "z"], infix=True)
    c += mtest(1, "sum", ["x"])
    c += mtest(1, "-", ["x", "y"], infix=True)
    c += mtest(1, ".*", ["x", "y"], infix=True)
    c += mtest(1, "./", ["x", "y"], {"y": nz}, infix=True)
    c += mtest(1, ".^", ["x", "y"], {"x": lambda x: 0.2 if x <= 0 else x}, infix=True)
    for f in ("sin", "abs", "cos", "exp"):
        c += mtest(1, f, ["x"])
    c += mtest(1, "log", ["x"], {"x": pos})
    c += mtest(1, "transpose", ["x"])
    c.append(("x'", {}, np.array([-3.0, 2.0, 0.0])))
    c += mtest(1, "max", ["x", "y"])
    c += mtest(1, "min", ["x", "y"])
    c += mtest(2, "^", ["x", "y"], infix=True)
    c += mtest(5, "/", ["x", "y"], {"y": nz}, infix=True)
    c += mtest(5, "*", ["x", "y"], infix=True)
    tz = V1.reshape(1, -1)
    c.append(("x*tz", {"tz": tz}, np.array([-3.0, 2.0, 0.0])))
    c.append(("tz*x", {"tz": tz}, V1.copy()))
    c.append(("v2ref*x", {"v2ref": V2}, np.array([-3.0, 2.0, 0.0])))
    c.append(("v2ref[:,1:2]*x", {"v2ref": V2}, np.array([[-3.0, 2.0, 0.0], [1.0, 1.0, -2.0]])))
    c += mtest(2, "dot", ["x", "y"])
    c += mtest(3, "dot", ["x", "y"])
    # continuous distributions, test.jl:176-184
    c += mtest(1, "logpdfNormal", ["mu", "sigma", "x"], {"sigma": pos})
    c += mtest(1, "logpdfUniform", ["a", "b", "x"], {"a": lambda a: a - 10, "b": lambda b: b + 10})
    clamp = lambda x: min(max(x, 0.01), 0.99)
    c += mtest(1, "logpdfBeta", ["a", "b", "x"], {"x": clamp, "a": pos, "b": pos})
    c += mtest(1, "logpdfTDist", ["df", "x"], {"df": pos})
    c += mtest(1, "logpdfExponential", ["sc", "x"], {"sc": pos, "x": pos})
    c += mtest(1, "logpdfGamma", ["sh", "sc", "x"], {"sh": pos, "sc": pos, "x": pos})
    c += mtest(1, "logpdfCauchy", ["mu", "sc", "x"], {"sc": pos})
    c += mtest(1, "logpdflogNormal", ["lmu", "lsc", "x"], {"lsc": pos, "x": pos})
    c += mtest(1, "logpdfWeibull", ["sh", "sc", "x"], {"sh": pos, "sc": pos, "x": pos})
    # discrete distributions, test.jl:187-191
    c += mtest(1, "logpdfBernoulli", ["prob", "x"], {"prob": clamp, "x": lambda x: float(x > 0)}, skip=("x",))
    c += mtest(1, "logpdfPoisson", ["l", "x"], {"l": pos, "x": lambda x: float(round(abs(x)))}, skip=("x",))
    c += mtest(1, "logpdfBinomial", ["n", "prob", "x"],
               {"prob": clamp, "x": lambda x: float(round(abs(x))), "n": lambda n: float(round(abs(n) + 10))},
               skip=("n", "x"))
    # refs, test.jl:211-224
    for r in ("x[2]", "x[2:3]", "x[2:end]"):
        c.append((r, {}, V1 + 0.0))
    for r in ("x[2:end]", "x[2]", "x[2:4]", "x[:,2]", "x[1,:]", "x[2:end,:]", "x[:,2:end]", "x[2]+x[1]",
              "log(x[2]^2+x[1]^2)"):
        c.append((r, {}, V2 + 0.0))
    return c
