"""ORACLE (test infrastructure only -- the product path never imports this file).

CPU restatement of SimpleMCMC.jl's model compiler and generated log-density closure:

  parseModel!        /root/reference/src/parsing.jl:76-120
  unfold!            /root/reference/src/parsing.jl:123-183
  uniqueVars!        /root/reference/src/parsing.jl:188-208
  categorizeVars!    /root/reference/src/parsing.jl:219-247
  backwardSweep!     /root/reference/src/parsing.jl:250-300
  setInit!/betaAssign/root/reference/src/parsing.jl:303-349
  preCalculate       /root/reference/src/parsing.jl:354-384
  generateModelFunction  /root/reference/src/parsing.jl:394-493
  @dfunc rule table  /root/reference/src/diff.jl:29-178   (RULES below, same formulas, our own text layout)
  derive / hint      /root/reference/src/diff.jl:182-220

The statement list (`m.exprs ++ m.dexprs`) is evaluated with NumPy under 2013-Julia array
semantics (column-major, scalar<->array broadcasting in + - .* ./, `*` = matrix product), in the
order the reference executes it.  Densities come from oracle/densities.py.

Parity status: the reference cannot be run in this container (no Julia; the source is Julia<=0.2
syntax), and it ships no golden vectors for this path, so the *statement lists* are pinned against
SURVEY.md Appendix A (hand-derived from the reference), the *rules* against central finite
differences with test/test.jl's fixtures and tolerance, and the libRmath-backed densities against
mpmath closed forms: "parity unpinned" for absolute density values.
"""
import itertools
import math

import numpy as np

from .jlparse import Expr, Sym, parse, parse_expr, show
from . import densities as D

ACC_SYM = Sym("__acc")      # src/SimpleMCMC.jl:14
PARAM_SYM = Sym("__beta")   # src/SimpleMCMC.jl:15
TEMP_NAME = "tmp"           # src/SimpleMCMC.jl:16
DERIV_PREFIX = "d"          # src/SimpleMCMC.jl:17


class GiveUpEval(Exception):
    """the string exception "give up eval" (src/distribs.jl:15,17,32,70)"""


class ModelError(Exception):
    pass


_gensym_counter = itertools.count(1)


def gensym(tag):
    return Sym("##%s#%d" % (tag, next(_gensym_counter)))


# --------------------------------------------------------------------------- symbols helpers
_CMP_OPS = {">", "<", ">=", "<=", ".>", ".<", ".<=", ".>=", "=="}


def get_symbols(ex):
    """src/parsing.jl:29-37"""
    if isinstance(ex, Sym):
        return {ex}
    if isinstance(ex, (list, tuple)):
        out = set()
        for e in ex:
            out |= get_symbols(e)
        return out
    if isinstance(ex, Expr):
        if ex.head == "call":
            return get_symbols(ex.args[1:])
        if ex.head == "ref":
            return get_symbols(ex.args) - {Sym(":"), Sym("end")}
        if ex.head == "comparison":
            return get_symbols(ex.args) - {Sym(o) for o in _CMP_OPS}
        return get_symbols(ex.args)
    return set()


def subst_symbols(ex, smap):
    """src/parsing.jl:41-46"""
    if isinstance(ex, Sym):
        return smap.get(ex, ex)
    if isinstance(ex, Expr):
        if ex.head == "call":
            return Expr("call", ex.args[0], *[subst_symbols(a, smap) for a in ex.args[1:]])
        return Expr(ex.head, *[subst_symbols(a, smap) for a in ex.args])
    return ex


class Param(object):
    """MCMCParams, src/parsing.jl:49-53 (map is a 0-based python slice/int here)"""

    def __init__(self, sym, size, start, count):
        self.sym, self.size, self.start, self.count = sym, list(size), start, count


class Model(object):
    """MCMCModel, src/parsing.jl:56-69"""

    def __init__(self):
        self.bsize = 0
        self.pars = []
        self.init = []
        self.source = None
        self.exprs = []
        self.dexprs = []
        self.finalacc = ACC_SYM
        self.varsset = set()
        self.pardesc = set()
        self.accanc = set()


# --------------------------------------------------------------------------- first pass
def parse_model(m, source):
    """parseModel!, src/parsing.jl:76-120"""
    found = [False]

    def explore(ex):
        if not isinstance(ex, Expr):
            return ex
        h = ex.head
        if h in ("ref", "=", "vcat"):
            return ex
        if h == "+=":
            return Expr("=", ex.args[0], Expr("call", Sym("+"), *ex.args))
        if h == "-=":
            return Expr("=", ex.args[0], Expr("call", Sym("-"), *ex.args))
        if h == "*=":
            return Expr("=", ex.args[0], Expr("call", Sym("*"), *ex.args))
        if h == "block":
            return Expr("block", *[explore(a) for a in ex.args])
        if h == "call":
            if ex.args[0] != "~":
                return ex
            found[0] = True
            dist = ex.args[2]
            fn = Sym("logpdf%s" % dist.args[0])
            return Expr("=", ACC_SYM, Expr("call", Sym("+"), ACC_SYM, Expr("call", fn, *(dist.args[1:] + [ex.args[1]]))))
        raise ModelError("[parseModel] unmanaged expr type %s" % h)

    if not (isinstance(source, Expr) and source.head == "block" and len(source.args) >= 1):
        raise ModelError("model should contain be a block with at least 1 statement")
    m.source = explore(source)
    if not found[0]:
        m.source.args[-1] = Expr("=", ACC_SYM, m.source.args[-1])


# --------------------------------------------------------------------------- unfold
def unfold(m):
    """unfold!, src/parsing.jl:123-183"""

    def explore(ex):
        if not isinstance(ex, Expr):
            return ex
        h = ex.head
        if h in ("ref", "comparison"):
            return ex
        if h == "vcat":
            return explore(Expr("call", Sym("vcat"), *ex.args))
        if h == "'":
            return explore(Expr("call", Sym("transpose"), ex.args[0]))
        if h == "block":
            last = None
            for a in ex.args:
                last = explore(a)
            return last
        if h == "=":
            lhs, rhs = ex.args
            if not (isinstance(lhs, Sym) or (isinstance(lhs, Expr) and lhs.head == "ref")):
                raise ModelError("[unfold] not a symbol on LHS of assigment %s" % show(ex))
            if isinstance(rhs, Sym) or isinstance(rhs, (int, float)):
                m.exprs.append(Expr("=", lhs, rhs))
            elif isinstance(rhs, Expr):
                ue = explore(rhs)
                m.exprs.append(Expr("=", lhs, ue))
            else:
                raise ModelError("[unfold] can't handle RHS of assignment %s" % show(ex))
            return lhs
        if h == "call":
            fname = ex.args[0]
            args = list(ex.args[1:])
            if fname in ("+", "*", "sum"):
                while len(args) > 2:        # right-nest n-ary calls, parsing.jl:160-166
                    a2 = args.pop()
                    a1 = args.pop()
                    args.append(Expr("call", fname, a1, a2))
            na = [fname]
            for e2 in args:
                if isinstance(e2, Expr):
                    ue = explore(e2)
                    nv = gensym(TEMP_NAME)
                    m.exprs.append(Expr("=", nv, ue))
                    na.append(nv)
                else:
                    na.append(e2)
            return Expr("call", *na)
        raise ModelError("[unfold] unmanaged expr type %s" % h)

    explore(m.source)


def unique_vars(m):
    """uniqueVars!, src/parsing.jl:188-208"""
    subst = {}
    used = {ACC_SYM}
    for st in m.exprs:
        st.args[1] = subst_symbols(st.args[1], subst)
        lhs = sorted(get_symbols(st.args[0]))[0] if not isinstance(st.args[0], Sym) else st.args[0]
        if isinstance(st.args[0], Expr):
            lhs = st.args[0].args[0]
        if lhs in used:
            subst[lhs] = gensym(str(lhs))
            st.args[0] = subst_symbols(st.args[0], subst)
        else:
            used.add(lhs)
    m.finalacc = subst.get(ACC_SYM, ACC_SYM)


def _lhs_symbol(st):
    return st.args[0] if isinstance(st.args[0], Sym) else st.args[0].args[0]


def categorize_vars(m):
    """categorizeVars!, src/parsing.jl:219-247"""
    m.varsset = {_lhs_symbol(st) for st in m.exprs}
    parset = {p.sym for p in m.pars}
    m.pardesc = set(parset)
    for st in m.exprs:
        if get_symbols(st.args[1]) & m.pardesc:
            m.pardesc.add(_lhs_symbol(st))
    m.accanc = {m.finalacc}
    for st in reversed(m.exprs):
        lhs = _lhs_symbol(st)
        rhs = get_symbols(st) - {lhs}
        if lhs in m.accanc:
            m.accanc |= rhs
    if m.finalacc not in m.pardesc:
        raise ModelError("Model parameters do not seem to influence model outcome")
    missing = parset - m.accanc
    if missing:
        raise ModelError("Model parameter(s) %s do not seem to influence model outcome" % sorted(missing))


# --------------------------------------------------------------------------- init / beta map
def set_init(m, init):
    """setInit!, src/parsing.jl:303-334 (kwarg order = beta order, matrices column-major)"""
    if len(init) < 1:
        raise ModelError("There should be at leat one parameter specified, none found")
    for par, val in init:
        par = Sym(par)
        if isinstance(val, (int, float)) and not isinstance(val, bool):
            m.pars.append(Param(par, [], m.bsize, 1))
            m.bsize += 1
            m.init.append(float(val))
        else:
            arr = np.asarray(val, dtype=np.float64)
            if arr.ndim == 1:
                nb = arr.shape[0]
                m.pars.append(Param(par, [nb], m.bsize, nb))
                m.bsize += nb
                m.init.extend(arr.tolist())
            elif arr.ndim == 2:
                nb = arr.shape[0] * arr.shape[1]
                m.pars.append(Param(par, list(arr.shape), m.bsize, nb))
                m.bsize += nb
                m.init.extend(arr.flatten(order="F").tolist())
            else:
                raise ModelError("[setInit] forbidden parameter type for %s" % par)


def beta_assign(m, env, beta):
    """betaAssign, src/parsing.jl:338-349"""
    for p in m.pars:
        seg = np.array(beta[p.start:p.start + p.count], dtype=np.float64)
        if len(p.size) == 0:
            env[p.sym] = float(seg[0])
        elif len(p.size) == 1:
            env[p.sym] = seg
        else:
            env[p.sym] = seg.reshape(p.size, order="F")


# --------------------------------------------------------------------------- Julia-0.2 array semantics
def is_real(v):
    return isinstance(v, (int, float, bool, np.floating, np.integer, np.bool_))


def _f(v):
    if isinstance(v, np.ndarray):
        return v if v.dtype == np.float64 else v.astype(np.float64)
    return float(v)


def _shape_merge(a, b):
    """promote_shape: equal numel with trailing singleton dims differing -> keep the larger ndims"""
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray) and a.shape != b.shape:
        if a.size == b.size:
            if a.ndim >= b.ndim:
                return a, b.reshape(a.shape, order="F")
            return a.reshape(b.shape, order="F"), b
        raise ModelError("dimension mismatch %s vs %s" % (a.shape, b.shape))
    return a, b


def jl_add(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    return a + b


def jl_sub(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    return a - b


def jl_emul(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    return a * b


def jl_ediv(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    with np.errstate(divide="ignore", invalid="ignore"):
        return a / b


def jl_epow(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return np.power(a, b) if (isinstance(a, np.ndarray) or isinstance(b, np.ndarray)) else _scalar_pow(a, b)


def _scalar_pow(a, b):
    try:
        return float(np.power(np.float64(a), np.float64(b)))
    except Exception:
        return float("nan")


def jl_mul(a, b):
    a, b = _f(a), _f(b)
    if is_real(a) or is_real(b):
        return a * b
    A = a if a.ndim == 2 else a.reshape(-1, 1)
    B = b if b.ndim == 2 else b.reshape(-1, 1)
    if A.shape[1] != B.shape[0]:
        raise ModelError("matrix product dimension mismatch %s * %s" % (a.shape, b.shape))
    R = A @ B
    if b.ndim == 1:
        return R[:, 0].copy()
    return R


def jl_div(a, b):
    a, b = _f(a), _f(b)
    if isinstance(a, np.ndarray) and isinstance(b, np.ndarray):
        raise ModelError("/ needs at least one scalar")
    with np.errstate(divide="ignore", invalid="ignore"):
        return a / b


def jl_pow(a, b):
    if not (is_real(a) and is_real(b)):
        raise ModelError("^ is scalar only")
    return _scalar_pow(a, b)


def jl_transpose(a):
    if is_real(a):
        return a
    a = _f(a)
    if a.ndim == 1:
        return a.reshape(1, -1).copy()
    return a.T.copy()


def jl_sum(a):
    if is_real(a):
        return float(a)          # sum(x::Real) = x, src/SimpleMCMC.jl:7
    return float(np.sum(_f(a).flatten(order="F")))


def jl_dot(a, b):
    if is_real(a) and is_real(b):
        return float(a) * float(b)
    return float(np.dot(_f(a).flatten(order="F"), _f(b).flatten(order="F")))


def _un(fn):
    def g(a):
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            r = fn(_f(a))
        return r if isinstance(r, np.ndarray) else float(r)
    return g


def jl_max(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    r = np.maximum(a, b)
    return r if isinstance(r, np.ndarray) else float(r)


def jl_min(a, b):
    a, b = _shape_merge(_f(a), _f(b))
    r = np.minimum(a, b)
    return r if isinstance(r, np.ndarray) else float(r)


def _cmp(op, a, b):
    a, b = _shape_merge(_f(a), _f(b))
    r = {">": np.greater, ".>": np.greater, "<": np.less, ".<": np.less, "<=": np.less_equal,
         ".<=": np.less_equal, ">=": np.greater_equal, ".>=": np.greater_equal, "==": np.equal}[op](a, b)
    return r


FUNCS = {
    "+": lambda *a: a[0] if len(a) == 1 else jl_add(*a),
    "-": lambda *a: (-_f(a[0])) if len(a) == 1 else jl_sub(*a),
    "*": jl_mul, ".*": jl_emul, "/": jl_div, "./": jl_ediv, "^": jl_pow, ".^": jl_epow,
    "sum": jl_sum, "dot": jl_dot, "transpose": jl_transpose,
    "exp": _un(np.exp), "log": _un(np.log), "sin": _un(np.sin), "cos": _un(np.cos),
    "abs": _un(np.abs), "sign": _un(np.sign), "sqrt": _un(np.sqrt),
    "digamma": _un(D.digamma), "max": jl_max, "min": jl_min,
    "zero": lambda a: 0.0 if is_real(a) else np.zeros_like(_f(a)),
}
for _name in D.DISTRIBUTIONS:
    FUNCS["logpdf" + _name] = getattr(D, "logpdf" + _name)


def _resolve_index(ix, dimlen, env):
    """one index of a ref -> (0-based int array, is_scalar)"""
    def ev(e):
        if isinstance(e, Sym) and e == "end":
            return dimlen
        return evaluate(e, env, end_value=dimlen)
    if isinstance(ix, Sym) and ix == ":":
        return np.arange(dimlen), False
    if isinstance(ix, Expr) and ix.head == ":":
        a, b = int(ev(ix.args[0])), int(ev(ix.args[1]))
        return np.arange(a - 1, b), False
    v = ev(ix)
    if is_real(v):
        return np.array([int(v) - 1]), True
    v = np.asarray(v)
    if v.dtype == np.bool_:
        return np.nonzero(v.flatten(order="F"))[0], False
    return v.flatten(order="F").astype(np.int64) - 1, False


def ref_indices(ref, env, base=None):
    """linear (column-major, 0-based) indices addressed by a ref + the Julia-0.2 result shape"""
    arr = env[ref.args[0]] if base is None else base
    idx = ref.args[1:]
    if not isinstance(arr, np.ndarray):
        raise ModelError("ref on a scalar: %s" % show(ref))
    if len(idx) == 1:
        lin, scalar = _resolve_index(idx[0], arr.size, env)
        return lin, (() if scalar else (len(lin),))
    if len(idx) == 2 and arr.ndim == 2:
        r, rs = _resolve_index(idx[0], arr.shape[0], env)
        c, cs = _resolve_index(idx[1], arr.shape[1], env)
        lin = (r[:, None] + c[None, :] * arr.shape[0]).flatten(order="F")
        if rs and cs:
            shape = ()
        elif cs:               # trailing scalar index dropped: x[:,2] is a vector
            shape = (len(r),)
        else:                  # x[1,:] stays 1xn in Julia 0.2
            shape = (len(r), len(c))
        return lin, shape
    raise ModelError("unsupported ref %s" % show(ref))


def get_ref(ref, env):
    arr = _f(env[ref.args[0]])
    lin, shape = ref_indices(ref, env)
    flat = arr.flatten(order="F")[lin]
    if shape == ():
        return float(flat[0])
    return flat.reshape(shape, order="F").copy()


def set_ref(ref, env, value):
    arr = _f(env[ref.args[0]])
    lin, shape = ref_indices(ref, env)
    flat = arr.flatten(order="F")
    if is_real(value):
        flat[lin] = value
    else:
        vals = _f(value).flatten(order="F")
        for k in range(len(lin)):     # sequential stores: with repeated indices the last one wins (Appendix C, Q2)
            flat[lin[k]] = vals[k]
    env[ref.args[0]] = flat.reshape(arr.shape, order="F")


def evaluate(ex, env, end_value=None):
    if isinstance(ex, Sym):
        if ex == "end" and end_value is not None:
            return end_value
        if ex not in env:
            raise ModelError("undefined variable %s" % ex)
        return env[ex]
    if not isinstance(ex, Expr):
        return ex
    h = ex.head
    if h == "call":
        fn = str(ex.args[0])
        args = [evaluate(a, env, end_value) for a in ex.args[1:]]
        if fn not in FUNCS:
            raise ModelError("unknown function %s" % fn)
        return FUNCS[fn](*args)
    if h == "ref":
        return get_ref(ex, env)
    if h == "comparison":
        vals = [evaluate(a, env, end_value) if i % 2 == 0 else a for i, a in enumerate(ex.args)]
        res = None
        for i in range(1, len(vals), 2):
            r = _cmp(str(vals[i]), vals[i - 1], vals[i + 1])
            res = r if res is None else np.logical_and(res, r)
        return res if isinstance(res, np.ndarray) else bool(res)
    if h == "'":
        return jl_transpose(evaluate(ex.args[0], env, end_value))
    if h == "block":
        last = None
        for a in ex.args:
            last = run_statement(a, env) if isinstance(a, Expr) and a.head == "=" else evaluate(a, env, end_value)
        return last
    if h == "=":
        return run_statement(ex, env)
    raise ModelError("cannot evaluate %s" % show(ex))


def run_statement(st, env):
    lhs, rhs = st.args
    val = evaluate(rhs, env)
    if isinstance(val, np.ndarray) and val.dtype == np.bool_:
        val = val.astype(np.float64)
    if isinstance(lhs, Sym):
        env[lhs] = val
    else:
        set_ref(lhs, env, val)
    return val


# --------------------------------------------------------------------------- derivation rules
# (function, argument names with optional Real/Array class, differentiated argument, adjoint body)
# Same mathematics as the @dfunc table of src/diff.jl:29-178; `ds` is the incoming adjoint.
R, A, X_ = "Real", "Array", None
RULES = []


def _rule(fn, sig, dv, body):
    RULES.append((fn, sig, dv, parse_expr(body)))


for _fn in ("+",):
    _rule(_fn, [("x", R), ("y", X_)], "x", "sum(ds)")
    _rule(_fn, [("x", A), ("y", X_)], "x", "+ds")
    _rule(_fn, [("x", X_), ("y", R)], "y", "sum(ds)")
    _rule(_fn, [("x", X_), ("y", A)], "y", "+ds")
_rule("-", [("x", X_)], "x", "-ds")
_rule("-", [("x", R), ("y", X_)], "x", "sum(ds)")
_rule("-", [("x", A), ("y", X_)], "x", "+ds")
_rule("-", [("x", X_), ("y", R)], "y", "-sum(ds)")
_rule("-", [("x", X_), ("y", A)], "y", "-ds")
_rule("sum", [("x", X_)], "x", "+ds")
_rule("dot", [("x", X_), ("y", X_)], "x", "y .* ds")
_rule("dot", [("x", X_), ("y", X_)], "y", "x .* ds")
_rule("log", [("x", X_)], "x", "ds ./ x")
_rule("exp", [("x", X_)], "x", "exp(x) .* ds")
_rule("sin", [("x", X_)], "x", "cos(x) .* ds")
_rule("cos", [("x", X_)], "x", "-sin(x) .* ds")
_rule("abs", [("x", X_)], "x", "sign(x) .* ds")
_rule("*", [("x", R), ("y", X_)], "x", "sum(ds .* y)")
_rule("*", [("x", A), ("y", X_)], "x", "ds * transpose(y)")
_rule("*", [("x", X_), ("y", R)], "y", "sum(ds .* x)")
_rule("*", [("x", X_), ("y", A)], "y", "transpose(x) * ds")
_rule(".*", [("x", R), ("y", X_)], "x", "sum(ds .* y)")
_rule(".*", [("x", A), ("y", X_)], "x", "ds .* y")
_rule(".*", [("x", X_), ("y", R)], "y", "sum(ds .* x)")
_rule(".*", [("x", X_), ("y", A)], "y", "ds .* x")
_rule("^", [("x", R), ("y", R)], "x", "y * x ^ (y-1) * ds")
_rule("^", [("x", R), ("y", R)], "y", "log(x) * x ^ y * ds")
_rule(".^", [("x", R), ("y", X_)], "x", "sum(y .* x .^ (y-1) .* ds)")
_rule(".^", [("x", A), ("y", X_)], "x", "y .* x .^ (y-1) .* ds")
_rule(".^", [("x", X_), ("y", R)], "y", "sum(log(x) .* x .^ y .* ds)")
_rule(".^", [("x", X_), ("y", A)], "y", "log(x) .* x .^ y .* ds")
_rule("/", [("x", R), ("y", X_)], "x", "sum(ds ./ y)")
_rule("/", [("x", A), ("y", R)], "x", "ds ./ y")
_rule("/", [("x", X_), ("y", R)], "y", "sum(- x ./ (y .* y) .* ds)")
_rule("/", [("x", R), ("y", A)], "y", "(- x ./ (y .* y)) .* ds")
_rule("./", [("x", R), ("y", X_)], "x", "sum(ds ./ y)")
_rule("./", [("x", A), ("y", X_)], "x", "ds ./ y")
_rule("./", [("x", X_), ("y", R)], "y", "sum(- x ./ (y .* y) .* ds)")
_rule("./", [("x", X_), ("y", A)], "y", "(- x ./ (y .* y)) .* ds")
_rule("max", [("x", R), ("y", X_)], "x", "sum((x .> y) .* ds)")
_rule("max", [("x", A), ("y", X_)], "x", "(x .> y) .* ds")
_rule("max", [("x", X_), ("y", R)], "y", "sum((x .< y) .* ds)")
_rule("max", [("x", X_), ("y", A)], "y", "(x .< y) .* ds")
_rule("min", [("x", R), ("y", X_)], "x", "sum((x .< y) .* ds)")
_rule("min", [("x", A), ("y", X_)], "x", "(x .< y) .* ds")
_rule("min", [("x", X_), ("y", R)], "y", "sum((x .> y) .* ds)")
_rule("min", [("x", X_), ("y", A)], "y", "(x .> y) .* ds")
_rule("transpose", [("x", R)], "x", "+ds")
_rule("transpose", [("x", A)], "x", "transpose(ds)")


def _dist_rules(fn, names, table):
    """table: {arg: elementwise-adjoint text}; emits the ::Real (summed) and ::Array variants"""
    for arg, (body, mult, pre) in table.items():
        for cls in (R, A):
            sig = [(n, cls if n == arg else X_) for n in names]
            core = ("sum(%s)" % body) if cls == R else ("(%s)" % body)
            txt = "%s %s ds" % (core, mult)
            if pre:
                txt = "(%s ; %s)" % (pre, txt)
            _rule(fn, sig, arg, txt)


_dist_rules("logpdfNormal", ["mu", "sigma", "x"], {
    "mu": ("(x - mu) ./ (sigma .* sigma)", "*", None),
    "sigma": ("((x - mu).*(x - mu) ./ (sigma.*sigma) - 1.) ./ sigma", "*", None),
    "x": ("(mu - x) ./ (sigma .* sigma)", "*", None)})
_dist_rules("logpdfUniform", ["a", "b", "x"], {
    "a": ("(a .<= x .<= b) ./ (b - a)", "*", None),
    "b": ("(a .<= x .<= b) ./ (a - b)", "*", None)})
_rule("logpdfUniform", [("a", X_), ("b", X_), ("x", X_)], "x", "zero(x)")
_dist_rules("logpdfWeibull", ["sh", "sc", "x"], {
    "sh": ("((1. - r.^sh) .* log(r) + 1./sh)", "*", "r = x./sc"),
    "sc": ("((x./sc).^sh - 1.) .* sh./sc", "*", None),
    "x": ("((1. - (x./sc).^sh) .* sh - 1.) ./ x", "*", None)})
_dist_rules("logpdfBeta", ["a", "b", "x"], {
    "x": ("(a-1) ./ x - (b-1) ./ (1-x)", "*", None),
    "a": ("digamma(a+b) - digamma(a) + log(x)", "*", None),
    "b": ("digamma(a+b) - digamma(b) + log(1-x)", "*", None)})
_dist_rules("logpdfTDist", ["df", "x"], {
    "x": ("-(df+1).*x ./ (df+x.*x)", ".*", None)})
# the df adjoint divides the (summed) bracket by 2 outside the sum, src/diff.jl:127-128
_rule("logpdfTDist", [("df", R), ("x", X_)], "df",
      "(tmp2 = (x.*x + df) ; sum( (x.*x-1)./tmp2 + log(df./tmp2) + digamma((df+1)/2) - digamma(df/2) ) / 2 .* ds )")
_rule("logpdfTDist", [("df", A), ("x", X_)], "df",
      "(tmp2 = (x.*x + df) ; ( (x.*x-1)./tmp2 + log(df./tmp2) + digamma((df+1)/2) - digamma(df/2) ) / 2 .* ds )")
_rule("logpdfExponential", [("sc", X_), ("x", R)], "x", "sum(-1/sc) .* ds")
_rule("logpdfExponential", [("sc", X_), ("x", A)], "x", "(- ds ./ sc)")
_rule("logpdfExponential", [("sc", R), ("x", X_)], "sc", "sum((x-sc)./(sc.*sc)) .* ds")
_rule("logpdfExponential", [("sc", A), ("x", X_)], "sc", "(x-sc) ./ (sc.*sc) .* ds")
_dist_rules("logpdfGamma", ["sh", "sc", "x"], {
    "x": ("-( sc + x - sh.*sc)./(sc.*x)", ".*", None),
    "sh": ("log(x) - log(sc) - digamma(sh)", ".*", None),
    "sc": ("(x - sc.*sh) ./ (sc.*sc)", ".*", None)})
_dist_rules("logpdfCauchy", ["mu", "sc", "x"], {
    "x": ("2(mu-x) ./ (sc.*sc + (x-mu).*(x-mu))", ".*", None),
    "mu": ("2(x-mu) ./ (sc.*sc + (x-mu).*(x-mu))", ".*", None),
    "sc": ("((x-mu).*(x-mu) - sc.*sc) ./ (sc.*(sc.*sc + (x-mu).*(x-mu)))", ".*", None)})
_dist_rules("logpdflogNormal", ["lmu", "lsc", "x"], {
    "x": ("(lmu - tmp2 - log(x)) ./ (tmp2.*x)", ".*", "tmp2=lsc.*lsc"),
    "lmu": ("(log(x) - lmu) ./ (lsc .* lsc)", ".*", None),
    "lsc": ("(lmu.*lmu - tmp2 - log(x).*(2lmu-log(x))) ./ (lsc.*tmp2)", ".*", "tmp2=lsc.*lsc")})
_dist_rules("logpdfBernoulli", ["p", "x"], {"p": ("1. ./ (p - (1. - x))", "*", None)})
_dist_rules("logpdfBinomial", ["n", "p", "x"], {"p": ("x ./ p - (n-x) ./ (1 - p)", "*", None)})
_dist_rules("logpdfPoisson", ["lambda", "x"], {"lambda": ("x ./ lambda - 1", "*", None)})
_rule("logpdfTestDiff", [("x", X_)], "x", "+ds")


def find_rule(fn, index, vals):
    """method dispatch of d_<fn>_x<index>(vals...) on Real vs Array sample values (diff.jl:197-202)"""
    best, best_score = None, -1
    for (rfn, sig, dv, body) in RULES:
        if rfn != fn or len(sig) != len(vals):
            continue
        names = [s[0] for s in sig]
        if names.index(dv) != index - 1:
            continue
        ok, score = True, 0
        for (nm, cls), v in zip(sig, vals):
            if cls == R:
                ok &= is_real(v)
                score += 1
            elif cls == A:
                ok &= isinstance(v, np.ndarray)
                score += 1
        if ok and score > best_score:
            best, best_score = (names, body), score
    return best


def derive(opex, index, dsym, vhint):
    """derive, src/diff.jl:192-220 -> list of statements, the last one `dv = dv + expr`"""
    vs = opex.args[index]
    ds = Sym("%s%s" % (DERIV_PREFIX, dsym)) if isinstance(dsym, Sym) else \
        Expr("ref", Sym("%s%s" % (DERIV_PREFIX, dsym.args[0])), *dsym.args[1:])
    args = opex.args[1:]
    vals = [vhint[a] if isinstance(a, Sym) else a for a in args]
    rule = find_rule(str(opex.args[0]), index, vals)
    if rule is None:
        raise ModelError("[derive] Doesn't know how to derive %s by argument %s" % (show(opex), vs))
    names, body = rule
    smap = {Sym(n): a for n, a in zip(names, args)}
    smap[Sym("ds")] = ds
    dexp = subst_symbols(body, smap)
    tmp = Model()
    tmp.source = Expr("=", Sym("dummy"), dexp)
    unfold(tmp)
    last = tmp.exprs[-1].args[1]
    dv = Sym("%s%s" % (DERIV_PREFIX, vs))
    tmp.exprs[-1] = Expr("=", dv, Expr("call", Sym("+"), dv, last))
    return tmp.exprs


def backward_sweep(m, vhint):
    """backwardSweep!, src/parsing.jl:250-300"""
    avars = m.accanc & m.pardesc

    def dname(s):
        return Sym("%s%s" % (DERIV_PREFIX, s))

    for st in reversed(m.exprs):
        lhs, rhs = st.args
        if isinstance(lhs, Sym):
            dsym, dsym2 = lhs, dname(lhs)
        elif isinstance(lhs, Expr) and lhs.head == "ref":
            dsym, dsym2 = lhs, Expr("ref", dname(lhs.args[0]), *lhs.args[1:])
        else:
            raise ModelError("[backwardSweep] not a symbol on LHS of assigment %s" % show(st))
        if isinstance(rhs, Sym):
            if rhs in avars:
                m.dexprs.append(Expr("=", dname(rhs), dsym2))
        elif not isinstance(rhs, Expr):
            pass
        elif rhs.head == "ref":
            if rhs.args[0] in avars:
                m.dexprs.append(Expr("=", Expr("ref", dname(rhs.args[0]), *rhs.args[1:]), dsym2))
        elif rhs.head == "call":
            for i in range(1, len(rhs.args)):
                vsym = rhs.args[i]
                if isinstance(vsym, Sym) and vsym in avars:
                    m.dexprs.extend(derive(rhs, i, dsym, vhint))
        else:
            raise ModelError("[backwardSweep] can't derive %s" % show(rhs))


# --------------------------------------------------------------------------- model function
class ModelFunction(object):
    """What generateModelFunction returns (src/parsing.jl:394-493): callable ll(beta)."""

    def __init__(self, m, data, gradient, vhint=None, source_nogr=None):
        self.m, self.data, self.gradient, self.vhint = m, data, gradient, vhint
        self.source_nogr = source_nogr
        self.nparams, self.pars, self.init = m.bsize, m.pars, np.array(m.init, dtype=np.float64)

    def statements(self):
        return [show(s) for s in self.m.exprs], [show(s) for s in self.m.dexprs]

    def _env0(self, beta):
        env = dict(self.data)
        beta_assign(self.m, env, np.asarray(beta, dtype=np.float64))
        env[ACC_SYM] = 0.0
        return env

    def __call__(self, beta):
        m = self.m
        if not self.gradient:
            # no-gradient closure runs the *un-unfolded* source, parsing.jl:469-481
            try:
                env = self._env0(beta)
                for st in m.source.args:
                    if isinstance(st, Expr):
                        run_statement(st, env) if st.head == "=" else evaluate(st, env)
                return float(env[ACC_SYM])
            except GiveUpEval:
                return -math.inf
        try:
            env = self._env0(beta)
            dsym = lambda v: Sym("%s%s" % (DERIV_PREFIX, v))
            env[dsym(m.finalacc)] = 1.0
            for v in (m.accanc & m.pardesc) - {m.finalacc}:    # parsing.jl:421-430
                vh = self.vhint[v]
                env[dsym(v)] = 0.0 if is_real(vh) else np.zeros(np.shape(vh), dtype=np.float64)
            for st in m.exprs:
                run_statement(st, env)
            for st in m.dexprs:
                run_statement(st, env)
            grads = []
            for p in m.pars:
                g = env[dsym(p.sym)]
                grads.append(np.atleast_1d(_f(g)).flatten(order="F"))
            return float(env[m.finalacc]), np.concatenate(grads)
        except GiveUpEval:
            return -math.inf, np.zeros(m.bsize)


def pre_calculate(m, data):
    """preCalculate, src/parsing.jl:354-384: run the model once at init and keep every value"""
    env = dict(data)
    beta_assign(m, env, np.array(m.init))
    env[ACC_SYM] = 0.0
    try:
        for st in m.exprs:
            run_statement(st, env)
    except GiveUpEval:
        raise ModelError("Initial values out of model support, try other values")
    res = env[m.finalacc]
    if not is_real(res):
        raise ModelError("Model outcome should be a scalar, %s found" % type(res).__name__)
    if res == -math.inf:
        raise ModelError("Initial values out of model support, try other values")
    return env


def generate_model_function(model, data=None, gradient=False, init=None, **kwinit):
    """generateModelFunction(model; gradient, init...) -> ModelFunction
    `model` is the text inside `quote ... end`; `init` is an ordered list of (name, value) pairs (or kwargs);
    `data` plays the role of the globals of Main (src/parsing.jl:484-485)."""
    pairs = list(init) if init is not None else []
    pairs += list(kwinit.items())
    data = {Sym(k): (np.array(v, dtype=np.float64) if isinstance(v, (list, tuple, np.ndarray)) and
                     np.asarray(v).dtype != np.int64 else v) for k, v in (data or {}).items()}
    for k, v in list(data.items()):
        if isinstance(v, np.ndarray) and v.dtype == np.bool_:
            data[k] = v.astype(np.float64)
    m = Model()
    set_init(m, pairs)
    src = parse(model) if isinstance(model, str) else model
    parse_model(m, src)
    unfold(m)
    unique_vars(m)
    categorize_vars(m)
    vhint = None
    if gradient:
        vhint = pre_calculate(m, data)
        backward_sweep(m, vhint)
    return ModelFunction(m, data, gradient, vhint)
