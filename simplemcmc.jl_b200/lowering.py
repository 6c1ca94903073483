"""Model -> device op tape.

Host-side counterpart of the changed reference subsystem `diff.jl`: where the reference turns the unfolded
statement list into Julia source (`backwardSweep!`/`derive`, /root/reference/src/parsing.jl:250-300,
/root/reference/src/diff.jl:192-220) and evals it, this module traces the model into typed SSA values
(static shapes take the place of preCalculate's `vhint`, src/parsing.jl:354-384), applies the same
derivation rules (src/diff.jl:29-178) in a reverse sweep, fuses the GLM pattern
`X*beta -> link -> logpdf` into one SMC_OP_GLM statement, and serialises the result in the layout of
include/simplemcmc_tape.h for libsimplemcmc_cuda.so.  Nothing here evaluates the model on the CPU.
"""
import struct

import numpy as np

from .jlsyntax import read_model

# ---- numbers from include/simplemcmc_tape.h (tests/test_abi.py checks they agree)
TAPE_MAGIC, TAPE_VERSION = 0x54434D53, 4
REG_IMM, REG_DATA, REG_HDR, REG_CHAIN, REG_PARAM, REG_GRAD = range(6)
F_ACC, F_TRANS_A, F_TRANS_B, F_FWD_ONLY = 1, 2, 4, 8
OP = dict(COPY=0, NEG=1, EXP=2, LOG=3, SIN=4, COS=5, ABS=6, SIGN=7, DIGAMMA=8, ZERO=9, SQRT=10,
          ADD=16, SUB=17, MUL=18, DIV=19, POW=20, MAX=21, MIN=22, GT=23, LT=24, GE=25, LE=26, EQ=27,
          LOGPDF=32, DLOGPDF=33, MATMUL=40, TRANSPOSE=41, GATHER=42, SCATTER_SET=43, SCATTER_ADD=44, GLM=48,
          AFFNORM=50)
DIST = dict(Normal=0, Uniform=1, Weibull=2, Beta=3, TDist=4, Exponential=5, Gamma=6, Cauchy=7, logNormal=8,
            Bernoulli=9, Binomial=10, Poisson=11, TestDiff=12)
DIST_ARITY = dict(Normal=3, Uniform=3, Weibull=3, Beta=3, TDist=2, Exponential=2, Gamma=3, Cauchy=3, logNormal=3,
                  Bernoulli=2, Binomial=3, Poisson=2, TestDiff=1)
# arguments with no adjoint: integer-valued ones (src/diff.jl:164-174)
DIST_NODERIV = dict(Bernoulli={1}, Binomial={0, 2}, Poisson={1})
GLM_NORMAL, GLM_BERNOULLI, GLM_BINOMIAL = 0, 1, 2

ACC_SYM = "__acc"   # src/SimpleMCMC.jl:14


class ModelError(Exception):
    """errors the reference raises while building the model function"""


class Val(object):
    __slots__ = ("vid", "shape", "kind", "imm", "slot", "name", "reg")

    def __init__(self, vid, shape, kind, imm=0.0, slot=0, name=None):
        self.vid, self.shape, self.kind, self.imm, self.slot, self.name = vid, tuple(shape), kind, imm, slot, name
        self.reg = None

    @property
    def numel(self):
        n = 1
        for d in self.shape:
            n *= d
        return n

    @property
    def is_array(self):
        return len(self.shape) > 0

    @property
    def varying(self):           # depends on the parameters (member of `pardesc`, src/parsing.jl:224-230)
        return self.kind in (REG_CHAIN, REG_PARAM, REG_GRAD)

    def __repr__(self):
        return "v%d%s%s" % (self.vid, list(self.shape), "iDHCPG"[self.kind])


class Stmt(object):
    __slots__ = ("op", "dst", "srcs", "flags", "aux0", "aux1", "aux2", "imm", "clone", "touched", "desc")

    def __init__(self, op, dst, srcs, flags=0, aux0=0, aux1=0, aux2=None, imm=0.0, desc=None):
        self.op, self.dst, self.srcs, self.flags = op, dst, list(srcs), flags
        self.aux0, self.aux1, self.aux2, self.imm = aux0, aux1, aux2, imm
        self.clone, self.touched = False, None
        self.desc = desc      # AFFNORM: AffNormDesc (terms, views, adjoint outputs)

    @property
    def extra_outs(self):
        """registers a statement writes besides dst: G of a GLM statement, the adjoint scalars of an AFFNORM statement"""
        if self.op == OP["GLM"]:
            return [self.aux2]
        if self.op == OP["AFFNORM"]:
            return [v for v in self.desc.outs if v is not None]
        return []

    def copy(self):
        st = Stmt(self.op, self.dst, self.srcs, flags=self.flags, aux0=self.aux0, aux1=self.aux1, aux2=self.aux2, imm=self.imm,
                  desc=self.desc)
        return st

    def __repr__(self):
        name = ([k for k, v in OP.items() if v == self.op] + ["FUSED"])[0]
        return "%r %s %s(%s)%s" % (self.dst, "+=" if self.flags & F_ACC else "=", name,
                                  ", ".join(repr(s) for s in self.srcs),
                                  "" if not (self.aux0 or self.aux1) else " aux=%s,%s" % (self.aux0, self.aux1))


class AffNormDesc(object):
    """SMC_OP_AFFNORM:  L = sum_e logpdfNormal(mu, sigma, r_e),  r_e = (((+-t_0) +- t_1) +- ...)  evaluated left to right with
    one rounding per operation exactly as the reference's element loop does, t_j = D_j[e] * s_j (kind 0) or s_j (kind 1);
    D_j = a parameter-independent array or a contiguous view (base array, offset) of one; s_j, mu, sigma scalars.
    Statement sources = [mu, sigma, s_0, ..., s_{T-1}];  outs[k] = dL/d(source k) (a per-chain scalar) or None."""

    def __init__(self, n, terms):
        self.n, self.terms = n, terms          # terms: list of (kind, sign, base Val or None, offset)
        self.outs = [None] * (2 + len(terms))


class Param(object):
    """MCMCParams (src/parsing.jl:49-53): symbol, size, position in the parameter vector"""

    def __init__(self, sym, size, start, count):
        self.sym, self.size, self.start, self.count = sym, list(size), start, count

    @property
    def map(self):
        return slice(self.start, self.start + self.count)


_ELEMENTWISE_UNARY = {"exp": "EXP", "log": "LOG", "sin": "SIN", "cos": "COS", "abs": "ABS", "sqrt": "SQRT"}


class Lowering(object):
    def __init__(self, model, init, data, gradient=True, ref_adjoint="assign", fuse=True, fuse_statements=True,
                 sufficient_stats=False, affnorm=None):
        self.vals, self.fwd, self.bwd = [], [], []
        self.env, self.data_slots, self.idx_tables = {}, [], []
        self.imm_cache = {}
        self.gradient, self.ref_adjoint, self.fuse = gradient, ref_adjoint, fuse
        self.fuse_statements = fuse_statements
        self.sufficient_stats = sufficient_stats    # opt-in: Normal-identity GLM statements evaluate from X'X, X'y, y'y
        self.pars, self.bsize, self.init = [], 0, []
        self.data = data
        self._set_init(init)
        stmts = read_model(model) if isinstance(model, str) else model
        if len(stmts) < 1:
            raise ModelError("model should contain be a block with at least 1 statement")
        self.acc = self.imm(0.0)
        self._trace(stmts)
        if self.acc.kind == REG_PARAM:       # the objective is a bare parameter: give it a computed register
            self.acc = self.emit("COPY", [self.acc], self.acc.shape)
        self.result = self.acc
        if not (self.result.varying and self.result.kind == REG_CHAIN):
            raise ModelError("Model parameters do not seem to influence model outcome")
        if self.result.is_array:
            raise ModelError("Model outcome should be a scalar, Array found")
        self._check_params_used()
        self.n_fused = 0
        self.n_affnorm = 0
        if affnorm is None:
            import os
            affnorm = os.environ.get("SMC_NO_AFFNORM") != "1"
        if fuse:
            self._fuse_glm()
            if affnorm:
                self._fuse_affnorm()
        self.n_fwd = None
        if gradient:
            self._reverse_sweep()
        self._dce()

    # ------------------------------------------------------------------ values
    def new(self, shape, kind, **kw):
        v = Val(len(self.vals), shape, kind, **kw)
        self.vals.append(v)
        return v

    def imm(self, x):
        x = float(x)
        key = repr(x)
        if key not in self.imm_cache:
            self.imm_cache[key] = self.new((), REG_IMM, imm=x)
        return self.imm_cache[key]

    def _set_init(self, init):
        """setInit!, src/parsing.jl:303-334"""
        if len(init) < 1:
            raise ModelError("There should be at leat one parameter specified, none found")
        for name, val in init:
            if isinstance(val, (bool, np.bool_)):
                raise ModelError("[setInit] forbidden parameter type for %s" % name)
            if isinstance(val, (int, float, np.integer, np.floating)):
                shape, flat = (), [float(val)]
            else:
                arr = np.asarray(val, dtype=np.float64)
                if arr.ndim not in (1, 2):
                    raise ModelError("[setInit] forbidden parameter type for %s" % name)
                shape, flat = arr.shape, arr.flatten(order="F").tolist()
            p = Param(name, shape, self.bsize, len(flat))
            self.pars.append(p)
            self.env[name] = self.new(shape, REG_PARAM, slot=self.bsize, name=name)
            self.bsize += len(flat)
            self.init.extend(flat)

    def data_val(self, name):
        if name not in self.data:
            raise ModelError("variable %s is not defined (not a parameter, not set by the model, not in `data`)" % name)
        raw = self.data[name]
        if isinstance(raw, (int, float, np.integer, np.floating, bool, np.bool_)):
            v = self.imm(float(raw))
            self.env[name] = v
            return v
        arr = np.asarray(raw)
        if arr.ndim not in (1, 2):
            raise ModelError("external variable %s must be a scalar, a vector or a matrix" % name)
        v = self.new(arr.shape, REG_DATA, slot=len(self.data_slots), name=name)
        self.data_slots.append((name, np.asarray(arr, dtype=np.float64)))
        self.env[name] = v
        return v

    def emit(self, op, srcs, shape, back=False, **kw):
        kind = REG_CHAIN if any(s.varying for s in srcs) else REG_HDR
        dst = self.new(shape, kind)
        (self.bwd if back else self.fwd).append(Stmt(OP[op], dst, srcs, **kw))
        return dst

    # ------------------------------------------------------------------ forward trace
    def _trace(self, stmts):
        tilde = any(s[0] == "tilde" for s in stmts)
        for k, st in enumerate(stmts):
            if st[0] == "assign":
                op, lhs, rhs = st[1], st[2], st[3]
                if lhs[0] != "var":
                    raise ModelError("[unfold] not a symbol on LHS of assigment (setting elements of an array is "
                                     "not supported, src/parsing.jl:186)")
                if op != "=":      # x += y  ->  x = x + y   (src/parsing.jl:86-88)
                    rhs = ("call", op[0], [lhs, rhs])
                self.env[lhs[1]] = self.expr(rhs)
            elif st[0] == "tilde":     # x ~ D(a,b)  ->  __acc = __acc + logpdfD(a,b,x)   (src/parsing.jl:103-109)
                lp = self.logpdf(st[2], [self.expr(a) for a in st[3]] + [self.expr(st[1])])
                self.acc = self.binary("ADD", self.acc, lp)
            else:
                v = self.expr(st)
                if not tilde and k == len(stmts) - 1:   # no `~`: the last expression is the objective (:117-119)
                    self.acc = v
        return self.acc

    def lookup(self, name):
        if name == ACC_SYM:
            return self.acc
        if name in self.env:
            return self.env[name]
        return self.data_val(name)

    def expr(self, node):
        tag = node[0]
        if tag == "num":
            return self.imm(node[1])
        if tag == "var":
            return self.lookup(node[1])
        if tag == "ref":
            return self.gather(node)
        if tag == "block":
            last = None
            for s in node[1]:
                if s[0] == "assign":
                    self.env[s[2][1]] = self.expr(s[3])
                else:
                    last = self.expr(s)
            return last
        if tag == "cmp":
            parts = node[1]
            res = None
            for i in range(1, len(parts), 2):
                opname = {"<": "LT", ".<": "LT", ">": "GT", ".>": "GT", "<=": "LE", ".<=": "LE", ">=": "GE",
                          ".>=": "GE", "==": "EQ"}[parts[i]]
                r = self.binary(opname, self.expr(parts[i - 1]), self.expr(parts[i + 1]))
                res = r if res is None else self.binary("MUL", res, r)
            return res
        if tag == "call":
            return self.call(node[1], node[2])
        raise ModelError("[unfold] unmanaged expr type %s" % tag)

    def call(self, fn, argnodes):
        if fn in ("+", "*") and len(argnodes) > 2:     # right-nest n-ary calls, src/parsing.jl:160-166
            tail = ("call", fn, argnodes[-2:])
            return self.call(fn, list(argnodes[:-2]) + [tail])
        args = [self.expr(a) for a in argnodes]
        if fn == "+":
            return args[0] if len(args) == 1 else self.binary("ADD", *args)
        if fn == "-":
            return self.unary("NEG", args[0]) if len(args) == 1 else self.binary("SUB", *args)
        if fn == ".*":
            return self.binary("MUL", *args)
        if fn == "./":
            return self.binary("DIV", *args)
        if fn == ".^":
            return self.binary("POW", *args)
        if fn == "^":
            if args[0].is_array or args[1].is_array:
                raise ModelError("[derive] Doesn't know how to derive ^ with array arguments (use .^)")
            return self.binary("POW", *args)
        if fn == "/":
            if args[0].is_array and args[1].is_array:
                raise ModelError("/ needs at least one scalar argument (use ./)")
            return self.binary("DIV", *args)
        if fn == "*":
            return self.times(*args)
        if fn in ("max", "min"):
            return self.binary(fn.upper(), *args)
        if fn in _ELEMENTWISE_UNARY:
            return self.unary(_ELEMENTWISE_UNARY[fn], args[0])
        if fn == "sum":
            if len(args) != 1:
                raise ModelError("sum takes one argument")
            return self.sum(args[0])
        if fn == "dot":
            a, b = args
            if a.numel != b.numel:
                raise ModelError("dot: arguments of different lengths")
            if not a.is_array and not b.is_array:
                return self.binary("MUL", a, b)
            return self.emit("MUL", [a, b], ())
        if fn == "transpose":
            return self.transpose(args[0])
        if fn.startswith("logpdf") and fn[6:] in DIST:
            return self.logpdf(fn[6:], args)
        if fn == "vcat":
            raise ModelError("[derive] Doesn't know how to derive vcat")
        raise ModelError("[derive] Doesn't know how to derive %s (unknown function)" % fn)

    @staticmethod
    def _bshape(a, b):
        if not a.is_array:
            return b.shape
        if not b.is_array:
            return a.shape
        if a.shape == b.shape:
            return a.shape
        if a.numel == b.numel:     # trailing singleton dimensions (promote_shape)
            return a.shape if len(a.shape) >= len(b.shape) else b.shape
        raise ModelError("dimension mismatch: %s vs %s" % (list(a.shape), list(b.shape)))

    def binary(self, op, a, b, back=False):
        """one 3-address statement -- after the identities the reference lists as TODO ("+0, *1", constant folding:
        TODO.md:4-9): x+0, 0+x, x-0, x*1, 1*x, x/1 are x itself and literal (op) literal is a literal.  All of them are
        exact in IEEE arithmetic, so values are unchanged; they only remove statements from the tape."""
        shape = self._bshape(a, b)
        ai, bi = a.kind == REG_IMM, b.kind == REG_IMM
        if ai and bi and op in ("ADD", "SUB", "MUL", "DIV") and not (op == "DIV" and b.imm == 0.0):
            x, y = a.imm, b.imm
            return self.imm(x + y if op == "ADD" else x - y if op == "SUB" else x * y if op == "MUL" else x / y)
        if op == "ADD" and ai and a.imm == 0.0 and b.shape == shape:
            return b
        if op in ("ADD", "SUB") and bi and b.imm == 0.0 and a.shape == shape:
            return a
        if op == "MUL" and ai and a.imm == 1.0 and b.shape == shape:
            return b
        if op in ("MUL", "DIV") and bi and b.imm == 1.0 and a.shape == shape:
            return a
        return self.emit(op, [a, b], shape, back=back)

    def unary(self, op, a, back=False):
        return self.emit(op, [a], a.shape, back=back)

    def sum(self, a, back=False):
        if not a.is_array:
            return a                      # sum(x::Real) = x, src/SimpleMCMC.jl:7
        return self.emit("COPY", [a], (), back=back)

    @staticmethod
    def _as_matrix(v):
        return v.shape if len(v.shape) == 2 else (v.shape[0], 1)

    def times(self, a, b, back=False):
        if not a.is_array or not b.is_array:
            return self.binary("MUL", a, b, back=back)
        (r, k), (k2, n) = self._as_matrix(a), self._as_matrix(b)
        if k != k2:
            raise ModelError("matrix product dimension mismatch: %s * %s" % (list(a.shape), list(b.shape)))
        shape = (r,) if len(b.shape) == 1 else (r, n)
        return self.emit("MATMUL", [a, b], shape, back=back, aux0=r, aux1=k, imm=float(n))

    def transpose(self, a, back=False):
        if not a.is_array:
            return a
        if len(a.shape) == 1:
            return self.emit("COPY", [a], (1, a.shape[0]), back=back)
        r, c = a.shape
        if r == 1 or c == 1:
            return self.emit("COPY", [a], (c, r), back=back)
        return self.emit("TRANSPOSE", [a], (c, r), back=back, aux0=r, aux1=c)

    def logpdf(self, dist, args):
        if len(args) != DIST_ARITY[dist]:
            raise ModelError("logpdf%s takes %d arguments" % (dist, DIST_ARITY[dist]))
        arrays = [a for a in args if a.is_array]
        n = arrays[0].numel if arrays else 1      # length of the first array argument, src/distribs.jl:99
        for a in arrays:
            if a.numel != n:
                raise ModelError("logpdf%s: array arguments of different lengths" % dist)
        if self.sufficient_stats and dist == "Normal" and n >= 64:
            lp = self._normal_affine(args, n)
            if lp is not None:
                return lp
        return self.emit("LOGPDF", args, (), aux0=DIST[dist], aux1=n)

    # ---- opt-in sufficient statistics for a Normal density whose variate is affine in data arrays
    def _producer(self, v):
        for st in self.fwd:
            if st.dst is v:
                return st
        return None

    def _affine(self, v, depth=0):
        """v (an array value) as sum_k coeff_k * D_k with D_k parameter-independent arrays (None = the all-ones array) and
        coeff_k scalar values; -> list of (D_k, coeff_k) or None when v is not of that form"""
        if depth > 32:
            return None
        if not v.is_array:
            return [(None, v)]                                   # a scalar broadcast over the index space
        if not v.varying:
            return [(v, self.imm(1.0))]                          # data / header array
        st = self._producer(v)
        if st is None or st.flags:
            return None
        O = OP
        if st.op in (O["ADD"], O["SUB"]):
            a, b = self._affine(st.srcs[0], depth + 1), self._affine(st.srcs[1], depth + 1)
            if a is None or b is None:
                return None
            if st.op == O["SUB"]:
                b = [(d, self.unary("NEG", c) if c.kind != REG_IMM else self.imm(-c.imm)) for d, c in b]
            return a + b
        if st.op == O["NEG"]:
            a = self._affine(st.srcs[0], depth + 1)
            return None if a is None else [(d, self.unary("NEG", c) if c.kind != REG_IMM else self.imm(-c.imm)) for d, c in a]
        if st.op in (O["MUL"], O["DIV"]) and st.dst.numel == v.numel:
            x, y = st.srcs
            if st.op == O["MUL"] and not x.is_array:
                x, y = y, x                                       # array * scalar
            if x.is_array and not y.is_array:
                a = self._affine(x, depth + 1)
                return None if a is None else [(d, self.binary("MUL" if st.op == O["MUL"] else "DIV", c, y)) for d, c in a]
        return None

    def _normal_affine(self, args, n):
        """sum_e logpdfNormal(mu, sigma, r_e) with scalar mu, sigma and r = sum_k c_k D_k  ==
        -(c' G c) / (2 sigma^2) - n log(sigma) - n log(sqrt(2 pi)),  G_kl = dot(D_k, D_l) evaluated once in the
        let-header.  A different arithmetic from the reference's element loop (src/distribs.jl:97-105): opt-in only."""
        mu, sigma, x = args
        if mu.is_array or sigma.is_array or not x.is_array or not x.varying:
            return None
        nfwd = len(self.fwd)
        terms = self._affine(x)
        if terms is not None:
            terms = terms + [(None, self.unary("NEG", mu) if mu.kind != REG_IMM else self.imm(-mu.imm))]
            merged = {}
            for d, c in terms:                                    # one coefficient per distinct array
                key = None if d is None else d.vid
                merged[key] = (d, c) if key not in merged else (d, self.binary("ADD", merged[key][1], c))
            terms = [t for t in merged.values() if not (t[1].kind == REG_IMM and t[1].imm == 0.0)]
        if terms is None or not (1 <= len(terms) <= 8) or any(d is not None and d.numel != n for d, _ in terms):
            del self.fwd[nfwd:]                                   # drop the scalar helpers emitted while probing
            return None

        def gram(a, b):
            if a is None and b is None:
                return self.imm(float(n))
            if a is None or b is None:
                return self.emit("COPY", [b if a is None else a], ())     # sum(D): parameter independent -> header
            return self.emit("MUL", [a, b], ())                            # dot(D_k, D_l)
        q = None
        for i, (di, ci) in enumerate(terms):
            for j in range(i, len(terms)):
                dj, cj = terms[j]
                t = self.binary("MUL", self.binary("MUL", ci, cj), gram(di, dj))
                if j > i:
                    t = self.binary("MUL", self.imm(2.0), t)
                q = t if q is None else self.binary("ADD", q, t)
        two_s2 = self.binary("MUL", self.imm(2.0), self.binary("MUL", sigma, sigma))
        lp = self.binary("SUB", self.unary("NEG", self.binary("DIV", q, two_s2)), self.binary("MUL", self.imm(float(n)), self.unary("LOG", sigma)))
        return self.binary("SUB", lp, self.imm(float(n) * 0.91893853320467274178))

    # ---- refs: static index lists (column major, 0 based)
    def _index_list(self, node, dimlen):
        def const(e):
            if e[0] == "num":
                return e[1]
            if e[0] == "end":
                return dimlen
            if e[0] == "call" and e[1] in ("+", "-", "*") and len(e[2]) == 2:
                a, b = const(e[2][0]), const(e[2][1])
                return a + b if e[1] == "+" else (a - b if e[1] == "-" else a * b)
            if e[0] == "call" and e[1] == "-" and len(e[2]) == 1:
                return -const(e[2][0])
            if e[0] == "var" and e[1] in self.data and np.ndim(self.data[e[1]]) == 0:
                return self.data[e[1]]
            raise ModelError("index expressions must be constants")
        if node[0] == "colon":
            return np.arange(dimlen), False
        if node[0] == "range":
            return np.arange(int(const(node[1])) - 1, int(const(node[2]))), False
        if node[0] == "var" and node[1] in self.data and np.ndim(self.data[node[1]]) >= 1:
            v = np.asarray(self.data[node[1]])
            if v.dtype == np.bool_:
                return np.nonzero(v.flatten(order="F"))[0], False
            return v.flatten(order="F").astype(np.int64) - 1, False
        return np.array([int(const(node)) - 1]), True

    def gather(self, node):
        base = self.expr(node[1])
        idx = node[2]
        if not base.is_array:
            raise ModelError("ref on a scalar")
        if len(idx) == 1:
            lin, scalar = self._index_list(idx[0], base.numel)
            shape = () if scalar else (len(lin),)
        elif len(idx) == 2 and len(base.shape) == 2:
            r, rs = self._index_list(idx[0], base.shape[0])
            c, cs = self._index_list(idx[1], base.shape[1])
            lin = (r[:, None] + c[None, :] * base.shape[0]).flatten(order="F")
            shape = () if (rs and cs) else ((len(r),) if cs else (len(r), len(c)))
        else:
            raise ModelError("unsupported ref")
        if lin.size and (lin.min() < 0 or lin.max() >= base.numel):
            raise ModelError("BoundsError in ref")
        self.idx_tables.append(lin.astype(np.int32))
        return self.emit("GATHER", [base], shape, aux0=len(self.idx_tables) - 1)

    # ------------------------------------------------------------------ analysis
    def _check_params_used(self):
        """every parameter must influence the outcome (categorizeVars!, src/parsing.jl:242-245)"""
        needed = {self.result.vid}
        for st in reversed(self.fwd):
            if st.dst.vid in needed:
                needed.update(s.vid for s in st.srcs)
        unused = [p.sym for p in self.pars if self.env_param(p).vid not in needed]
        if unused:
            raise ModelError("Model parameter(s) %s do not seem to influence model outcome" % unused)
        self.needed = needed

    def env_param(self, p):
        for v in self.vals:
            if v.kind == REG_PARAM and v.name == p.sym:
                return v
        raise KeyError(p.sym)

    # ------------------------------------------------------------------ GLM fusion
    def _consumers(self, v):
        return [st for st in self.fwd if any(s is v for s in st.srcs)]

    def _fuse_glm(self):
        for mm in list(self.fwd):
            if mm.op != OP["MATMUL"] or mm.flags:
                continue
            X, b = mm.srcs
            if X.varying or len(X.shape) != 2 or b.kind != REG_PARAM or len(b.shape) != 1:
                continue
            N, M = X.shape
            if M > 64 or N < 64:
                continue
            plan = self._match_glm(mm.dst, N)
            if plan is None:
                continue
            family, sign, par, y, lp_stmt, dead = plan
            if self.sufficient_stats and family == GLM_NORMAL:
                sign |= 4       # SMC_GLM_F_GRAM (include/simplemcmc_tape.h)
            G = self.new((M,), REG_CHAIN)
            glm = Stmt(OP["GLM"], lp_stmt.dst, [X, b, y], aux0=family, aux1=sign, aux2=G, imm=par)
            pos = self.fwd.index(lp_stmt)
            self.fwd[pos] = glm
            for st in dead + [mm]:
                self.fwd.remove(st)
            self.n_fused += 1

    def _single_use(self, v):
        c = self._consumers(v)
        return c[0] if len(c) == 1 and v is not self.result else None

    def _match_glm(self, eta, N):
        """returns (family, sign bits, parameter, y, logpdf statement, statements made dead) or None"""
        def is_row_data(v):
            return (not v.varying) and (v.numel == N or not v.is_array)

        def is_const(v, x=None):
            return v.kind == REG_IMM and (x is None or v.imm == x)

        st = self._single_use(eta)
        if st is None:
            return None
        dead = []
        # ---- Normal, identity link
        if st.op == OP["SUB"]:
            a, b = st.srcs
            other = b if a is eta else a
            if not is_row_data(other) or other is eta:
                return None
            sa_neg, sb_neg = (0, 1) if a is eta else (1, 0)      # z = eta - y | y - eta
            lp = self._single_use(st.dst)
            if lp is None or lp.op != OP["LOGPDF"] or lp.aux0 != DIST["Normal"]:
                return None
            mu, sigma, x = lp.srcs
            if x is not st.dst or not is_const(mu, 0.0) or not is_const(sigma) or sigma.imm <= 0:
                return None
            return GLM_NORMAL, sa_neg | (sb_neg << 1), sigma.imm, other, lp, [st]
        if st.op == OP["LOGPDF"] and st.aux0 == DIST["Normal"]:
            mu, sigma, x = st.srcs
            if not is_const(sigma) or sigma.imm <= 0:
                return None
            if mu is eta and x is not eta and is_row_data(x):       # z = x - eta
                return GLM_NORMAL, 1, sigma.imm, x, st, []
            if x is eta and mu is not eta and is_row_data(mu):      # z = eta - mu
                return GLM_NORMAL, 2, sigma.imm, mu, st, []
            return None
        # ---- Bernoulli, prob = 1/(1+exp(+-eta))
        sign = 0
        if st.op == OP["NEG"]:
            dead.append(st)
            sign = 1
            st = self._single_use(st.dst)
            if st is None:
                return None
        if st.op != OP["EXP"]:
            return None
        dead.append(st)
        add = self._single_use(st.dst)
        if add is None or add.op != OP["ADD"]:
            return None
        o = add.srcs[1] if add.srcs[0] is st.dst else add.srcs[0]
        if not is_const(o, 1.0):
            return None
        dead.append(add)
        div = self._single_use(add.dst)
        if div is None or div.op != OP["DIV"] or div.srcs[1] is not add.dst or not is_const(div.srcs[0], 1.0):
            return None
        dead.append(div)
        lp = self._single_use(div.dst)
        if lp is not None and lp.op == OP["LOGPDF"] and lp.aux0 == DIST["Binomial"] and lp.srcs[1] is div.dst:
            # counts: logpdfBinomial(n, prob, Y) with a literal number of trials (src/distribs.jl:43, adjoint src/diff.jl:169-170)
            ntr, _, y = lp.srcs
            if not is_const(ntr) or ntr.imm < 0 or ntr.imm != round(ntr.imm) or y.kind != REG_DATA or y.numel != N:
                return None
            yv = self.data_slots[y.slot][1]
            if not np.all((yv >= 0) & (yv <= ntr.imm) & (yv == np.round(yv))):     # outside the support: leave it to the generic path
                return None
            return GLM_BINOMIAL, sign, ntr.imm, y, lp, dead
        if lp is None or lp.op != OP["LOGPDF"] or lp.aux0 != DIST["Bernoulli"] or lp.srcs[0] is not div.dst:
            return None
        y = lp.srcs[1]
        if y.kind != REG_DATA or y.numel != N:
            return None
        yv = self.data_slots[y.slot][1]
        if not np.all((yv == 0.0) | (yv == 1.0)):    # the fused kernel assumes a 0/1 response (src/distribs.jl:18-19 errors otherwise)
            return None
        return GLM_BERNOULLI, sign, 0.0, y, lp, dead

    # ------------------------------------------------------------------ affine residual into a Normal density
    AFFNORM_MIN_N = 2048

    def _fuse_affnorm(self):
        """`resid = D0 - D1 * s1 - s2; resid ~ Normal(mu, sigma)` (examples/ornstein.jl:24-26 and every model of that shape:
        a variate built from parameter-independent arrays and per-chain scalars by + and -) becomes one AFFNORM statement:
        one hand-written kernel reads the arrays once, keeps r_e in registers and returns the log-density together with
        the sums the reverse sweep needs (src/diff.jl:94-99 for the Normal adjoints, :29-45 for + - *), instead of
        three element-wise temporaries forward and five backward."""
        for lp in list(self.fwd):
            if lp.op != OP["LOGPDF"] or lp.aux0 != DIST["Normal"] or lp.flags or lp not in self.fwd:
                continue
            mu, sigma, x = lp.srcs
            n = lp.aux1
            if mu.is_array or sigma.is_array or not x.is_array or x.numel != n or not x.varying or n < self.AFFNORM_MIN_N:
                continue
            if self._single_use(x) is not lp:
                continue
            dead = []
            terms = self._linearise(x, n, dead, top=True)
            if terms is None or not (1 <= len(terms) <= 4) or not any(t[0] == 0 for t in terms):
                continue
            desc = AffNormDesc(n, [(k, sg, base, off) for (k, sg, base, off, sc) in terms])
            srcs = [mu, sigma] + [t[4] for t in terms]
            st = Stmt(OP["AFFNORM"], lp.dst, srcs, aux1=n, desc=desc)
            self.fwd[self.fwd.index(lp)] = st
            for d in dead:
                self.fwd.remove(d)
            self.n_affnorm += 1

    def _view_of(self, v):
        """a parameter-independent array as (array to read, element offset): a contiguous range of a data vector
        (x[2:end]) is a view of that vector, so that x[2:end] and x[1:end-1] are read as ONE stream"""
        if v.kind == REG_HDR:
            st = self._producer(v)
            if st is not None and st.op == OP["GATHER"] and not st.flags and st.srcs[0].kind == REG_DATA and len(st.srcs[0].shape) == 1:
                idx = self.idx_tables[st.aux0]
                if len(idx) == v.numel and len(idx) >= 1 and np.array_equal(idx, idx[0] + np.arange(len(idx), dtype=idx.dtype)):
                    return st.srcs[0], int(idx[0])
        return v, 0

    def _linearise(self, v, n, dead, top=False):
        """v as a left-to-right sum of signed terms with the reference's roundings, or None.  IEEE addition commutes and
        negation is exact, so a caterpillar tree of + / - / unary minus (at most one non-leaf child per node) has an
        evaluation order that is a plain chain.  -> list of (kind, sign, base, offset, scalar)"""
        def leaf(u):
            if not u.is_array:
                return (1, 1.0, None, 0, u)                                  # a scalar broadcast over the index space
            if not u.varying and u.numel == n and u.kind in (REG_DATA, REG_HDR):
                base, off = self._view_of(u)
                return (0, 1.0, base, off, self.imm(1.0))
            st = self._producer(u)
            if (st is not None and st.op == OP["MUL"] and not st.flags and u.numel == n and self._single_use(u) is not None
                    and u.kind == REG_CHAIN):
                a, b = st.srcs
                if not a.is_array:
                    a, b = b, a
                if a.is_array and not a.varying and a.numel == n and a.kind in (REG_DATA, REG_HDR) and not b.is_array:
                    base, off = self._view_of(a)
                    return (0, 1.0, base, off, b), st
            return None

        def as_leaf(u):
            r = leaf(u)
            if r is None:
                return None, None
            if isinstance(r[0], tuple):
                return r[0], r[1]
            return r, None

        lf, lst = as_leaf(v)
        if lf is not None:
            if lst is not None:
                dead.append(lst)
            return [lf]
        st = self._producer(v)
        if st is None or st.flags or v.kind != REG_CHAIN or v.numel != n or (not top and self._single_use(v) is None):
            return None
        neg = lambda ts: [(k, -sg, b, o, sc) for (k, sg, b, o, sc) in ts]
        if st.op == OP["NEG"]:
            inner = self._linearise(st.srcs[0], n, dead)
            if inner is None:
                return None
            dead.append(st)
            return neg(inner)
        if st.op not in (OP["ADD"], OP["SUB"]):
            return None
        a, b = st.srcs
        la, lsa = as_leaf(a)
        lb, lsb = as_leaf(b)
        sgb = 1.0 if st.op == OP["ADD"] else -1.0
        if lb is not None:                       # chain(a) +- leaf(b)
            head = [la] if la is not None else self._linearise(a, n, dead)
            if head is None:
                return None
            if la is not None and lsa is not None:
                dead.append(lsa)
            if lsb is not None:
                dead.append(lsb)
            dead.append(st)
            return head + [(lb[0], sgb * lb[1], lb[2], lb[3], lb[4])]
        if la is not None:                       # leaf(a) +- chain(b)  ==  (+-chain(b)) + leaf(a)
            tail = self._linearise(b, n, dead)
            if tail is None:
                return None
            if lsa is not None:
                dead.append(lsa)
            dead.append(st)
            return (tail if sgb > 0 else neg(tail)) + [la]
        return None

    # ------------------------------------------------------------------ reverse sweep
    def _adjoint_target(self, v):
        if v.vid not in self.adj:
            if v.kind == REG_PARAM:
                self.adj[v.vid] = self.new(v.shape, REG_GRAD, slot=v.slot, name=v.name)
            else:
                self.adj[v.vid] = self.new(v.shape, REG_CHAIN)
            self.adj_init[v.vid] = False
        return self.adj[v.vid]

    def _own_adjoint(self, target):
        """copy on write: an adjoint that so far was only another name for a value gets its own register"""
        if target.vid in self.adj_alias:
            alias = self.adj.pop(target.vid)
            self.adj_alias.discard(target.vid)
            dst = self._adjoint_target(target)
            self.bwd.append(Stmt(OP["COPY"], dst, [alias]))
            self.adj_init[target.vid] = True

    def contrib(self, target, op, srcs, **kw):
        """adjoint(target) += op(srcs)   -- the `dv = dv + ...` statement of src/diff.jl:214.
        The seed d(finalacc) = 1.0 (src/parsing.jl:419) travels down the accumulator chain as a literal: `x * 1.0` is x,
        and a first contribution that is a plain same-shape copy makes the adjoint another name for its source instead
        of a statement (exact identities; they remove the all-ones per-chain registers the literal derivation creates)."""
        if op == "MUL" and not kw and len(srcs) == 2:
            a, b = srcs
            if b.kind == REG_IMM and b.imm == 1.0:
                op, srcs = "COPY", [a]
            elif a.kind == REG_IMM and a.imm == 1.0:
                op, srcs = "COPY", [b]
        if (op == "COPY" and not kw and target.kind != REG_PARAM and target.vid not in self.adj
                and srcs[0].shape == target.shape):
            self.adj[target.vid] = srcs[0]
            self.adj_init[target.vid] = True
            self.adj_alias.add(target.vid)
            return
        self._own_adjoint(target)
        dst = self._adjoint_target(target)
        n = max([s.numel for s in srcs] + [1])
        if op not in ("MATMUL", "TRANSPOSE") and not (dst.numel == n or dst.numel == 1 or n == 1):
            raise ModelError("internal: adjoint shape mismatch for %r" % (target,))
        flags = kw.pop("flags", 0) | (F_ACC if self.adj_init[target.vid] else 0)
        self.bwd.append(Stmt(OP[op], dst, srcs, flags=flags, **kw))
        self.adj_init[target.vid] = True

    def b(self, op, *srcs, **kw):
        """temporary in the backward list, elementwise with broadcasting"""
        shape = ()
        for s in srcs:
            if s.is_array:
                shape = s.shape if (not shape or len(s.shape) >= len(shape)) else shape
        return self.emit(op, list(srcs), shape, back=True, **kw)

    def _reverse_sweep(self):
        self.adj, self.adj_init, self.adj_alias = {}, {}, set()
        needed = set()
        # active = descendants of parameters that are ancestors of the accumulator (src/parsing.jl:291)
        anc = {self.result.vid}
        for st in reversed(self.fwd):
            outs = [st.dst] + st.extra_outs
            if any(o.vid in anc for o in outs):
                anc.update(s.vid for s in st.srcs)
        one = self.imm(1.0)
        for st in reversed(self.fwd):
            if st.dst.vid not in anc or not st.dst.varying:
                continue
            if st.dst is self.result and self.result.vid not in self.adj:
                ds = one                                 # d(finalacc) = 1.0, src/parsing.jl:419
            elif st.dst.vid in self.adj:
                ds = self.adj[st.dst.vid]
            else:
                continue
            for k, a in enumerate(st.srcs):
                if a.varying and a.vid in anc:
                    self._rule(st, k, a, ds)
        for p in self.pars:
            pv = self.env_param(p)
            if pv.vid not in self.adj:
                raise ModelError("Model parameter(s) %s do not seem to influence model outcome" % [p.sym])

    def _rule(self, st, k, x, ds):
        """adjoint of argument k of statement st (rules: src/diff.jl:29-178)"""
        op = st.op
        s = st.srcs
        O = OP
        if op == O["ADD"]:
            self.contrib(x, "COPY", [ds])
        elif op == O["SUB"]:
            self.contrib(x, "COPY" if k == 0 else "NEG", [ds])
        elif op == O["NEG"]:
            self.contrib(x, "NEG", [ds])
        elif op == O["COPY"]:                 # sum(x), transpose of a vector: adjoint flows through unchanged
            self.contrib(x, "COPY", [ds])
        elif op == O["LOG"]:
            self.contrib(x, "DIV", [ds, x])
        elif op == O["EXP"]:
            self.contrib(x, "MUL", [st.dst, ds])
        elif op == O["SIN"]:
            self.contrib(x, "MUL", [self.b("COS", x), ds])
        elif op == O["COS"]:
            self.contrib(x, "MUL", [self.b("NEG", self.b("SIN", x)), ds])
        elif op == O["ABS"]:
            self.contrib(x, "MUL", [self.b("SIGN", x), ds])
        elif op == O["SQRT"]:
            self.contrib(x, "DIV", [ds, self.b("MUL", self.imm(2.0), st.dst)])
        elif op == O["MUL"]:
            other = s[1 - k]
            if st.dst.numel == 1 and (s[0].is_array or s[1].is_array):    # dot(x, y)
                self.contrib(x, "MUL", [other, ds])
            else:
                self.contrib(x, "MUL", [ds, other])
        elif op == O["DIV"]:
            if k == 0:
                self.contrib(x, "DIV", [ds, s[1]])
            else:
                t = self.b("NEG", self.b("DIV", s[0], self.b("MUL", s[1], s[1])))
                self.contrib(x, "MUL", [t, ds])
        elif op == O["POW"]:
            if k == 0:
                t = self.b("MUL", s[1], self.b("POW", s[0], self.b("SUB", s[1], self.imm(1.0))))
                self.contrib(x, "MUL", [t, ds])
            else:
                t = self.b("MUL", self.b("LOG", s[0]), st.dst)
                self.contrib(x, "MUL", [t, ds])
        elif op in (O["MAX"], O["MIN"]):
            first = (op == O["MAX"]) == (k == 0)
            self.contrib(x, "MUL", [self.b("GT" if first else "LT", s[0], s[1]), ds])
        elif op in (O["GT"], O["LT"], O["GE"], O["LE"], O["EQ"], O["SIGN"], O["ZERO"]):
            raise ModelError("[derive] Doesn't know how to derive a comparison by argument %d" % (k + 1))
        elif op == O["MATMUL"]:
            r, kk, n = st.aux0, st.aux1, int(st.imm)
            if k == 0:      # ds * transpose(y)
                self.contrib(x, "MATMUL", [ds, s[1]], flags=F_TRANS_B, aux0=r, aux1=n, imm=float(kk))
            else:           # transpose(x) * ds
                self.contrib(x, "MATMUL", [s[0], ds], flags=F_TRANS_A, aux0=kk, aux1=r, imm=float(n))
        elif op == O["TRANSPOSE"]:
            self.contrib(x, "TRANSPOSE", [ds], aux0=st.aux1, aux1=st.aux0)
        elif op == O["GATHER"]:
            self._own_adjoint(x)
            if x.vid not in self.adj or not self.adj_init[x.vid]:
                self.contrib(x, "ZERO", [x])
            dst = self.adj[x.vid]
            mode = "SCATTER_SET" if self.ref_adjoint == "assign" else "SCATTER_ADD"
            self.bwd.append(Stmt(OP[mode], dst, [ds], aux0=st.aux0))
        elif op == O["LOGPDF"]:
            dist = [d for d, i in DIST.items() if i == st.aux0][0]
            if k in DIST_NODERIV.get(dist, ()):
                raise ModelError("[derive] Doesn't know how to derive logpdf%s by argument %d (integer argument)"
                                 % (dist, k + 1))
            if dist == "Uniform" and k == 2:
                self.contrib(x, "ZERO", [x])
                return
            t = self.emit("DLOGPDF", list(s), self._arg_shape(s), back=True, aux0=st.aux0, aux1=k)
            self.contrib(x, "MUL", [t, ds])
        elif op == O["GLM"]:
            if k != 1:
                raise ModelError("internal: GLM adjoint requested for a non-parameter argument")
            self.contrib(x, "MUL", [st.aux2, ds])
        elif op == O["AFFNORM"]:
            # dL/d(source k) comes out of the kernel as a per-chain scalar (Normal adjoints src/diff.jl:94-99 pushed through
            # the + - * rules :29-45 inside the kernel's finish step)
            if st.desc.outs[k] is None:
                st.desc.outs[k] = self.new((), REG_CHAIN)
            self.contrib(x, "MUL", [st.desc.outs[k], ds])
        else:
            raise ModelError("[derive] Doesn't know how to derive op %d" % op)

    @staticmethod
    def _arg_shape(srcs):
        for a in srcs:
            if a.is_array:
                return a.shape
        return ()

    # ------------------------------------------------------------------ clean-up and serialisation
    def _dce(self):
        live = {self.result.vid}
        for v in self.vals:
            if v.kind == REG_GRAD:
                live.add(v.vid)
        allst = self.fwd + self.bwd
        keep = [False] * len(allst)
        for i in range(len(allst) - 1, -1, -1):
            st = allst[i]
            outs = [st.dst] + st.extra_outs
            if any(o.vid in live for o in outs):
                keep[i] = True
                live.update(s.vid for s in st.srcs)
        nf = len(self.fwd)
        self.fwd = [st for i, st in enumerate(self.fwd) if keep[i]]
        self.bwd = [st for i, st in enumerate(self.bwd) if keep[nf + i]]

    def statements(self):
        return list(self.fwd), list(self.bwd)

    def serialise(self):
        fwd, bwd, programs = self.fwd, self.bwd, []
        self.tape_result = self.result
        if self.fuse_statements:
            from . import fusion
            fwd, bwd, programs = fusion.fuse(self)
        self.tape_fwd, self.tape_bwd, self.programs = fwd, bwd, programs
        stmts = fwd + bwd
        used = []
        seen = set()

        def use(v):
            if v.vid not in seen:
                seen.add(v.vid)
                used.append(v)
        for st in stmts:
            for v in st.srcs:
                use(v)
            use(st.dst)
            if st.op == OP["GLM"]:
                use(st.aux2)
            if st.op == OP["AFFNORM"]:
                for v in st.extra_outs:
                    use(v)
                for (_, _, base, _) in st.desc.terms:
                    if base is not None:
                        use(base)
            if st.touched:
                for v in st.touched:
                    use(v)
        use(self.tape_result)
        data_used = sorted({v.slot for v in used if v.kind == REG_DATA})
        slot_map = {s: i for i, s in enumerate(data_used)}
        for i, v in enumerate(used):
            v.reg = i
        # AFFNORM descriptors travel as index tables (register numbers are known only now):
        #   [T, reg(dL/dmu) or -1, reg(dL/dsigma) or -1,  then per term: kind, sign, array reg or -1, offset, scalar reg, reg(dL/ds) or -1]
        # (index tables no surviving statement refers to -- e.g. the ranges of x[2:end] once AFFNORM reads x as views -- stay home)
        idx_ops = (OP["GATHER"], OP["SCATTER_SET"], OP["SCATTER_ADD"])
        idx_used = sorted({st.aux0 for st in stmts if st.op in idx_ops})
        idx_map = {old: new for new, old in enumerate(idx_used)}
        tables = [self.idx_tables[i] for i in idx_used]
        aff_table = {}
        for st in stmts:
            if st.op == OP["AFFNORM"]:
                d = st.desc
                row = [len(d.terms)] + [(-1 if o is None else o.reg) for o in d.outs[:2]]
                for j, (kind, sg, base, off) in enumerate(d.terms):
                    o = d.outs[2 + j]
                    row += [kind, int(sg), -1 if base is None else base.reg, off, st.srcs[2 + j].reg, -1 if o is None else o.reg]
                aff_table[id(st)] = len(tables)
                tables.append(np.array(row, dtype=np.int32))
        out = [struct.pack("<12I", TAPE_MAGIC, TAPE_VERSION, len(used), len(stmts), self.bsize, len(data_used),
                           len(tables), self.tape_result.reg, 0, len(fwd), len(programs), 0)]
        for v in used:
            if len(v.shape) == 0:
                rows, cols = 1, 1
            elif len(v.shape) == 1:
                rows, cols = v.shape[0], 1
            else:
                rows, cols = v.shape
            aux = slot_map[v.slot] if v.kind == REG_DATA else v.slot
            out.append(struct.pack("<4Iqd", v.kind, rows, cols, len(v.shape), aux, v.imm))
        for st in stmts:
            srcs, aux0 = st.srcs, st.aux0
            if st.op in idx_ops:
                aux0 = idx_map[st.aux0]
            if st.op == OP["AFFNORM"]:
                srcs, aux0 = st.srcs[:2], aff_table[id(st)]          # (mu, sigma); everything else is in the descriptor
            src = [s.reg for s in srcs] + [0] * (4 - len(srcs))
            aux2 = st.aux2.reg if st.aux2 is not None else 0
            out.append(struct.pack("<4I4I4Id", st.op, st.flags, st.dst.reg, len(srcs), *src, aux0, st.aux1, aux2, 0,
                                   float(st.imm)))
        bound = []
        for s in data_used:
            name, arr = self.data_slots[s]
            rows, cols = (arr.shape[0], 1) if arr.ndim == 1 else arr.shape
            out.append(struct.pack("<56s2I", name.encode()[:55], rows, cols))
            bound.append((name, arr))
        for t in tables:
            pad = t if len(t) % 2 == 0 else np.concatenate([t, np.zeros(1, np.int32)])
            out.append(struct.pack("<2I", len(t), 0) + pad.astype("<i4").tobytes())
        regmap = {v.vid: v.reg for v in used}
        for prog in programs:
            out.append(prog.pack(regmap))
        return b"".join(out), bound
