"""Seeded random models in the reference's DSL (multi-statement graphs with re-used intermediates, mixed scalar / vector /
matrix operands, slices, gathers, reductions and several `~` statements), for parity runs of the device path against the
oracle beyond the one-operation cases of rule_cases.py (test/test.jl:142-224 only ever differentiates a single call).

Every operator used here has a rule in src/diff.jl:29-178.  Values are kept in tame ranges by construction (bounded
arguments for exp, shifted absolute values for log, ./ and .^), so that the 1e-12 parity bar is meaningful.
"""
import re

import numpy as np

N, M = 23, 5


def make_data(rng, n=N, m=M):
    return dict(
        X=rng.standard_normal((n, m)),
        Xt=rng.standard_normal((m, n)),
        y=rng.standard_normal(n),
        z=0.5 + rng.random(n),
        yb=(rng.random(n) < 0.5).astype(np.float64),
        ll=rng.integers(1, m + 1, size=n).astype(np.int64),
        k=1.7,
        D3=rng.standard_normal((n, 3)),
        one3=np.ones(3),
    )


def make_params(rng, n=N, m=M):
    """ordered (name, value) pairs: a, s scalars (s is used as a scale: positive), b an m-vector, w an n-vector,
    Bm an m x 3 matrix"""
    return [("a", 0.3 + 0.2 * rng.standard_normal()), ("s", 1.3 + 0.1 * rng.random()),
            ("b", 0.5 * rng.standard_normal(m)), ("w", 0.5 * rng.standard_normal(n)),
            ("Bm", 0.5 * rng.standard_normal((m, 3)))]


class _Gen(object):
    def __init__(self, rng):
        self.rng = rng
        self.lines = []
        self.pool = {"S": ["a", "s", "k"], "V": ["w", "y", "z"], "U": [], "P": ["b"], "A": ["D3"]}   # A: n x 3 matrices
        self.active = {"a", "s", "b", "w", "Bm"}      # names that depend on a parameter
        self.nt = 0

    def pick(self, kind, want_active=False):
        c = self.pool[kind]
        if want_active:
            act = [x for x in c if x in self.active]
            if act:
                c = act
        return c[int(self.rng.integers(len(c)))]

    def emit(self, kind, expr, active):
        self.nt += 1
        name = "t%d" % self.nt
        self.lines.append("%s = %s" % (name, expr))
        self.pool[kind].append(name)
        if active:
            self.active.add(name)
        return name

    def is_active(self, *names):
        return any(n in self.active for n in names)

    def step(self):
        r = self.rng
        vec = "V" if (not self.pool["U"] or r.random() < 0.7) else "U"
        choice = int(r.integers(20))
        if choice >= 16:       # two-dimensional values (examples/hierarchical.jl:31: (beta[ll,:] .* X) * oneL)
            acts = [x for x in self.pool["A"] if x in self.active]
            if choice == 16 or not acts:
                if r.random() < 0.5:
                    return self.emit("A", "X * Bm", True)
                return self.emit("A", "Bm[ll,:]", True)
            x = acts[int(r.integers(len(acts)))]
            if choice == 17:
                yv = self.pick("A")
                op = ["+", "-", ".*"][int(r.integers(3))]
                return self.emit("A", "%s %s %s" % (x, op, yv), True)
            if choice == 18:
                f = ["sin(%s)", "exp(0.2 * %s)", "%s .* %s" % ("%s", self.pick("S")), "%s ./ (1.5 + abs(D3))"][int(r.integers(4))]
                return self.emit("A", f % x, True)
            if r.random() < 0.4:
                return self.emit("S", "sum(%s)" % x, True)
            if r.random() < 0.5:
                return self.emit("V", "%s * one3" % x, True)
            return self.emit("V", "%s[:,2]" % x, True)
        if choice <= 2:        # elementwise binary, vector (op) vector / scalar
            x = self.pick(vec, True)
            ykind = vec if r.random() < 0.6 else "S"
            yv = self.pick(ykind)
            op = ["+", "-", ".*"][int(r.integers(3))]
            if r.random() < 0.5:
                x, yv = yv, x
            return self.emit(vec, "%s %s %s" % (x, op, yv), self.is_active(x, yv))
        if choice == 3:        # guarded division
            x = self.pick(vec, True)
            d = self.pick(vec if r.random() < 0.5 else "S")
            if r.random() < 0.5:
                return self.emit(vec, "%s ./ (1.5 + abs(%s))" % (x, d), self.is_active(x, d))
            return self.emit(vec, "%s ./ (1.5 + abs(%s))" % (d, x), self.is_active(x, d))
        if choice == 4:        # unary
            x = self.pick(vec, True)
            f = ["sin(%s)", "cos(%s)", "exp(0.2 * %s)", "log(1.5 + abs(%s))", "abs(%s)", "0.5 - %s"][int(r.integers(6))]
            return self.emit(vec, f % x, self.is_active(x))
        if choice == 5:        # powers
            x = self.pick(vec, True)
            if r.random() < 0.5:
                return self.emit(vec, "(1.5 + abs(%s)) .^ 1.7" % x, self.is_active(x))
            e = self.pick("S")
            return self.emit(vec, "(1.5 + abs(%s)) .^ (0.3 * %s)" % (x, e), self.is_active(x, e))
        if choice == 6:        # max / min
            x, yv = self.pick(vec, True), self.pick(vec)
            return self.emit(vec, "%s(%s, %s)" % (["max", "min"][int(r.integers(2))], x, yv), self.is_active(x, yv))
        if choice == 7:        # matrix * parameter vector
            p = self.pick("P", True)
            return self.emit("V", "X * %s" % p, self.is_active(p))
        if choice == 8:        # matrix * n-vector -> m-vector
            x = self.pick("V", True)
            return self.emit("P", "Xt * %s" % x, self.is_active(x))
        if choice == 9:        # reductions
            x = self.pick(vec, True)
            if r.random() < 0.5:
                return self.emit("S", "sum(%s)" % x, self.is_active(x))
            yv = self.pick(vec)
            return self.emit("S", "dot(%s, %s)" % (x, yv), self.is_active(x, yv))
        if choice == 10:       # scalar arithmetic
            x, yv = self.pick("S", True), self.pick("S")
            op = ["+", "-", "*"][int(r.integers(3))]
            return self.emit("S", "%s %s 0.1 * %s" % (x, op, yv) if op != "*" else "%s * %s" % (x, yv), self.is_active(x, yv))
        if choice == 11:       # scalar functions
            x = self.pick("S", True)
            f = ["sin(%s)", "exp(0.1 * %s)", "log(2. + abs(%s))", "(1.5 + abs(%s))^2", "%s / (2. + abs(k))"][int(r.integers(5))]
            return self.emit("S", f % x, self.is_active(x))
        if choice == 12:       # slices (shifted views of one vector, as in examples/ornstein.jl:23)
            x = self.pick("V", True)
            sl = ["%s[2:end]", "%s[1:end-1]"][int(r.integers(2))]
            return self.emit("U", sl % x, self.is_active(x))
        if choice == 13 and self.pool["U"]:
            x, yv = self.pick("U", True), self.pick("U")
            return self.emit("U", "%s - %s .* %s" % (x, yv, self.pick("S")), True if self.is_active(x, yv) else self.is_active(x))
        if choice == 14:       # gather through a data index (examples/hierarchical.jl:31)
            p = self.pick("P", True)
            return self.emit("V", "%s[ll]" % p, self.is_active(p))
        # parameter vector arithmetic
        p, q = self.pick("P", True), self.pick("P")
        return self.emit("P", "%s .* 0.5 + %s" % (p, q), self.is_active(p, q))

    def likelihood(self):
        r = self.rng
        choice = int(r.integers(7))
        vec = "V" if (not self.pool["U"] or r.random() < 0.75) else "U"
        x = self.pick(vec, True)
        if choice == 0:
            self.lines.append("%s ~ Normal(0, 1.5)" % x)
        elif choice == 1:
            mu = self.pick("S", True)
            self.lines.append("%s ~ Normal(%s, s)" % (x, mu))
        elif choice == 2 and vec == "V":
            self.lines.append("y ~ Normal(%s, 2.0)" % x)
        elif choice == 3 and vec == "V":
            self.nt += 1
            self.lines.append("t%d = 1 / (1. + exp(%s))" % (self.nt, x))
            self.lines.append("yb ~ Bernoulli(t%d)" % self.nt)
        elif choice == 4:
            sc = self.pick("S", True)
            self.lines.append("%s ~ Normal(0, 1.0)" % sc)
        elif choice == 5 and vec == "V":
            self.lines.append("z ~ Gamma(2.0, 1.5 + abs(%s))" % x)
        else:
            self.lines.append("%s ~ Cauchy(0, 2.0)" % x)


def random_model(seed, n_steps=None):
    """-> model text.  Parameters a, s, b, w, Bm (make_params), data X, Xt, y, z, yb, ll, k, D3, one3 (make_data)."""
    rng = np.random.default_rng(1000 + seed)
    g = _Gen(rng)
    n_steps = n_steps or int(rng.integers(4, 22))
    for i in range(n_steps):
        g.step()
        if rng.random() < 0.2:
            g.likelihood()
    for _ in range(int(rng.integers(1, 4))):
        g.likelihood()
    # A parameter-dependent assignment that never reaches a `~` makes the reference's generated function read an
    # uninitialised adjoint (backwardSweep! visits it, the init block covers only accanc ∩ pardesc: src/parsing.jl:250-300,
    # :417-430), so dead intermediates are pruned: the fuzz is about the live graph.
    lines = list(g.lines)
    while True:
        used = set()
        for ln in lines:
            text = ln if "~" in ln else ln.split("=", 1)[1]
            used |= set(re.findall(r"\bt\d+\b", text))
        keep = [ln for ln in lines if "~" in ln or ln.split("=", 1)[0].strip() in used]
        if len(keep) == len(lines):
            break
        lines = keep
    g.lines = lines
    # every parameter has a prior, so none is outside the active set
    priors = ["a ~ Normal(0, 3.0)", "s ~ Gamma(2.0, 1.0)", "b ~ Normal(0, 2.0)", "w ~ Normal(0, 2.0)", "Bm ~ Normal(0, 2.0)"]
    return "\n".join(priors + g.lines)
