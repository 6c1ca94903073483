"""A NumPy executor of the device op tape (include/simplemcmc_tape.h) as simplemcmc.jl_b200/lowering.py + fusion.py emit it --
TEST INFRASTRUCTURE: it lets the CPU suite check what the host front end hands to the library (statement order, register
programs, SMC_F_ACC / SMC_F_FWD_ONLY semantics) against the oracle's model function (oracle/refmodel.py) without a GPU.  One
chain; column-major element order; `evaluate(low, beta, with_grad)` follows eval_body of csrc/smc_core.cu: with a gradient
the forward statements flagged SMC_F_FWD_ONLY are skipped, without one everything past the forward part is.

SMC_OP_GLM and SMC_OP_AFFNORM are executed from their definitions in the header (eta = X * beta, the family's density, G =
X' * dl/deta; the affine residual into a Normal density with its adjoint scalars), not the way the kernels compute them.
"""
import numpy as np
from scipy.special import digamma

from oracle import densities

from simplemcmc_jl_b200 import fusion
from simplemcmc_jl_b200 import lowering as L

OPN = {v: k for k, v in L.OP.items()}
DISTN = {v: k for k, v in L.DIST.items()}


def dlogpdf(dist, k, a0, a1, a2):
    """d logpdf / d argument k per element: the rules of src/diff.jl:94-174 (the ::Real variants are these summed)"""
    name = DISTN[dist]
    if name == "Normal":
        d, s2 = a2 - a0, a1 * a1
        return d / s2 if k == 0 else ((d * d / s2 - 1.0) / a1 if k == 1 else (a0 - a2) / s2)
    if name == "Weibull":
        r = a2 / a1
        rs = r ** a0
        return (1.0 - rs) * np.log(r) + 1.0 / a0 if k == 0 else ((rs - 1.0) * a0 / a1 if k == 1 else ((1.0 - rs) * a0 - 1.0) / a2)
    if name == "Gamma":
        if k == 2:
            return -(a1 + a2 - a0 * a1) / (a1 * a2)
        return np.log(a2) - np.log(a1) - digamma(a0) if k == 0 else (a2 - a1 * a0) / (a1 * a1)
    if name == "Bernoulli":
        return 1.0 / (a0 - (1.0 - a1))
    if name == "Exponential":
        return -1.0 / a0 if k == 1 else (a1 - a0) / (a0 * a0)
    if name == "Poisson":
        return a1 / a0 - 1.0
    if name == "Binomial":
        return a2 / a1 - (a0 - a2) / (1.0 - a1)
    if name == "Cauchy":
        d = a2 - a0
        den = a1 * a1 + d * d
        return 2.0 * (a0 - a2) / den if k == 2 else (2.0 * d / den if k == 0 else (d * d - a1 * a1) / (a1 * den))
    if name == "Uniform":
        inside = ((a0 <= a2) & (a2 <= a1)).astype(np.float64)
        return inside / (a1 - a0) if k == 0 else (inside / (a0 - a1) if k == 1 else 0.0 * inside)
    if name == "Beta":
        if k == 2:
            return (a0 - 1.0) / a2 - (a1 - 1.0) / (1.0 - a2)
        return digamma(a0 + a1) - digamma(a0) + np.log(a2) if k == 0 else digamma(a0 + a1) - digamma(a1) + np.log(1.0 - a2)
    if name == "TDist":
        if k == 1:
            return -(a0 + 1.0) * a1 / (a0 + a1 * a1)
        t = a1 * a1 + a0
        return ((a1 * a1 - 1.0) / t + np.log(a0 / t) + digamma((a0 + 1.0) * 0.5) - digamma(a0 * 0.5)) * 0.5
    if name == "logNormal":
        t, lx = a1 * a1, np.log(a2)
        if k == 2:
            return (a0 - t - lx) / (t * a2)
        return (lx - a0) / t if k == 0 else (a0 * a0 - t - lx * (2.0 * a0 - lx)) / (a1 * t)
    if name == "TestDiff":
        return 1.0 + 0.0 * a0
    raise NotImplementedError("dlogpdf of " + name)


def apply(op, aux0, aux1, x):
    """one elementwise operation on broadcastable float64 operands (smc_apply, csrc/exec_common.cuh)"""
    name = OPN[op]
    a = x[0] if len(x) > 0 else 0.0
    b = x[1] if len(x) > 1 else 0.0
    c = x[2] if len(x) > 2 else 0.0
    with np.errstate(all="ignore"):
        if name == "COPY":
            return a + 0.0
        if name == "NEG":
            return -a
        if name in ("EXP", "LOG", "SIN", "COS", "ABS", "SIGN", "SQRT"):
            return getattr(np, {"ABS": "fabs"}.get(name, name.lower()))(a)
        if name == "DIGAMMA":
            return digamma(a)
        if name == "ZERO":
            return np.zeros_like(np.asarray(a, dtype=np.float64))
        if name == "ADD":
            return a + b
        if name == "SUB":
            return a - b
        if name == "MUL":
            return a * b
        if name == "DIV":
            return a / b
        if name == "POW":
            return np.power(a, b)
        if name == "MAX":
            return np.fmax(a, b)
        if name == "MIN":
            return np.fmin(a, b)
        if name in ("GT", "LT", "GE", "LE", "EQ"):
            f = {"GT": np.greater, "LT": np.less, "GE": np.greater_equal, "LE": np.less_equal, "EQ": np.equal}[name]
            return f(a, b).astype(np.float64)
        if name == "LOGPDF":
            return densities.elementwise(DISTN[aux0], *x)
        if name == "DLOGPDF":
            return dlogpdf(aux0, aux1, a, b, c)
    raise NotImplementedError(name)


class Machine(object):
    def __init__(self, low, beta):
        self.low = low
        self.beta = np.array(beta, dtype=np.float64)
        self.grad = np.zeros_like(self.beta)
        self.regs = {}
        self.bad = False
        self.launches = 0          # statements executed (one device launch each, give or take the split reductions)

    def get(self, v):
        if v.kind == L.REG_IMM:
            return np.full(1, v.imm)
        if v.kind == L.REG_DATA:
            return np.asarray(self.low.data_slots[v.slot][1], dtype=np.float64).flatten(order="F")
        if v.kind == L.REG_PARAM:
            return self.beta[v.slot:v.slot + v.numel]
        if v.kind == L.REG_GRAD:
            return self.grad[v.slot:v.slot + v.numel]
        return self.regs[v.vid]

    def put(self, v, val, acc=False):
        val = np.broadcast_to(np.asarray(val, dtype=np.float64).ravel(), (v.numel,)) if np.size(val) == 1 else np.asarray(val, dtype=np.float64).ravel()
        assert val.size == v.numel, (v, val.size)
        if v.kind == L.REG_GRAD:
            cur = self.grad[v.slot:v.slot + v.numel]
            self.grad[v.slot:v.slot + v.numel] = cur + val if acc else val
        else:
            assert v.kind in (L.REG_CHAIN, L.REG_HDR), v
            self.regs[v.vid] = (self.regs[v.vid] + val) if acc else val.copy()

    def elementwise(self, op, aux0, aux1, xs):
        r = apply(op, aux0, aux1, xs)
        if op == L.OP["LOGPDF"]:
            r = np.asarray(r, dtype=np.float64)
            give_up = np.isnan(r) | (r == -np.inf)
            if give_up.any():          # "give up eval", src/distribs.jl:69-70
                self.bad = True
                r = np.where(give_up, 0.0, r)
        return r

    def program(self, prog):
        n = prog.n
        loc = {}
        sums = [0.0] * len(prog.reductions)
        for (op, dst, srcs, flags, aux0, aux1, vid) in prog.prologue + prog.body:
            if op in (fusion.P_LOADS, fusion.P_LOADV):
                val = self.get(self.low.vals[vid])
                assert val.size in (1, n), (val.size, n)
                loc[dst] = val.copy()
            elif op == fusion.P_STOREV:
                self.put(self.low.vals[vid], np.broadcast_to(loc[srcs[0]], (n,)) if self.low.vals[vid].numel == n else loc[srcs[0]])
            elif op == fusion.P_REDUCE:
                sums[aux0] = sums[aux0] + float(np.sum(np.broadcast_to(loc[srcs[0]], (n,))))
            else:
                r = self.elementwise(op, aux0, aux1, [loc[s] for s in srcs])
                loc[dst] = (loc[dst] + r) if (flags & L.F_ACC) else np.asarray(r, dtype=np.float64)
        for k, (vid, acc) in enumerate(prog.reductions):
            self.put(self.low.vals[vid], sums[k], acc=bool(acc))

    def glm(self, st, acc):
        """SMC_OP_GLM (include/simplemcmc_tape.h): eta = X * beta; L = sum_i l(eta_i, y_i); G = X' * dl/deta"""
        Xv, b, y = st.srcs
        N, M = Xv.shape
        X = self.get(Xv).reshape((N, M), order="F")
        eta = X @ self.get(b)
        yv = self.get(y)
        fam, bits = st.aux0, st.aux1
        with np.errstate(all="ignore"):
            if fam == L.GLM_NORMAL:                        # l = logpdfNormal(0, sigma, a * eta + b * y), a, b = +-1
                a, bb, sigma = (-1.0 if bits & 1 else 1.0), (-1.0 if bits & 2 else 1.0), st.imm
                z = a * eta + bb * yv
                l = densities.elementwise("Normal", 0.0, sigma, z)
                dl = a * (0.0 - z) / (sigma * sigma)
            else:                                          # prob = 1 / (1 + exp(s * eta)), s = -1 when aux1 == 1
                sg = -1.0 if bits == 1 else 1.0
                pr = 1.0 / (1.0 + np.exp(sg * eta))
                if fam == L.GLM_BERNOULLI:
                    l = densities.elementwise("Bernoulli", pr, yv)
                    dlp = 1.0 / (pr - (1.0 - yv))
                elif fam == L.GLM_BINOMIAL:
                    l = densities.elementwise("Binomial", st.imm, pr, yv)
                    dlp = yv / pr - (st.imm - yv) / (1.0 - pr)
                else:
                    raise NotImplementedError("GLM family %d" % fam)
                dl = dlp * (-sg) * pr * (1.0 - pr)
        l = np.asarray(l, dtype=np.float64)
        if (np.isnan(l) | (l == -np.inf)).any():
            self.bad = True
            l, dl = np.zeros_like(l), np.zeros_like(dl)
        self.put(st.dst, np.sum(l), acc)
        self.put(st.aux2, X.T @ dl)

    def affnorm(self, st, acc):
        """SMC_OP_AFFNORM: dst (+)= sum_e logpdfNormal(mu, sigma, r_e), r_e = ((+-t_0) +- t_1) ..., t_j = D_j[e] * s_j | s_j; the
        adjoint scalars dL/dmu, dL/dsigma, dL/ds_j are assigned by the same statement"""
        d = st.desc
        n = d.n
        mu, sigma = float(self.get(st.srcs[0])[0]), float(self.get(st.srcs[1])[0])
        r, cols = None, []
        for j, (kind, sg, base, off) in enumerate(d.terms):
            sj = float(self.get(st.srcs[2 + j])[0])
            col = self.get(base)[off:off + n] if kind == 0 else np.ones(n)
            cols.append(col)
            t = sg * (col * sj)
            r = t if r is None else r + t
        dd = r - mu
        S2 = float(np.sum(dd * dd))
        if not sigma > 0.0:
            self.bad = True
            return self.put(st.dst, 0.0, acc)
        self.put(st.dst, -0.5 * S2 / sigma ** 2 - n * np.log(sigma) - n * 0.9189385332046727418, acc)
        outs = d.outs
        if outs[0] is not None:
            self.put(outs[0], np.sum(dd) / sigma ** 2)
        if outs[1] is not None:
            self.put(outs[1], (S2 / sigma ** 2 - n) / sigma)
        for j, (kind, sg, base, off) in enumerate(d.terms):
            if outs[2 + j] is not None:
                self.put(outs[2 + j], -sg * float(np.sum(dd * cols[j])) / sigma ** 2)

    def statement(self, st):
        self.launches += 1
        acc = bool(st.flags & L.F_ACC)
        name = OPN.get(st.op, "FUSED" if st.op == fusion.OP_FUSED else "?")
        if name == "FUSED":
            return self.program(self.low.programs[st.aux0])
        if name == "GLM":
            return self.glm(st, acc)
        if name == "AFFNORM":
            return self.affnorm(st, acc)
        if name == "MATMUL":
            r, k, n = st.aux0, st.aux1, int(st.imm)
            A, B = self.get(st.srcs[0]), self.get(st.srcs[1])
            A = A.reshape((k, r) if st.flags & L.F_TRANS_A else (r, k), order="F")
            B = B.reshape((n, k) if st.flags & L.F_TRANS_B else (k, n), order="F")
            A = A.T if st.flags & L.F_TRANS_A else A
            B = B.T if st.flags & L.F_TRANS_B else B
            return self.put(st.dst, (A @ B).flatten(order="F"), acc)
        if name == "TRANSPOSE":
            return self.put(st.dst, self.get(st.srcs[0]).reshape((st.aux0, st.aux1), order="F").T.flatten(order="F"), acc)
        if name == "GATHER":
            return self.put(st.dst, self.get(st.srcs[0])[self.low.idx_tables[st.aux0]], acc)
        if name in ("SCATTER_SET", "SCATTER_ADD"):
            src, idx = self.get(st.srcs[0]), self.low.idx_tables[st.aux0]
            cur = self.get(st.dst).copy()
            for e, i in enumerate(idx):                    # a sequential pass over the list (k_scatter_csr)
                cur[i] = cur[i] + src[e] if name == "SCATTER_ADD" else src[e]
            return self.put(st.dst, cur)
        xs = [self.get(s) for s in st.srcs]
        n_el = max([1] + [x.size for x in xs])
        if st.op == L.OP["LOGPDF"]:
            n_el = max(st.aux1, 1)
            xs = [x if x.size == 1 else x[:n_el] for x in xs]
        r = np.broadcast_to(self.elementwise(st.op, st.aux0, st.aux1, xs), (n_el,)) if n_el > 1 else np.asarray(
            self.elementwise(st.op, st.aux0, st.aux1, xs), dtype=np.float64).ravel()
        if st.dst.numel == 1 and n_el > 1:
            r = np.array([np.sum(r)])
        self.put(st.dst, r, acc)


def evaluate(low, beta, with_grad=True):
    """-> (logp, grad or None, statements executed); `low` is a Lowering whose serialise() has run"""
    m = Machine(low, beta)
    fwd, bwd = low.tape_fwd, low.tape_bwd
    if low.tape_result.kind == L.REG_CHAIN:
        m.regs[low.tape_result.vid] = np.zeros(1)          # zero between evaluations (smc_clear_result, csrc/smc_core.cu)
    for st in fwd + bwd:                                   # the let-header, once (src/parsing.jl:432-442)
        if st.dst.kind == L.REG_HDR:
            m.statement(st)
    m.launches = 0
    for st in fwd:
        if st.dst.kind == L.REG_HDR or (with_grad and (st.flags & L.F_FWD_ONLY)):
            continue
        if not with_grad and st.dst.kind == L.REG_GRAD:
            continue
        m.statement(st)
    if with_grad:
        for st in bwd:
            if st.dst.kind != L.REG_HDR:
                m.statement(st)
    logp = float(m.get(low.tape_result)[0])
    if m.bad or not logp > -np.inf:
        return -np.inf, (np.zeros_like(m.grad) if with_grad else None), m.launches
    return logp, (m.grad if with_grad else None), m.launches
