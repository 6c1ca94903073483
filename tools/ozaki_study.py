"""Numerics of an FP64-accurate GLM contraction from int8 slices (Ozaki scheme), emulated in NumPy -- the feasibility
study behind DESIGN.md section 8 "What would lift the FP64 tensor bound itself".  Not part of the product path.

Both contractions of configs[1] (eta = X*beta, G = X'*d with d = y - eta) are computed three ways on a row subset:
  truth   80-bit long double accumulation
  fp64    what the DMMA kernel does (FP64 products and sums)
  sliced  operands split into S signed 7-bit slices against a power-of-two scale; slice products summed exactly in
          integers (int32 on the tensor cores; int64 here), groups of equal shift recombined in FP64.  X is sliced once
          with one global scale; beta per chain; the adjoint d per (128-row tile, chain) because its magnitude is only
          known after the family stage.
Prints the worst relative errors (relative to max|.| per chain, the parity metric of tests/test_gpu_parity.py)."""
import json, sys
import numpy as np

N, M, C, TILE = int(sys.argv[1]) if len(sys.argv) > 1 else 32768, 64, 32, 128
rng = np.random.default_rng(2)
X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
b0 = rng.standard_normal(M)
Y = X @ b0 + rng.standard_normal(N)
mode = np.linalg.solve(X.T @ X + np.eye(M), X.T @ Y)
B = mode[:, None] + 1e-3 * rng.standard_normal((M, C))          # chains near the posterior mode, as in bench.py


def slices(A, scale, S):
    """A / scale in (-1, 1) -> S integer arrays q_k in [-127, 127] with A = scale * sum_k q_k 2^(-7k) + O(2^(-7S))"""
    r = A / scale
    out = []
    for _ in range(S):
        r = r * 128.0
        q = np.trunc(r)
        out.append(q.astype(np.int64))
        r = r - q                                             # exact: power-of-two scaling, integer part removed
    return out


def pow2_above(v):
    return 2.0 ** np.ceil(np.log2(np.maximum(v, 1e-300)) + 1e-12)


def sliced_matmul(QA, QB, S):
    """sum over slice pairs with i + j <= S + 1 (the others are below 2^(-7S)); equal shifts are added as integers"""
    groups = {}
    for i, qa in enumerate(QA, 1):
        for j, qb in enumerate(QB, 1):
            if i + j <= S + 1:
                groups[i + j] = groups.get(i + j, 0) + qa @ qb
    acc = 0.0
    for sh in sorted(groups, reverse=True):                   # smallest contributions first
        acc = acc + groups[sh].astype(np.float64) * 2.0 ** (-7 * sh)
    return acc, sum(1 for i in range(1, S + 1) for j in range(1, S + 1) if i + j <= S + 1)


Xl, Bl, Yl = X.astype(np.longdouble), B.astype(np.longdouble), Y.astype(np.longdouble)
eta_t = Xl @ Bl
d_t = Yl[:, None] - eta_t
G_t = Xl.T @ d_t
L_t = -0.5 * np.sum(d_t * d_t, axis=0)
eta64 = X @ B
d64 = Y[:, None] - eta64
G64 = X.T @ d64
L64 = -0.5 * np.sum(d64 * d64, axis=0)


def rel(a, t):
    return float(np.max(np.abs(a.astype(np.longdouble) - t) / np.max(np.abs(t), axis=0)))


res = {"N": N, "M": M, "chains": C, "fp64": {"eta": rel(eta64, eta_t), "logp": float(np.max(np.abs((L64 - L_t) / L_t))), "grad": rel(G64, G_t)}}
sx = pow2_above(np.max(np.abs(X)))
for S in (6, 7, 8):
    QX = slices(X, sx, S)
    sb = pow2_above(np.max(np.abs(B), axis=0))                # per chain
    QB = slices(B, sb[None, :], S)
    eta_s, npairs = sliced_matmul(QX, QB, S)
    eta_s = eta_s * sx * sb[None, :]
    d_s = Y[:, None] - eta_s
    L_s = -0.5 * np.sum(d_s * d_s, axis=0)
    G_s = np.zeros((M, C))
    QXT = [q.T for q in QX]
    for r0 in range(0, N, TILE):                              # second product: per-tile scale of the adjoint
        dt = d_s[r0:r0 + TILE]
        sd = pow2_above(np.max(np.abs(dt), axis=0))
        QD = slices(dt, sd[None, :], S)
        g, _ = sliced_matmul([q[:, r0:r0 + TILE] for q in QXT], QD, S)
        G_s += g * sx * sd[None, :]
    res["S=%d" % S] = {"slice_pairs": npairs, "eta": rel(eta_s, eta_t), "logp": float(np.max(np.abs((L_s - L_t) / L_t))),
                       "grad": rel(G_s, G_t)}


# ---- the scheme the device kernels ended up with (tools/ozaki_glm_proto.cu v4/v5, csrc/glm_i8.cu): seven BALANCED 8-bit
# slices (digits in [-128, 127] of round(v 2^54 / scale) + 0x80...80), pairs with i + j <= 6 (0-based): 28 int8 GEMMs
H8 = sum(128 << (8 * k) for k in range(7))


def slices8(A, scale):
    t = np.rint(A * (2.0 ** 54 / scale)).astype(np.int64)       # to nearest, as the kernels slice X and beta
    u = (t + H8).astype(np.uint64)
    return [((u >> np.uint64(8 * (6 - j))) & np.uint64(255)).astype(np.int64) - 128 for j in range(7)]


def sliced_matmul8(QA, QB):
    groups = {}
    for i, qa in enumerate(QA):
        for j, qb in enumerate(QB):
            if i + j <= 6:
                groups[i + j] = groups.get(i + j, 0) + qa @ qb
    acc = 0.0
    for g in sorted(groups, reverse=True):
        acc = acc + groups[g].astype(np.float64) * 2.0 ** (-12 - 8 * g)
    return acc


QX8 = slices8(X, sx)
sb = pow2_above(np.max(np.abs(B), axis=0))
eta_8 = sliced_matmul8(QX8, slices8(B, sb[None, :])) * sx * sb[None, :]
d_8 = Y[:, None] - eta_8
L_8 = -0.5 * np.sum(d_8 * d_8, axis=0)
G_8 = np.zeros((M, C))
QXT8 = [q.T for q in QX8]
for r0 in range(0, N, TILE):
    dt = d_8[r0:r0 + TILE]
    sd = pow2_above(np.max(np.abs(dt), axis=0))
    G_8 += sliced_matmul8([q[:, r0:r0 + TILE] for q in QXT8], slices8(dt, sd[None, :])) * sx * sd[None, :]
res["7 balanced 8-bit slices"] = {"slice_pairs": 28, "eta": rel(eta_8, eta_t), "logp": float(np.max(np.abs((L_8 - L_t) / L_t))),
                                  "grad": rel(G_8, G_t)}
print(json.dumps(res, indent=1))
