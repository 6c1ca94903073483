"""CPU check of the arithmetic of the sliced-integer GLM prototype (tools/ozaki_glm_proto.cu): the
exact integer algorithm of the v4/v5 kernels -- seven balanced 8-bit slices per operand, slice pairs with i + j <= 6
accumulated per group i + j in (emulated) int32, the stacked two-slice arrangement of product 2 whose second lane block
holds the groups shifted by one, recombination from two int64 halves, a power-of-two scale per (128-row tile, chain) for
the adjoint -- restated in NumPy and compared with 80-bit long double on configs[1]-shaped data.  The bar is the parity
bar of the GPU suite: 1e-12 relative to max |.| per chain.  No GPU, no oracle: this pins the algorithm, the GPU test
`test_zz_gpu_prototype.py` pins the kernel."""
import numpy as np
import pytest

NS, TR, H8 = 7, 128, sum(128 << (8 * k) for k in range(7))


def _scale(m):
    """power of two strictly above m (1 for m == 0), as scale_of8 takes it from the exponent field"""
    m = np.asarray(m, dtype=np.float64)
    return np.where(m > 0, 2.0 ** np.frexp(np.where(m > 0, m, 1.0))[1], 1.0)


def _slices(v, scale):
    """slice j (0 = most significant) = byte 6-j of (round(v 2^54 / scale) + H8), minus 128.  (X and beta are rounded to
    nearest in the kernels; the adjoint is truncated by slice7_bits or rounded by slice7_magic -- its share of the error is
    1e-16 either way.)"""
    t = np.rint(v * (2.0 ** 54 / scale)).astype(np.int64)
    u = (t + H8).astype(np.uint64)
    assert np.all(u < (1 << 56))
    return [((u >> np.uint64(8 * (6 - j))) & np.uint64(255)).astype(np.int64) - 128 for j in range(NS)]


def _recombine(acc, shift_groups):
    """sum_cb acc[cb] 2^(-12 - 8 (cb + shift_groups)) through two int64 halves"""
    assert np.all(np.abs(acc) < 2 ** 31)                      # what the tensor core accumulates in int32
    hi = (acc[0] << 24) + (acc[1] << 16) + (acc[2] << 8) + acc[3]
    lo = (acc[4] << 16) + (acc[5] << 8) + acc[6]
    assert np.all(np.abs(hi) < 2 ** 51)
    s = 2.0 ** (-8 * shift_groups)
    return hi.astype(np.float64) * (2.0 ** -36 * s) + lo.astype(np.float64) * (2.0 ** -60 * s)


def sliced_glm(X, Y, B):
    """(logp, G) of Y - X B ~ Normal(0, 1) without the constant, the way k_ozaki_glm_v5 computes them"""
    N, M = X.shape
    C = B.shape[1]
    sx = float(_scale(np.abs(X).max()))
    QX = _slices(X, sx) + [np.zeros(X.shape, np.int64)]       # slot 7 = zeros (second half of the pair (6, 7))
    sb = _scale(np.abs(B).max(axis=0))
    QB = _slices(B, sb[None, :])
    L, G = np.zeros(C), np.zeros((M, C))
    for r0 in range(0, N, TR):
        rows = slice(r0, r0 + TR)
        acc = np.zeros((NS, TR, C), np.int64)                 # product 1: column block cb = i + j
        for i in range(NS):
            for j in range(NS - i):
                acc[i + j] += QX[i][rows] @ QB[j]
        eta = _recombine(acc, 0) * (sx * sb[None, :])
        d = Y[rows, None] - eta
        L += -0.5 * np.sum(d * d, axis=0)
        sd = _scale(np.abs(d).max(axis=0))                    # per (tile, chain)
        QD = _slices(d, sd[None, :])
        acc2 = np.zeros((NS, 2, M, C), np.int64)              # product 2: lane block ii holds group cb + ii
        for i in range(0, NS, 2):
            for j in range(NS - i):
                for ii in range(2):
                    acc2[i + j, ii] += QX[i + ii][rows].T @ QD[j]
        for ii in range(2):
            G += _recombine(acc2[:, ii], ii) * (sx * sd[None, :])
    return L, G


def _problem(N, M, C, seed, spread=1e-3):
    rng = np.random.default_rng(seed)
    X = np.column_stack([np.ones(N), rng.standard_normal((N, M - 1))])
    b0 = rng.standard_normal(M)
    Y = X @ b0 + rng.standard_normal(N)
    B = b0[:, None] + spread * rng.standard_normal((M, C))
    return X, Y, B


def _truth(X, Y, B):
    Xl, Yl, Bl = X.astype(np.longdouble), Y.astype(np.longdouble), B.astype(np.longdouble)
    d = Yl[:, None] - Xl @ Bl
    return -0.5 * np.sum(d * d, axis=0), Xl.T @ d


@pytest.mark.parametrize("N,M,C,seed,spread", [(2048, 64, 8, 2, 1e-3), (1024, 64, 5, 7, 1.0), (512, 10, 3, 11, 1e-6)])
def test_sliced_glm_meets_the_parity_bar(N, M, C, seed, spread):
    X, Y, B = _problem(N, M, C, seed, spread)
    L, G = sliced_glm(X, Y, B)
    Lt, Gt = _truth(X, Y, B)
    assert float(np.max(np.abs((L - Lt) / Lt))) <= 1e-13
    assert float(np.max(np.abs(G - Gt) / np.max(np.abs(Gt), axis=0))) <= 1e-12
    # and it is as good as the FP64 evaluation it is meant to replace (within a small factor)
    G64 = X.T @ (Y[:, None] - X @ B)
    e64 = float(np.max(np.abs(G64 - Gt) / np.max(np.abs(Gt), axis=0)))
    assert float(np.max(np.abs(G - Gt) / np.max(np.abs(Gt), axis=0))) <= max(20 * e64, 1e-13)


def test_slices_are_an_exact_representation():
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(1000) * 10.0 ** rng.integers(-12, 3, 1000), [0.0, -0.0, 1.0, -1.0, 4.0 - 2 ** -50]])
    scale = float(_scale(np.abs(v).max()))
    q = _slices(v, scale)
    for s in q:
        assert s.min() >= -128 and s.max() <= 127              # int8
    t = sum(qj * (1 << (8 * (6 - j))) for j, qj in enumerate(q))
    assert np.array_equal(t, np.rint(v * (2.0 ** 54 / scale)).astype(np.int64))
    assert np.max(np.abs(t.astype(np.float64) * (scale / 2.0 ** 54) - v)) <= scale * 2.0 ** -54


def test_degenerate_tiles():
    # a chain that reproduces Y exactly (zero adjoint: the scale falls back to 1) and rows of zeros (padding)
    rng = np.random.default_rng(5)
    N, M = 256, 8
    X = np.round(rng.standard_normal((N, M)) * 8) / 8          # exactly representable products
    b = np.round(rng.standard_normal(M) * 4) / 4
    X[200:] = 0.0
    Y = X @ b
    B = np.column_stack([b, b + 0.5])
    L, G = sliced_glm(X, Y, B)
    Lt, Gt = _truth(X, Y, B)
    assert L[0] == 0.0 and np.all(G[:, 0] == 0.0)
    assert abs(L[1] - float(Lt[1])) <= 1e-13 * abs(float(Lt[1]))
    assert np.max(np.abs(G[:, 1] - Gt[:, 1].astype(np.float64))) <= 1e-12 * float(np.max(np.abs(Gt[:, 1])))


def test_magic_number_slicing():
    """slice7_magic of the prototype's v7: t = rint(v 2^54 / scale) from two magic-number roundings (low words of
    x + 1.5 2^52), as IEEE double arithmetic in NumPy; equal to the direct rounding, |t| <= 2^54, so t + H8 fits 56 bits"""
    magic = np.float64(6755399441055744.0)

    def lo32(a):
        return (a.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).astype(np.int64)

    rng = np.random.default_rng(0)
    for trial in range(40):
        v = rng.standard_normal(4000) * 10.0 ** rng.integers(-8, 3)
        v[:5] = [0.0, -0.0, 1e-300, -1e-300, v[5] * 1e-20]
        emax = max(int(np.frexp(np.abs(v).max())[1]) + 1022, 64)          # biased exponent of the maximum, clamped as in the kernel
        f, f27 = np.ldexp(1.0, 54 - (emax - 1022)), np.ldexp(1.0, 27 - (emax - 1022))
        a = v * f27 + magic                       # the product is exact (power of two), so this is the fused multiply-add
        h = lo32(a)
        r = v * f - (a - magic) * np.ldexp(1.0, 27)
        t = h * (1 << 27) + lo32(r + magic)
        assert np.array_equal(t, np.rint(v * f).astype(np.int64))
        assert np.abs(t).max() <= (1 << 54)
        assert np.all((t + H8) >= 0) and np.all((t + H8) < (1 << 56))


def _slices_beta8(v, scale):
    """slice8_word of csrc/glm_i8.cu: eight balanced slices of round(v 2^62 / scale), slice j = byte 7-j of (t + 0x80..80)"""
    t = np.rint(v * (2.0 ** 62 / scale)).astype(np.int64)
    u = t.astype(np.uint64) + np.uint64(0x8080808080808080)          # wraps modulo 2^64 like the device arithmetic
    q = [((u >> np.uint64(8 * (7 - j))) & np.uint64(255)).astype(np.int64) - 128 for j in range(8)]
    assert np.array_equal(sum(int(1 << (8 * (7 - j))) * qj for j, qj in enumerate(q)), t)
    return q


def sliced_eta_beta8(X, B):
    """product 1 with eight beta slices (SMC_GLM_I8_BETA8=1): groups i + j <= 7, recombine8"""
    sx = float(_scale(np.abs(X).max()))
    QX = _slices(X, sx)
    sb = _scale(np.abs(B).max(axis=0))
    QB = _slices_beta8(B, sb[None, :])
    acc = np.zeros((8,) + (X.shape[0], B.shape[1]), np.int64)
    for i in range(NS):
        for j in range(8 - i):
            acc[i + j] += QX[i] @ QB[j]
    assert np.all(np.abs(acc) < 2 ** 31)
    hi = (acc[0] << 24) + (acc[1] << 16) + (acc[2] << 8) + acc[3]
    lo = (acc[4] << 24) + (acc[5] << 16) + (acc[6] << 8) + acc[7]
    return (hi.astype(np.float64) * 2.0 ** -36 + lo.astype(np.float64) * 2.0 ** -68) * (sx * sb[None, :])


def test_eight_beta_slices():
    X, Y, B = _problem(4096, 64, 6, 13)
    B[5, :] *= 1e-4                                              # a component far below the chain's maximum
    sx, sb = float(_scale(np.abs(X).max())), _scale(np.abs(B).max(axis=0))
    QX, QB7 = _slices(X, sx), _slices(B, sb[None, :])
    acc7 = np.zeros((NS, X.shape[0], B.shape[1]), np.int64)
    for i in range(NS):
        for j in range(NS - i):
            acc7[i + j] += QX[i] @ QB7[j]
    eta7 = _recombine(acc7, 0) * (sx * sb[None, :])
    eta8 = sliced_eta_beta8(X, B)
    truth = X.astype(np.longdouble) @ B.astype(np.longdouble)
    e7 = float(np.max(np.abs(eta7 - truth)) / np.max(np.abs(truth)))
    e8 = float(np.max(np.abs(eta8 - truth)) / np.max(np.abs(truth)))
    assert e8 <= 1e-15 and e8 <= e7
    # the gradient of the quadratic form amplifies the resolution of beta by N: X'(X (B8 - B)) against X'(X (B7 - B))
    B7 = np.rint(B * (2.0 ** 54 / sb)) * (sb / 2.0 ** 54)
    B8 = np.rint(B * (2.0 ** 62 / sb)) * (sb / 2.0 ** 62)
    assert np.max(np.abs(B8 - B)) <= np.max(np.abs(B7 - B)) / 100 or np.max(np.abs(B7 - B)) == 0.0
