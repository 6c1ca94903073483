"""ORACLE (test infrastructure only).  Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy
as 1, 2, 3", SC'11) in NumPy, pinned by the Random123 known-answer vectors (tests/test_philox.py), plus the
counter layout and the uniform / Box-Muller transforms of the device stream
(simplemcmc.jl_b200/csrc/philox.cuh).  The reference has no counterpart: it draws from Julia's global RNG
(randn(nparams), rand(): /root/reference/src/samplers.jl:151,162)."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32"""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def _block(seed, chain, step, stream, block):
    block = np.asarray(block, dtype=np.uint64)
    ctr = np.zeros(block.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = block & MASK
    ctr[..., 1] = np.uint32(stream | (((chain >> 32) & 0xFFFFFF) << 8))
    ctr[..., 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32(chain & 0xFFFFFFFF)
    key = np.zeros(block.shape + (2,), dtype=np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32(((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF)
    return philox4x32_10(ctr, key)


def _u53(hi, lo):
    x = ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))
    return (x.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def uniforms(seed, chain, step, n):
    k = np.arange(n)
    r = _block(seed, chain, step, 1, k >> 1)
    even = _u53(r[:, 0], r[:, 1])
    odd = _u53(r[:, 2], r[:, 3])
    return np.where(k & 1, odd, even)


def normals(seed, chain, step, n):
    nb = (n + 1) // 2
    r = _block(seed, chain, step, 0, np.arange(nb))
    u1, u2 = _u53(r[:, 0], r[:, 1]), _u53(r[:, 2], r[:, 3])
    rad = np.sqrt(-2.0 * np.log(u1))
    z = np.empty(2 * nb)
    z[0::2] = rad * np.cos(2.0 * np.pi * u2)
    z[1::2] = rad * np.sin(2.0 * np.pi * u2)
    return z[:n]
