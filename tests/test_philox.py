"""Philox4x32-10 restatement against the Random123 known-answer vectors (kat_vectors of the Random123 distribution)."""
import numpy as np

from oracle import philox


def _run(ctr, key):
    return [int(x) for x in philox.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))]


def test_known_answers():
    assert _run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert _run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_streams_are_uniform_and_normal():
    u = np.concatenate([philox.uniforms(7, c, 3, 1000) for c in range(20)])
    z = np.concatenate([philox.normals(7, c, 3, 1000) for c in range(20)])
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1) < 0.03
    assert not np.array_equal(philox.uniforms(7, 0, 3, 8), philox.uniforms(7, 1, 3, 8))
    assert not np.array_equal(philox.uniforms(7, 0, 3, 8), philox.uniforms(7, 0, 4, 8))
