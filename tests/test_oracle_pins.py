"""Pins for the oracle: the only externally published known answers on this path.

The reference ships no tests; PCG32 is a published generator whose demo vector
(pcg32-demo, seed 42 / stream 54) pins rnd.Generator (src/base/random/generator.zig:1-47).
"""

import numpy as np

import oracle_lib as oracle
from zyg_b200 import scenes


def test_pcg32_published_vector():
    got = oracle.pcg32_uints(42, 54, 6)
    want = np.array([0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E], np.uint32)
    assert np.array_equal(got, want)


def test_pcg32_float_mantissa_trick():
    u = oracle.pcg32_uints(7, 3, 64)
    f = oracle.pcg32_floats(7, 3, 64)
    want = ((u & 0x007FFFFF) | 0x3F800000).view(np.float32) - np.float32(1.0)
    assert np.array_equal(f, want)
    assert (f >= 0).all() and (f < 1).all()


def test_host_pcg32_matches_oracle():
    seqs = np.array([0, 1, 2, 12345, 2**40 + 17], np.uint64)
    g = scenes.PCG32(0, seqs)
    draws = np.stack([g.uint() for _ in range(16)], axis=1)
    for row, s in zip(draws, seqs):
        assert np.array_equal(row, oracle.pcg32_uints(0, int(s), 16))
