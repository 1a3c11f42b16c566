"""Pins the oracle's (and the generators') Philox4x32-10 to the Random123 known-answer vectors."""
import numpy as np
import pytest

import oracle
from embiggen_b200.graph import philox4x32

# (key0, key1, c0, c1, c2, c3) -> output; Random123 kat_vectors, philox4x32 with 10 rounds
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000, 0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0xa4093822, 0x299f31d0, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("inputs,expected", KAT)
def test_oracle_known_answers(inputs, expected):
    k0, k1, c0, c1, c2, c3 = inputs
    assert oracle.philox(k0 | (k1 << 32), c0, c1, c2, c3) == expected


@pytest.mark.parametrize("inputs,expected", KAT)
def test_numpy_known_answers(inputs, expected):
    k0, k1, c0, c1, c2, c3 = inputs
    out = philox4x32(k0 | (k1 << 32), c0, c1, c2, c3)
    assert tuple(int(x) for x in out) == expected


def test_numpy_matches_oracle_on_random_counters():
    rng = np.random.default_rng(0)
    counters = rng.integers(0, 2 ** 32, size=(64, 4), dtype=np.uint64)
    seed = 0x1234567890ABCDEF
    out = philox4x32(seed, *(counters[:, i] for i in range(4)))
    for row in range(64):
        expected = oracle.philox(seed, *(int(x) for x in counters[row]))
        assert tuple(int(o[row]) for o in out) == expected
