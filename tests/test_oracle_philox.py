"""CPU: the host restatement of the product's Philox stream (oracle.philox4x32_10 / philox_uniforms).

The block function is pinned by the known-answer vectors of Random123 (``kat_vectors``, philox4x32 10 rounds); the
double construction is numpy's ``random_sample`` one.  tests/test_gpu_parity.py::test_bench_batch_philox_mode_vs_oracle
then feeds these doubles to the oracle and compares a Philox-mode GPU run value for value."""
import numpy as np

from oracle import densify_oracle as O

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = O.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want


def test_philox_block_is_vectorised_consistently():
    ctr = np.array([k[0] for k in KAT[:1]] * 3 + [(5, 0, 9, O.PHILOX_DOMAIN_TAG)], dtype=np.uint32)
    got = O.philox4x32_10(ctr, (0, 0))
    assert tuple(int(x) for x in got[0]) == KAT[0][2] and np.array_equal(got[0], got[2])
    assert not np.array_equal(got[0], got[3])


def test_uniform_construction_matches_the_kernel_recipe():
    seed, stream = 0x1234567890ABCDEF, 77
    u = O.philox_uniforms(seed, stream, 11, first=4)
    for j, d in enumerate(range(4, 15)):
        r = O.philox4x32_10(np.array([d >> 1, 0, stream, O.PHILOX_DOMAIN_TAG], dtype=np.uint32),
                            (seed & 0xFFFFFFFF, seed >> 32))
        a, b = (int(r[2]), int(r[3])) if d & 1 else (int(r[0]), int(r[1]))
        assert u[j] == ((a >> 5) * 67108864.0 + (b >> 6)) / 9007199254740992.0      # numpy random_sample's 53-bit double
    assert np.all((u >= 0.0) & (u < 1.0))
    # a window of the stream equals the same draws of a longer stream (counter-based: no state)
    assert np.array_equal(u, O.philox_uniforms(seed, stream, 40)[4:15])


def test_streams_and_seeds_are_independent():
    a = O.philox_uniforms(1, 0, 64)
    assert not np.array_equal(a, O.philox_uniforms(1, 1, 64))
    assert not np.array_equal(a, O.philox_uniforms(2, 0, 64))
    assert abs(float(O.philox_uniforms(3, 5, 20000).mean()) - 0.5) < 0.01
