"""Pins oracle/philox.py with the Random123 Philox4x32-10 known-answer vectors."""
import numpy as np

from oracle import philox

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def test_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(*ctr, *key)
        assert tuple(int(x) for x in got) == want


def test_vectorised_matches_scalar():
    g = np.arange(1000, dtype=np.uint64)
    w = philox.draw_block(0x123456789ABCDEF, g, 77, philox.STREAM_BANDIT)
    for i in (0, 1, 500, 999):
        s = philox.draw_block(0x123456789ABCDEF, i, 77, philox.STREAM_BANDIT)
        assert all(int(w[k][i]) == int(s[k]) for k in range(4))


def test_step_words_grouping():
    g = np.arange(3, 35)
    for t in (0, 1, 2, 3, 2 ** 33 + 5):
        a, b = philox.step_words(9, g, np.full(len(g), t, np.uint64))
        for i, gi in enumerate(g):
            wt = philox.draw_block(9, int(gi) >> 2, t, philox.STREAM_TRANSITION)
            wp = philox.draw_block(9, int(gi) >> 2, t, philox.STREAM_POLICY)
            assert int(a[i]) == int(wt[int(gi) & 3]) and int(b[i]) == int(wp[int(gi) & 3])


def test_u23_open_interval_and_fp32_exact():
    w = np.array([0, 1, 511, 512, 2 ** 31, 2 ** 32 - 1], np.uint32)
    u = philox.u23(w)
    assert (u > 0).all() and (u < 1).all()
    assert (u.astype(np.float32).astype(np.float64) == u).all()
    assert ((1.0 - u).astype(np.float32).astype(np.float64) == 1.0 - u).all()
