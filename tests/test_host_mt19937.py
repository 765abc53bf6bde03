"""Host logic (no GPU) of the stream-exact bandit mode: per-env RandomState states in the device layout and the numpy twin
of the generator arithmetic the kernel restates, pinned against numpy's own RandomState."""
import numpy as np
import pytest

from rl_envs_forge_b200 import mt19937

SEEDS = [0, 1, 42, 12345, 2 ** 31, 2 ** 32 - 1]


def test_initial_state_is_numpy_state_after_the_default_means():
    for k in (4, 5):                                  # odd k leaves the polar method's cached normal live
        key, pos, hg, gauss, means = mt19937.initial_state(SEEDS, k)
        assert key.shape == (624, len(SEEDS)) and key.dtype == np.uint32
        for i, s in enumerate(SEEDS):
            rs = np.random.RandomState(s)
            assert np.array_equal(means[i], rs.randn(k))
            st = rs.get_state()
            assert np.array_equal(key[:, i], st[1]) and pos[i] == st[2] and hg[i] == st[3] and gauss[i] == st[4]
        assert int(hg[0]) == k % 2


def test_generator_twin_matches_random_sample_across_regenerations():
    key, pos, _, _, _ = mt19937.initial_state(SEEDS, 4)
    refs = [np.random.RandomState(s) for s in SEEDS]
    for r in refs:
        r.randn(4)
    for _ in range(1000):                             # 2000 words per env: three regenerations of the 624-word state
        got = mt19937.next_double(key, pos)
        assert np.array_equal(got, np.array([r.random_sample() for r in refs]))
    for i, r in enumerate(refs):
        st = r.get_state()
        assert np.array_equal(key[:, i], st[1]) and pos[i] == st[2]


def test_env_seeds():
    assert mt19937.env_seeds(10, None, 5, 3) == [15, 16, 17]
    assert mt19937.env_seeds(None, [7, 8], 99, 2) == [7, 8]
    with pytest.raises(ValueError, match="2\\*\\*32 - 1"):
        mt19937.env_seeds(2 ** 32 - 2, None, 0, 4)
    with pytest.raises(ValueError, match="seed sequence"):
        mt19937.env_seeds(None, [1, 2, 3], 0, 4)
