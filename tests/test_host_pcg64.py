"""Host logic (no GPU): the vectorised SeedSequence / PCG64 initialisation and the 53-bit thresholds of the stream-exact
GridWorld mode, pinned against numpy's own SeedSequence, PCG64 and Generator."""
import numpy as np

from rl_envs_forge_b200 import pcg64

SEEDS = [0, 1, 2, 42, 12345, 2 ** 31, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 17, 2 ** 63 + 5, 2 ** 64 - 1, 2 ** 64, 2 ** 100 + 3,
         2 ** 127 + 9]


def test_seed_sequence_pool_matches_numpy():
    pool = pcg64.seed_sequence_pool(SEEDS)
    for i, s in enumerate(SEEDS):
        assert [int(x) for x in pool[i]] == [int(x) for x in np.random.SeedSequence(s).pool]


def test_initial_state_matches_numpy_pcg64():
    state, inc = pcg64.pcg64_initial_state(SEEDS)
    for i, s in enumerate(SEEDS):
        st = np.random.PCG64(np.random.SeedSequence(s)).state["state"]
        assert (int(state[i, 0]) << 64) | int(state[i, 1]) == st["state"]
        assert (int(inc[i, 0]) << 64) | int(inc[i, 1]) == st["inc"]


def test_uniform_stream_matches_generator_random():
    seeds = [0, 42, 2 ** 40 + 17]
    state, inc = pcg64.pcg64_initial_state(seeds)
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))) for s in seeds]
    for _ in range(200):
        m = pcg64.pcg64_next_uniform53(state, inc)
        for i, g in enumerate(gens):
            assert g.random() == int(m[i]) * 2.0 ** -53


def test_thresholds53_equal_numpy_choice_rule():
    rng = np.random.default_rng(0)
    for p in (0.0, 0.25, 0.5, 0.7, 0.8, 0.9, 1 - 2 ** -30, float(np.nextafter(1.0, 0.0))):
        thr, cap = pcg64.thresholds53(p)
        cdf = np.cumsum([p] + [(1 - p) / 3] * 3)
        cdf /= cdf[-1]
        ms = {0, 1, 2 ** 53 - 1, 2 ** 52}
        for t in thr:
            ms |= {max(t - 1, 0), min(t, 2 ** 53 - 1), min(t + 1, 2 ** 53 - 1)}
        ms |= set(int(x) for x in rng.integers(0, 2 ** 53, 3000))
        for m in ms:
            want = int(np.searchsorted(cdf, m * 2.0 ** -53, side="right"))
            assert min(sum(m >= t for t in thr), cap) == want, (p, m)
