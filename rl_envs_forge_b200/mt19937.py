"""Host side of the stream-exact bandit mode (`rng="mt19937"`): per-env numpy `RandomState` initial states in the layout
the device kernel reads, plus a numpy twin of the generator arithmetic the kernel restates.

The reference builds `np.random.RandomState(seed)` per instance (rl_envs_forge/envs/k_armed_bandit/k_armed_bandit.py:145,
again in reset :215) and draws its default means with `randn(k)` (:162) before any reward.  `initial_state` does exactly
that with numpy itself -- seeding (init_genrand / init_by_array) is numpy's own code, not a restatement -- and returns the
generator state AFTER the means were drawn:
    key [624, n] uint32 word-major, pos [n] int32, has_gauss [n] int32, gauss [n] float64, means [n, k] float64.
`next_double` is the host twin of csrc/bandit_mt19937.cu (mt19937_gen + tempering + numpy's legacy next_double) on that
layout; tests/test_host_mt19937.py pins it against `RandomState.random_sample`.
"""
import numpy as np

N, M = 624, 397
_MATRIX_A, _UPPER, _LOWER = np.uint32(0x9908B0DF), np.uint32(0x80000000), np.uint32(0x7FFFFFFF)


def env_seeds(seed, seed_list, env_offset, num_envs):
    """Seed of every env: `seed_list` verbatim, else seed + global env index; numpy's own range check and message."""
    seeds = list(seed_list) if seed_list is not None else [int(seed) + int(env_offset) + i for i in range(num_envs)]
    if len(seeds) != num_envs:
        raise ValueError(f"seed sequence must hold {num_envs} integers")
    for s in seeds:
        if s < 0 or s > 0xFFFFFFFF:
            raise ValueError("Seed must be between 0 and 2**32 - 1")
    return seeds


def initial_state(seeds, k):
    n = len(seeds)
    key = np.empty((N, n), np.uint32)
    pos, has_gauss = np.empty(n, np.int32), np.empty(n, np.int32)
    gauss, means = np.empty(n, np.float64), np.empty((n, k), np.float64)
    for i, s in enumerate(seeds):
        rs = np.random.RandomState(s)
        means[i] = rs.randn(k)
        st = rs.get_state()
        key[:, i], pos[i], has_gauss[i], gauss[i] = st[1], st[2], st[3], st[4]
    return key, pos, has_gauss, gauss, means


def _regenerate(key, cols):
    """mt19937_gen on the selected columns, in place (sequential in the word index like the C loop)."""
    mt = key[:, cols].astype(np.uint32)
    for kk in range(N - M):
        y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
        mt[kk] = mt[kk + M] ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), _MATRIX_A, np.uint32(0))
    for kk in range(N - M, N - 1):
        y = (mt[kk] & _UPPER) | (mt[kk + 1] & _LOWER)
        mt[kk] = mt[kk + (M - N)] ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), _MATRIX_A, np.uint32(0))
    y = (mt[N - 1] & _UPPER) | (mt[0] & _LOWER)
    mt[N - 1] = mt[M - 1] ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), _MATRIX_A, np.uint32(0))
    key[:, cols] = mt


def next_u32(key, pos):
    """One tempered 32-bit output per env; `key` [624, n] and `pos` [n] are updated in place."""
    full = np.nonzero(pos == N)[0]
    if len(full):
        _regenerate(key, full)
        pos[full] = 0
    y = key[pos, np.arange(key.shape[1])].astype(np.uint32)
    pos += 1
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


def next_double(key, pos):
    """numpy's legacy next_double: ((a >> 5) * 2**26 + (b >> 6)) / 2**53 from two consecutive outputs."""
    a = (next_u32(key, pos) >> np.uint32(5)).astype(np.float64)
    b = (next_u32(key, pos) >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) / 9007199254740992.0
