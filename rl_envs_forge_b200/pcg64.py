"""Vectorised numpy SeedSequence -> PCG64 initial state, for the stream-exact GridWorld mode (`rng="pcg64"`).

The reference seeds each GridWorld with `gym.utils.seeding.np_random(seed)` = `Generator(PCG64(SeedSequence(seed)))`
(rl_envs_forge/envs/grid_world/grid_world.py:69) and draws one `Generator.random()` double per step inside
`Generator.choice` (:515).  To reproduce `GridWorld(seed=s)` without injecting draws, every env carries its own PCG64
state (128-bit state + 128-bit increment) in device memory; this module computes those initial states for an array of
integer seeds exactly like numpy does (SeedSequence pool mixing, `generate_state(4, uint64)`, `pcg64_set_seed`), and the
53-bit integer thresholds equivalent to numpy's `searchsorted(cdf, u, 'right')` on `u = (raw >> 11) * 2**-53`.
Pinned against numpy itself in tests/test_host_pcg64.py.
"""
import math

import numpy as np

_INIT_A, _MULT_A = 0x43B0D7E5, 0x931E8875
_INIT_B, _MULT_B = 0x8B51F9DD, 0x58F38DED
_MIX_L, _MIX_R = 0xCA01F9DD, 0x4973F715
_POOL = 4
_M32 = np.uint64(0xFFFFFFFF)
PCG_MULT_HI, PCG_MULT_LO = 0x2360ED051FC65DA4, 0x4385DF649FCCF645
_MASK64 = (1 << 64) - 1
_MASK128 = (1 << 128) - 1


def _u32(x):
    return np.asarray(x, dtype=np.uint64) & _M32


class _HashConst:
    def __init__(self, v):
        self.v = v & 0xFFFFFFFF

    def hashmix(self, value):
        value = _u32(value) ^ np.uint64(self.v)
        self.v = (self.v * _MULT_A) & 0xFFFFFFFF
        value = (value * np.uint64(self.v)) & _M32
        return value ^ (value >> np.uint64(16))


def _mix(x, y):
    r = (np.uint64(_MIX_L) * x - np.uint64(_MIX_R) * y) & _M32
    return r ^ (r >> np.uint64(16))


def seed_sequence_pool(seeds):
    """SeedSequence(seed).pool for every non-negative integer seed (< 2**128), as uint64-held uint32 [N,4]."""
    seeds = [int(s) for s in np.asarray(seeds, dtype=object).reshape(-1)]
    if any(s < 0 for s in seeds):
        raise ValueError("seeds must be non-negative integers")
    n_words = max(1, max((s.bit_length() + 31) // 32 for s in seeds))
    words = np.zeros((len(seeds), n_words), np.uint64)
    lens = np.zeros(len(seeds), np.int64)
    for i, s in enumerate(seeds):
        k = 0
        while True:
            words[i, k] = s & 0xFFFFFFFF
            s >>= 32
            k += 1
            if s == 0:
                break
        lens[i] = k
    if n_words > _POOL and len(set(lens.tolist())) > 1:
        raise ValueError("seeds wider than 128 bits must all have the same width")
    hc = _HashConst(_INIT_A)
    pool = np.zeros((len(seeds), _POOL), np.uint64)
    for i in range(_POOL):
        # numpy hashes entropy[i] when i < len(entropy) else 0 -- the padding words are zero already
        pool[:, i] = hc.hashmix(words[:, i] if i < n_words else np.zeros(len(seeds), np.uint64))
    for i_src in range(_POOL):
        for i_dst in range(_POOL):
            if i_src != i_dst:
                pool[:, i_dst] = _mix(pool[:, i_dst], hc.hashmix(pool[:, i_src]))
    for i_src in range(_POOL, n_words):
        for i_dst in range(_POOL):
            pool[:, i_dst] = _mix(pool[:, i_dst], hc.hashmix(words[:, i_src]))
    return pool


def _generate_state_u64x4(pool):
    hc = _INIT_B
    out32 = np.zeros((pool.shape[0], 8), np.uint64)
    for i in range(8):
        v = pool[:, i % _POOL] ^ np.uint64(hc)
        hc = (hc * _MULT_B) & 0xFFFFFFFF
        v = (v * np.uint64(hc)) & _M32
        out32[:, i] = v ^ (v >> np.uint64(16))
    return out32[:, 0::2] | (out32[:, 1::2] << np.uint64(32))  # little-endian uint32 pairs -> uint64 [N,4]


def pcg64_initial_state(seeds):
    """(state uint64 [N,2], inc uint64 [N,2]) with column 0 = high word, column 1 = low word, equal to
    `np.random.PCG64(np.random.SeedSequence(seed)).state['state']` for every seed."""
    s4 = _generate_state_u64x4(seed_sequence_pool(seeds))
    n = s4.shape[0]
    state = np.zeros((n, 2), np.uint64)
    inc = np.zeros((n, 2), np.uint64)
    mult = (PCG_MULT_HI << 64) | PCG_MULT_LO
    for i in range(n):  # 128-bit arithmetic in Python ints (one-off, ~1 us per env)
        initstate = (int(s4[i, 0]) << 64) | int(s4[i, 1])
        initseq = (int(s4[i, 2]) << 64) | int(s4[i, 3])
        ic = ((initseq << 1) | 1) & _MASK128
        st = ic  # state = 0 -> step -> inc
        st = (st + initstate) & _MASK128
        st = (st * mult + ic) & _MASK128
        state[i] = (st >> 64, st & _MASK64)
        inc[i] = (ic >> 64, ic & _MASK64)
    return state, inc


def pcg64_next_uniform53(state, inc):
    """Advance uint64 [N,2] states in place (Python ints; test helper) and return the 53-bit integers m with
    Generator.random() == m * 2**-53."""
    mult = (PCG_MULT_HI << 64) | PCG_MULT_LO
    out = np.zeros(state.shape[0], np.uint64)
    for i in range(state.shape[0]):
        st = (int(state[i, 0]) << 64) | int(state[i, 1])
        ic = (int(inc[i, 0]) << 64) | int(inc[i, 1])
        st = (st * mult + ic) & _MASK128
        hi, lo = st >> 64, st & _MASK64
        x = hi ^ lo
        rot = st >> 122
        raw = ((x >> rot) | (x << ((64 - rot) & 63))) & _MASK64
        out[i] = raw >> 11
        state[i] = (hi, lo)
    return out


def thresholds53(p_success):
    """Integer thresholds T_k = ceil(cdf_k * 2**53), k = 0..2, and idx_cap, for the outcome vector [p, q, q, q]:
    numpy's searchsorted(cdf, m * 2**-53, 'right') == min(sum(m >= T_k), idx_cap)."""
    p = float(p_success)
    cdf = np.cumsum(np.array([p] + [(1.0 - p) / 3.0] * 3, np.float64))
    cdf /= cdf[-1]
    t = [int(math.ceil(float(cdf[k]) * 2.0 ** 53)) for k in range(3)]  # scaling by 2**53 is exact in binary64
    cap = sum(1 for v in t if v < 2 ** 53)
    return tuple(min(v, 2 ** 53) for v in t), cap
