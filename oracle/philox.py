"""TEST INFRASTRUCTURE ONLY (oracle) -- host twin of the in-kernel Philox4x32-10 generator.

The reference has no counter-based RNG (it draws from numpy's PCG64 / MT19937 streams, see
rl_envs_forge/envs/grid_world/grid_world.py:69,515 and k_armed_bandit.py:145).  The engine replaces
those streams by Philox4x32-10 keyed by (seed) and counted by (global env index, step, stream) so that
every draw is reproducible on the host; the parity tests then feed *the same* draws to the oracle /
the reference (injected through `np_random` shims).  This file restates the published Philox4x32-10
algorithm (Salmon et al., SC'11; Random123) in numpy and is pinned by the Random123 known-answer
vectors in tests/test_oracle_philox.py.

Stream layout shared with rl_envs_forge_b200/csrc/ef_common.cuh (keep both in sync):
    key = (seed & 0xffffffff, seed >> 32)
    ctr = (x, c_lo, c_hi, stream | block << 8)
  per-env streams    (RESET_*, BANDIT*):    x = g (global env index, u32), all four words belong to env g
  grouped streams    (TRANSITION, POLICY):  x = g >> 2, env g owns word g & 3 -- one Philox call serves four
                                            consecutive envs, which is what a thread of the step kernels processes
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)

# stream ids (ctr[3] low byte)
STREAM_TRANSITION = 0   # grouped, per (g>>2, t): word g&3 = transition (outcome) draw of env g at step t
STREAM_RESET_AUTO = 1   # per (env, t): initial-state draws of the episode that starts after step t
STREAM_RESET_USER = 2   # per (env, reset epoch): initial-state draws of an explicit reset()
STREAM_BANDIT = 3       # per (env, t), block j: primitive draws of the arm sampler
STREAM_BANDIT_MEANS = 4  # per (env, arm): default-arm means
STREAM_POLICY = 5       # grouped, per (g>>2, t): word g&3 = policy draw (random action) of env g at step t


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable, values taken mod 2**32.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def draw_block(seed, g, counter, stream, block=0):
    """4 uint32 words for (seed, global env index g, 64-bit counter, stream, block)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    counter = np.asarray(counter, dtype=np.uint64)
    c1 = counter & MASK32
    c2 = counter >> np.uint64(32)
    c3 = np.uint64((int(stream) & 0xFF) | ((int(block) & 0xFFFFFF) << 8))
    return philox4x32_10(np.asarray(g, dtype=np.uint64), c1, c2, c3, seed & 0xFFFFFFFF, seed >> 32)


def group_word(seed, g, t, stream):
    """Word of env g in a grouped stream: word (g & 3) of the block at x = g >> 2, counter t."""
    g = np.asarray(g, dtype=np.uint64)
    w = np.stack(draw_block(seed, g >> np.uint64(2), t, stream), axis=-1)
    lane = (g & np.uint64(3)).astype(np.int64)
    return np.take_along_axis(w, lane[..., None], axis=-1)[..., 0]


def step_words(seed, g, t):
    """(transition word, policy word) of step index t for envs g."""
    return group_word(seed, g, t, STREAM_TRANSITION), group_word(seed, g, t, STREAM_POLICY)


def u23(w):
    """uint32 word -> float64 uniform in (0,1), exactly representable in fp32: ((w>>9)+0.5) * 2**-23."""
    return ((np.asarray(w, dtype=np.uint32) >> np.uint32(9)).astype(np.float64) + 0.5) * (2.0 ** -23)


def uniform_f32(w, lo, hi):
    """Device rule for a real uniform: fp32 `lo + (hi-lo)*u23(w)` with separate (non-fused) mul and add."""
    u = u23(w).astype(np.float32)
    lo = np.float32(lo)
    span = np.float32(np.float32(hi) - lo)
    return (u * span + lo).astype(np.float32)
