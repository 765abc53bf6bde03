"""GPU parity tests for Labyrinth stepping: VecLabyrinth through the C ABI vs oracle/labyrinth.py
(pinned against the reference's own trajectories) and directly against the golden trajectories.  All integer / byte work:
bit-exact."""
import json

import numpy as np
import pytest

from oracle import labyrinth as OL
from tests._golden import golden_files, load

pytestmark = pytest.mark.gpu
FILES = golden_files("labyrinth")


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("labyrinth_")[-1][:-4])
def test_golden_steps_as_batch(path):
    """Every recorded step of the unmodified reference becomes one player on the reference-generated maze."""
    import torch
    from rl_envs_forge_b200 import VecLabyrinth
    g = load(path)
    kw = json.loads(str(g["config"]))
    n = len(g["action"])
    env = VecLabyrinth(n, g["grid"], g["start"], g["target"], state_vision_range=kw.get("state_vision_range"),
                       autoreset=False)
    assert np.array_equal(env.state[0].cpu().numpy(), g["initial_obs"])
    env.set_state(g["pre"])
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(g["action"].astype(np.int32)).cuda())
    assert np.array_equal(env.player_position.cpu().numpy(), g["post"])
    assert np.array_equal(rew.cpu().numpy(), g["reward"].astype(np.float32))
    assert np.array_equal(term.cpu().numpy(), g["done"]) and not bool(trunc.any())
    assert np.array_equal(obs.cpu().numpy(), g["obs"])
    assert np.array_equal(env.state.cpu().numpy(), g["obs"])


def _random_mazes(rng, m, rows, cols):
    grids = (rng.random((m, rows, cols)) > 0.3).astype(np.uint8)
    grids[rng.random((m, rows, cols)) < 0.05] = 3                  # START marks are walkable like any non-WALL value
    starts, targets = [], []
    for k in range(m):
        free = np.argwhere(grids[k] != 0)
        s, t = free[rng.choice(len(free), 2, replace=False)]
        starts.append(s)
        targets.append(t)
    return grids, np.array(starts), np.array(targets)


@pytest.mark.parametrize("rows,cols,vision,explicit_index", [(16, 16, None, False), (10, 10, None, True), (9, 7, None, False),
                                                             (12, 17, 2, True), (30, 30, 3, False), (11, 13, 9, False)])
def test_many_mazes_match_oracle(rows, cols, vision, explicit_index):
    """M mazes, N players, auto-reset + step limit + rejected actions; 16-, 4- and 1-byte render paths, register-resident
    and generic vision windows."""
    import torch
    from rl_envs_forge_b200 import VecLabyrinth
    rng = np.random.default_rng(rows * 100 + cols)
    m, n, off, limit = 37, 20_011, 5000, 25
    grids, starts, targets = _random_mazes(rng, m, rows, cols)
    mi = rng.integers(0, m, n) if explicit_index else (off + np.arange(n)) % m
    env = VecLabyrinth(n, grids, starts, targets, maze_index=mi if explicit_index else None, state_vision_range=vision,
                       episode_length_limit=limit, env_offset=off, seed=3)
    assert np.array_equal(env.maze_of_env(), mi)
    flat = grids.reshape(m, -1)
    start_c, tgt_c = starts[:, 0] * cols + starts[:, 1], targets[:, 0] * cols + targets[:, 1]
    pos = start_c[mi].copy()
    ln = np.zeros(n, np.int64)
    ret = np.zeros(n, np.float32)
    assert np.array_equal(env.state.cpu().numpy(), OL.vec_obs(flat, mi, tgt_c, pos, cols, vision))
    episodes = errors = 0
    for t in range(40):
        act = rng.integers(0, 4, n).astype(np.int32)
        act[rng.random(n) < 0.01] = 4 + t
        obs, rew, term, trunc, info = env.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        pos, rw, done, invalid = OL.vec_step(flat, mi, tgt_c, pos, act, cols)
        ln = ln + ~invalid
        ret = (ret + rw).astype(np.float32)
        tr = (ln >= limit) & ~invalid
        assert np.array_equal(info["final_observation"].cpu().numpy(), OL.vec_obs(flat, mi, tgt_c, pos, cols, vision)), t
        fin = done | tr
        episodes += int(fin.sum())
        errors += int(invalid.sum())
        pos = np.where(fin, start_c[mi], pos)
        ln = np.where(fin, 0, ln)
        ret = np.where(fin, np.float32(0), ret)
        assert np.array_equal(rew.cpu().numpy(), rw), t
        assert np.array_equal(term.cpu().numpy(), done) and np.array_equal(trunc.cpu().numpy(), tr), t
        assert np.array_equal(obs.cpu().numpy(), OL.vec_obs(flat, mi, tgt_c, pos, cols, vision)), t
    assert np.array_equal(env.pos.cpu().numpy(), pos) and np.array_equal(env.ep_len.cpu().numpy(), ln)
    assert np.array_equal(env.ep_ret.cpu().numpy(), ret)
    st = env.episode_stats()
    assert st["episodes"] == episodes and st["env_steps"] == 40 * n - errors and int(env.err[0]) == errors
    assert episodes > 0 and errors > 0


def test_reference_test_cases_on_vec_surface():
    """tests/envs/labyrinth/test_labyrinth.py: predetermined layout (:296-314), valid / invalid move (:166-191), reaching
    the target (:193-203), vision windows (:316-420), set_state errors (:75-104)."""
    from rl_envs_forge_b200 import Action, VecLabyrinth
    W, P = OL.WALL, OL.PATH
    pre = np.array([[W, W, W, W, W], [W, P, P, P, W], [W, P, W, P, W], [W, P, P, P, W], [W, W, W, W, W]], np.uint8)
    lab = VecLabyrinth(1, pre, [(1, 1)], [(3, 3)], autoreset=False, strict=True)
    assert lab.player_position.cpu().tolist() == [[1, 1]]
    _, r, d, _, _ = lab.step(np.array([Action.RIGHT]))
    assert lab.player_position.cpu().tolist() == [[1, 2]] and r[0] == np.float32(-0.01) and not d[0]
    _, r, d, _, _ = lab.step(np.array([Action.UP]))
    assert lab.player_position.cpu().tolist() == [[1, 2]] and r[0] == -1
    lab.set_state((2, 3))
    s, r, d, _, _ = lab.step(np.array([Action.DOWN]))
    assert r[0] == 10 and d[0] and s[0, 3, 3] == OL.PLAYER
    with pytest.raises(ValueError):
        lab.set_state((2, 2))                                       # wall
    with pytest.raises(ValueError):
        lab.set_state((3, 3))                                       # target
    with pytest.raises(ValueError):
        lab.step(np.array([7]))                                     # Action(7)
    with pytest.raises(AssertionError, match="state_vision_range must be a positive integer or None"):
        VecLabyrinth(1, pre, [(1, 1)], [(3, 3)], state_vision_range=-1)
    # vision 3 in the middle of a 10x10 maze / vision 2 in a corner
    pre10 = np.zeros((10, 10), np.uint8)
    pre10[3:7, 3:7] = P
    lab = VecLabyrinth(1, pre10, [(5, 5)], [(6, 6)], state_vision_range=3)
    s = lab.state[0].cpu().numpy()
    assert s.shape == (7, 7) and s[3, 3] == OL.PLAYER and s[4, 4] == OL.TARGET
    assert np.array_equal(s, OL.LabyrinthPort(pre10, (5, 5), (6, 6), 3).state)
    pre10 = np.zeros((10, 10), np.uint8)
    pre10[0:3, 0:3] = P
    lab = VecLabyrinth(1, pre10, [(0, 0)], [(2, 2)], state_vision_range=2)
    s = lab.state[0].cpu().numpy()
    assert s.shape == (5, 5) and s[4, 4] == OL.TARGET and (s[:2] == W).all() and (s[:, :2] == W).all()


def test_host_buffer_path_from_mazes_and_snapshots():
    import torch
    from types import SimpleNamespace
    from rl_envs_forge_b200 import VecLabyrinth
    rng = np.random.default_rng(8)
    grids, starts, targets = _random_mazes(rng, 6, 20, 20)
    mazes = [SimpleNamespace(grid=grids[k].astype(int), start_position=tuple(starts[k]), target_position=tuple(targets[k]))
             for k in range(6)]                                    # what the reference's Labyrinth.maze exposes
    n = 3001
    a = VecLabyrinth.from_mazes(mazes, n, seed=1)
    b = VecLabyrinth(n, grids, starts, targets, seed=1)
    for t in range(15):
        act = rng.integers(0, 4, n).astype(np.int32)
        oa = a.step(act)                                           # numpy in -> numpy out
        ob = b.step(torch.from_numpy(act).cuda())
        assert isinstance(oa[0], np.ndarray) and oa[0].shape == (n, 20, 20) and oa[0].dtype == np.uint8
        for x, y in zip(oa[:4], ob[:4]):
            assert np.array_equal(x, y.cpu().numpy())
    snap = b.get_state()
    before = b.state.clone()
    b.step(torch.zeros(n, dtype=torch.int32, device="cuda"), return_obs=False)
    b.set_state(snap)
    assert torch.equal(b.state, before)
    obs, _ = b.reset(mask=torch.arange(n, device="cuda") < 10)
    assert torch.equal(b.pos[:10].cpu(), torch.from_numpy((starts[:, 0] * 20 + starts[:, 1]).astype(np.int32))[np.arange(10) % 6])


def test_rollout_equals_step_sequence():
    """K fused steps (no rendering) == K step() calls: external actions with emitted rewards / flags, then the in-kernel
    random policy through its host twin (policy word >> 30)."""
    import torch
    from oracle import philox
    from rl_envs_forge_b200 import VecLabyrinth
    rng = np.random.default_rng(12)
    grids, starts, targets = _random_mazes(rng, 9, 14, 11)
    n, K, seed, off = 50_003, 48, 6, 300
    kw = dict(episode_length_limit=20, seed=seed, env_offset=off)
    a, b = VecLabyrinth(n, grids, starts, targets, **kw), VecLabyrinth(n, grids, starts, targets, **kw)
    acts = torch.randint(0, 4, (K, n), dtype=torch.int32, device="cuda")
    acts[5, ::77] = 6                                                # rejected actions
    rew, flg = a.rollout(K, acts, return_rewards=True)
    for k in range(K):
        o = b.step(acts[k], return_obs=False)
        assert torch.equal(rew[k], o[1]) and torch.equal(flg[k], o[2].to(torch.uint8) | (o[3].to(torch.uint8) << 1)), k
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    assert int(a.err[0]) == int(b.err[0]) > 0
    g = np.arange(n) + off
    a.rollout(K)
    for k in range(K):
        wp = philox.group_word(seed, g, K + k, philox.STREAM_POLICY)
        b.step(torch.from_numpy((wp >> np.uint32(30)).astype(np.int32)).cuda(), return_obs=False)
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    assert torch.equal(a.state, b.state)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] > 0 and sa["length_sum"] == sb["length_sum"] and sa["env_steps"] == sb["env_steps"]


@pytest.mark.parametrize("n", [1, 2, 31, 255, 257])
def test_ragged_sizes_and_sharding(n):
    """Partial CTAs of the 256-players-per-CTA kernel; a shard (env_offset) reproduces its slice of the whole batch."""
    import torch
    from rl_envs_forge_b200 import VecLabyrinth
    rng = np.random.default_rng(n)
    grids, starts, targets = _random_mazes(rng, 5, 10, 10)
    off = 1000
    for vision in (None, 2):
        big = VecLabyrinth(off + 600, grids, starts, targets, state_vision_range=vision, seed=3)
        small = VecLabyrinth(n, grids, starts, targets, state_vision_range=vision, seed=3, env_offset=off)
        for t in range(8):
            act = torch.randint(0, 4, (off + 600,), dtype=torch.int32, device="cuda")
            ob, os_ = big.step(act), small.step(act[off:off + n].clone())
            for x, y in zip(ob[:4], os_[:4]):
                assert torch.equal(x[off:off + n], y), (n, vision, t)
        big.rollout(7)
        small.rollout(7)
        assert torch.equal(big.pos[off:off + n], small.pos) and torch.equal(big.state[off:off + n], small.state)
