"""GPU parity tests for mixed-layout GridWorld batches: one launch steps envs that live on
different grids (walls / terminals / special transitions / start state per layout).  Checked bit-exactly against the oracle
(oracle/gridworld.py, pinned to the reference's golden trajectories) applied layout by layout."""
import numpy as np
import pytest

from oracle import philox
from oracle.gridworld import FlatTables, GridWorldPort

pytestmark = pytest.mark.gpu


def _random_layouts(rng, m, rows, cols):
    cells = [(r, c) for r in range(rows) for c in range(cols)]
    out = []
    for _ in range(m):
        pick = rng.permutation(len(cells))
        walls = {cells[i] for i in pick[:rng.integers(0, rows)]}
        free = [cells[i] for i in pick[rows:]]
        terms = {free[0]: float(rng.integers(1, 5)), free[1]: -1.0}
        specials = {(free[2], int(rng.integers(0, 4))): (free[3], 0.25 * int(rng.integers(-4, 5)))}
        out.append(dict(start_state=free[4], terminal_states=terms, walls=walls, special_transitions=specials))
    return out


@pytest.mark.parametrize("rows,cols,p,m,explicit", [(8, 8, 0.8, 7, False), (8, 8, 1.0, 3, True), (40, 40, 0.7, 5, True)])
def test_mixed_layout_batch_matches_oracle(rows, cols, p, m, explicit):
    """(40x40 x 5 layouts = 1.0 MB of tables: the global-memory table path; 8x8: all layouts staged in shared memory.)"""
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorldLayouts
    from oracle.gridworld import vec_step
    rng = np.random.default_rng(rows + m)
    layouts = _random_layouts(rng, m, rows, cols)
    n, seed, off, limit = 30_011, 21, 777, 30
    li = rng.integers(0, m, n) if explicit else (off + np.arange(n)) % m
    vec_layouts = [dict(l, special_transitions={(k[0], Action(k[1])): v for k, v in l["special_transitions"].items()})
                   for l in layouts]
    env = VecGridWorldLayouts(n, vec_layouts, rows=rows, cols=cols, p_success=p, episode_length_limit=limit,
                              layout_index=li if explicit else None, seed=seed, env_offset=off)
    assert np.array_equal(env.layout_of_env(), li)
    ports = [GridWorldPort(rows=rows, cols=cols, p_success=p, episode_length_limit=limit, **l) for l in layouts]
    tabs = [FlatTables(pt) for pt in ports]
    starts = np.array([pt.start_state[0] * cols + pt.start_state[1] for pt in ports])
    pos, ln, ret = starts[li].astype(np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
    assert np.array_equal(env.pos.cpu().numpy(), pos)
    g = np.arange(n) + off
    errors = episodes = 0
    for t in range(45):
        wt = philox.group_word(seed, g, t, philox.STREAM_TRANSITION)
        act = rng.integers(0, 4, n).astype(np.int32)
        act[rng.random(n) < 0.005] = 9
        obs, rew, term, trunc, info = env.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        want = dict(obs=np.zeros((n, 2), np.int32), reward=np.zeros(n, np.float32), terminated=np.zeros(n, bool),
                    truncated=np.zeros(n, bool), final=np.zeros(n, np.int64))
        for l in range(m):
            sel = li == l
            o = vec_step(tabs[l], pos[sel], ln[sel], ret[sel], act[sel], wt[sel], limit=limit, start_pos=starts[l], autoreset=True)
            pos[sel], ln[sel], ret[sel] = o["pos"], o["ep_len"], o["ep_ret"]
            want["obs"][sel], want["reward"][sel] = o["obs"], o["reward"]
            want["terminated"][sel], want["truncated"][sel], want["final"][sel] = o["terminated"], o["truncated"], o["final_pos"]
            errors += int(o["error"].sum())
            episodes += int(o["finished"].sum())
        assert np.array_equal(obs.cpu().numpy(), want["obs"]), t
        assert np.array_equal(rew.cpu().numpy(), want["reward"]), t
        assert np.array_equal(term.cpu().numpy(), want["terminated"]) and np.array_equal(trunc.cpu().numpy(), want["truncated"]), t
        fo = info["final_observation"].cpu().numpy()
        assert np.array_equal(fo[:, 0] * cols + fo[:, 1], want["final"]), t
    assert np.array_equal(env.pos.cpu().numpy(), pos) and np.array_equal(env.ep_len.cpu().numpy(), ln)
    assert np.array_equal(env.ep_ret.cpu().numpy(), ret)
    st = env.episode_stats()
    assert st["episodes"] == episodes and int(env.err[0]) == errors and st["env_steps"] == 45 * n - errors
    assert episodes > 0 and errors > 0


def test_layout_exports_reset_and_equivalence_with_single_layout_env():
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld, VecGridWorldLayouts
    lay = dict(start_state=(0, 0), walls={(1, 1), (1, 2), (2, 5)}, terminal_states={(7, 7): 1.0, (0, 7): -1.0},
               special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
    other = dict(start_state=(2, 2), terminal_states={(5, 5): 2.0})
    n = 4099
    multi = VecGridWorldLayouts(n, [lay, other], rows=8, cols=8, p_success=0.8, episode_length_limit=20, seed=3,
                                layout_index=np.zeros(n, np.int64))
    single = VecGridWorld(n, rows=8, cols=8, p_success=0.8, episode_length_limit=20, seed=3, **lay)
    rng = np.random.default_rng(0)
    for _ in range(40):
        act = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
        for x, y in zip(multi.step(act)[:4], single.step(act)[:4]):
            assert torch.equal(x, y)
    assert multi.mdp_of(0) == single.mdp and np.array_equal(multi.P_of(0), single.P)
    assert multi.mdp_of(1)[((2, 2), Action.UP)][0][0] == (1, 2)
    multi2 = VecGridWorldLayouts(10, [lay, other], rows=8, cols=8)
    assert multi2.state.cpu().tolist() == [[0, 0], [2, 2]] * 5
    multi2.step(np.full(10, 2))                                    # DOWN
    assert multi2.state.cpu().tolist() == [[1, 0], [3, 2]] * 5
    multi2.reset(mask=torch.arange(10, device="cuda") < 2)
    assert multi2.state.cpu().tolist() == [[0, 0], [2, 2]] + [[1, 0], [3, 2]] * 4
    with pytest.raises(ValueError):
        VecGridWorldLayouts(4, [dict(rows=3)], rows=8, cols=8)


@pytest.mark.parametrize("rows,cols,p,m,explicit,n", [(8, 8, 0.8, 7, False, 30_011), (8, 8, 1.0, 3, True, 4096),
                                                       (40, 40, 0.7, 5, True, 10_002), (8, 8, 0.8, 4, True, 3)])
@pytest.mark.parametrize("external", [False, True])
@pytest.mark.parametrize("layout", ["packed", "split"])
def test_mixed_layout_rollout_equals_step_sequence(rows, cols, p, m, explicit, n, external, layout):
    """ef_gridworld_rollout_layouts (K fused steps, state in registers) against K calls of the mixed-layout step, which the
    test above pins to the oracle: same observations, rewards, flags, final state, statistics and error count -- with the
    in-kernel random policy (its host twin feeds step()) and with external actions that include invalid ones."""
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorldLayouts
    rng = np.random.default_rng(rows * 7 + m)
    layouts = _random_layouts(rng, m, rows, cols)
    vec_layouts = [dict(l, special_transitions={(k[0], Action(k[1])): v for k, v in l["special_transitions"].items()})
                   for l in layouts]
    seed, off, limit, K = 5, (4096 + 8 if n % 2 == 0 else 777), 25, 37
    li = rng.integers(0, m, n) if explicit else None
    mk = lambda: VecGridWorldLayouts(n, vec_layouts, rows=rows, cols=cols, p_success=p, episode_length_limit=limit,  # noqa: E731
                                     layout_index=li, seed=seed, env_offset=off, state_layout=layout)
    a, b = mk(), mk()
    # a few plain steps first so that the rollout starts from a mixed state and a non-zero step counter
    for _ in range(3):
        act = torch.from_numpy(rng.integers(0, 4, n).astype(np.int32)).cuda()
        a.step(act)
        b.step(act)
    g = np.arange(n) + off
    if external:
        acts = rng.integers(0, 4, (K, n)).astype(np.int32)
        acts[rng.random((K, n)) < 0.003] = 11
    else:
        acts = np.stack([(philox.group_word(seed, g, 3 + k, philox.STREAM_POLICY) >> np.uint32(30)).astype(np.int32)
                         for k in range(K)])
    d_acts = torch.from_numpy(acts).cuda()
    obs, rew, flg = a.rollout(K, d_acts if external else None, return_obs=True)
    for k in range(K):
        o, r, te, tr, _ = b.step(d_acts[k])
        assert torch.equal(obs[k], o), k
        assert torch.equal(rew[k], r), k
        assert torch.equal(flg[k] & 1, te.to(torch.uint8)) and torch.equal(flg[k] >> 1, tr.to(torch.uint8)), k
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
    assert sa[0] == sb[0] and sa[1] == sb[1] and sa[4] == sb[4] and np.allclose(sa[2:4], sb[2:4], rtol=1e-12)
    assert int(a.err[0]) == int(b.err[0])
    assert a.step_count == b.step_count == 3 + K
    # statistics-only rollout (nothing materialised) continues the same trajectories
    a.rollout(5, d_acts[:5] if external else None)
    rw, fl = b.rollout(5, d_acts[:5] if external else None, return_rewards=True)
    assert rw.shape == (5, n) and torch.equal(a.pos, b.pos) and torch.equal(a.ep_ret, b.ep_ret)
