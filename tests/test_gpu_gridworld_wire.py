"""GPU parity tests of the narrow wire format (`wire="compact"` -> ef_gridworld_step_u8): uint8 actions in, uint8
(row, col) observations out.  Same trajectories as the canonical int32 call and as the oracle, bit for bit."""
import ctypes

import numpy as np
import pytest

from oracle import philox
from oracle.gen_golden import CONFIG2_LAYOUT
from oracle.gridworld import FlatTables, GridWorldPort, vec_step
from tests._golden import golden_files, grid_kwargs, grid_runtime_specials, load

pytestmark = pytest.mark.gpu
FILES = golden_files("gridworld")


def _vec_kwargs(kw):
    from rl_envs_forge_b200 import Action
    kw = dict(kw)
    if kw.get("special_transitions"):
        kw["special_transitions"] = {(k[0], Action(int(k[1]))): v for k, v in kw["special_transitions"].items()}
    return kw


def _config2(p, limit=200):
    return dict(CONFIG2_LAYOUT, p_success=p, episode_length_limit=limit)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_golden_steps_as_compact_batch(path):
    """Every recorded reference step as one env of a compact-wire batch (injected draws: one env per thread kernel)."""
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld
    g = load(path)
    kw = grid_kwargs(g)
    n = len(g["action"])
    env = VecGridWorld(n, autoreset=False, wire="compact", **_vec_kwargs(kw))
    for fs, a, ts, r in grid_runtime_specials(g):
        env.add_special_transition(fs, Action(a), ts, r)
    env.state = g["pre"]
    raised = g["raised"] > 0
    env.ep_len = torch.from_numpy((g["ep_len"] - (~raised)).astype(np.int32))
    r32 = np.floor(g["u"] * 2.0 ** 32).astype(np.uint64).astype(np.uint32)
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(g["action"].astype(np.uint8)).cuda(),
                                        draws=torch.from_numpy(r32.view(np.int32)).cuda())
    assert obs.dtype == torch.uint8 and tuple(obs.shape) == (n, 2)
    obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
    ok = ~raised
    assert np.array_equal(obs[ok], g["post"][ok])
    assert np.array_equal(obs[raised], g["pre"][raised])
    assert np.array_equal(rew[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(term[ok], g["done"][ok]) and np.array_equal(trunc[ok], g["trunc"][ok])
    assert int(env.err.cpu().numpy()[0]) == int(raised.sum())


@pytest.mark.parametrize("p,n", [(0.8, 65536), (1.0, 65536), (0.8, 4099), (0.8, 3), (0.8, 1 << 20)])
def test_compact_philox_step_matches_oracle(p, n):
    """Device-side Philox draws, auto-reset, vector kernel (and its ragged tail): bit-exact against the restatement."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    steps = 24 if n >= 1 << 20 else 200
    limit, seed = 50, 99
    cfg = _config2(p, limit)
    tab = FlatTables(GridWorldPort(**cfg))
    env = VecGridWorld(n, seed=seed, wire="compact", **_vec_kwargs(cfg))
    pos, ln, ret = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
    g = np.arange(n)
    eps = errors = 0
    for t in range(steps):
        wt, wp = philox.step_words(seed, g, np.full(n, t, np.uint64))
        act = (wp >> np.uint32(30)).astype(np.uint8)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        o = vec_step(tab, pos, ln, ret, act.astype(np.int32), wt, limit=limit, start_pos=0, autoreset=True)
        eps += int(o["finished"].sum())
        errors += int(o["error"].sum())
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
        assert obs.dtype == torch.uint8
        assert np.array_equal(obs.cpu().numpy(), o["obs"]), t
        assert np.array_equal(rew.cpu().numpy(), o["reward"]), t
        assert np.array_equal(term.cpu().numpy(), o["terminated"]) and np.array_equal(trunc.cpu().numpy(), o["truncated"]), t
    assert np.array_equal(env.pos.cpu().numpy(), pos) and np.array_equal(env.ep_len.cpu().numpy(), ln)
    assert np.array_equal(env.ep_ret.cpu().numpy(), ret)
    st = env.episode_stats()
    assert int(st["episodes"]) == eps and int(st["env_steps"]) == n * steps - errors
    assert int(env.err.cpu().numpy()[0]) == errors


@pytest.mark.parametrize("layout,autoreset", [("packed", True), ("split", True), ("split", False)])
def test_compact_equals_canonical(layout, autoreset):
    """Same seed and actions through both wire formats, including invalid actions (4, 7, 200, 255 and, on the canonical
    side, their int32 twins): observations, rewards, flags, final observations, state, statistics and error words agree."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n = 20_003
    cfg = _vec_kwargs(_config2(0.8, 40))
    a = VecGridWorld(n, seed=11, autoreset=autoreset, state_layout=layout, **cfg)
    b = VecGridWorld(n, seed=11, autoreset=autoreset, state_layout=layout, wire="compact", **cfg)
    assert torch.equal(a.reset()[0], b.reset()[0].to(torch.int32))
    rng = np.random.default_rng(3)
    for t in range(120):
        act = rng.integers(0, 4, n).astype(np.uint8)
        bad = rng.random(n) < 0.002
        act[bad] = rng.choice(np.array([4, 7, 200, 255], np.uint8), int(bad.sum()))
        fo = t % 3 == 0
        oa = a.step(torch.from_numpy(act.astype(np.int32)).cuda(), return_final_obs=fo)
        ob = b.step(torch.from_numpy(act).cuda(), return_final_obs=fo)
        assert ob[0].dtype == torch.uint8
        assert torch.equal(oa[0], ob[0].to(torch.int32)) and torch.equal(oa[1], ob[1])
        assert torch.equal(oa[2], ob[2]) and torch.equal(oa[3], ob[3])
        if fo:
            assert ob[4]["final_observation"].dtype == torch.uint8
            assert torch.equal(oa[4]["final_observation"], ob[4]["final_observation"].to(torch.int32))
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
    assert sa[0] == sb[0] and sa[1] == sb[1] and sa[4] == sb[4]      # episodes, length sum, env-steps: integers, exact
    assert np.allclose(sa[2:4], sb[2:4], rtol=1e-12)                 # float sums: atomics in a different order
    ea, eb = a.err.cpu().numpy(), b.err.cpu().numpy()
    assert int(ea[0]) == int(eb[0]) > 0
    with pytest.raises(ValueError):
        b.check_errors()


@pytest.mark.parametrize("transport", ["zerocopy", "staged", "pipelined", "pipelined+h2d"])
def test_compact_host_buffer_path_equals_device_path(transport):
    """numpy / pinned uint8 actions in -> numpy out (uint8 observations); wider integer dtypes are narrowed on the host
    with out-of-range values kept invalid."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(0.8))
    n = 10_001
    a = VecGridWorld(n, seed=4, wire="compact", **cfg)
    a.host_transport, a.host_stage_actions, a.host_chunks = transport.split("+")[0], transport.endswith("+h2d"), 3
    b = VecGridWorld(n, seed=4, wire="compact", **cfg)
    rng = np.random.default_rng(0)
    for i in range(24):
        act = rng.integers(0, 4, n).astype(np.uint8)
        if i % 4 == 1:
            src = torch.from_numpy(act).pin_memory()
        elif i % 4 == 2:
            src = act.astype(np.int64)
            src[5], act[5] = 256, 255        # 256 must not wrap to action 0
            src[6], act[6] = -3, 255
        elif i % 4 == 3:
            src = torch.from_numpy(act.astype(np.int32))
        else:
            src = act
        oa = a.step(src, return_final_obs=True)
        ob = b.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        assert isinstance(oa[0], np.ndarray) and oa[0].dtype == np.uint8 and oa[0].shape == (n, 2)
        assert oa[1].dtype == np.float32 and oa[2].dtype == np.bool_
        assert np.array_equal(oa[0], ob[0].cpu().numpy()) and np.array_equal(oa[1], ob[1].cpu().numpy())
        assert np.array_equal(oa[2], ob[2].cpu().numpy()) and np.array_equal(oa[3], ob[3].cpu().numpy())
        assert np.array_equal(oa[4]["final_observation"], ob[4]["final_observation"].cpu().numpy())
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_ret, b.ep_ret)
    # 12 narrowed invalid actions + the env-steps on the cell the reference raises on (p_success < 1 special transition)
    assert int(a.err.cpu().numpy()[0]) == int(b.err.cpu().numpy()[0]) >= 12


@pytest.mark.parametrize("wire,transport", [("compact", "zerocopy"), ("compact", "staged"), ("canonical", "staged"),
                                            ("canonical", "zerocopy")])
def test_step_async_two_sets_in_flight_equal_plain_steps(wire, transport):
    """step_async() / wait(): two env sets stepped alternately with both steps in flight give, set by set, exactly what
    plain step() calls with CUDA actions give (same kernels; only the host wait moved)."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(0.8))
    n = 20_003
    adt = np.uint8 if wire == "compact" else np.int32
    sets = [VecGridWorld(n, seed=9, wire=wire, env_offset=k * n, **cfg) for k in range(2)]
    refs = [VecGridWorld(n, seed=9, wire=wire, env_offset=k * n, **cfg) for k in range(2)]
    for e in sets:
        e.async_transport = transport
    rng = np.random.default_rng(1)
    acts = [[torch.from_numpy(rng.integers(0, 4, n).astype(adt)).pin_memory() for _ in range(12)] for _ in range(2)]
    with pytest.raises(RuntimeError):
        sets[0].wait()
    sets[0].step_async(acts[0][0])
    with pytest.raises(RuntimeError):
        sets[0].step_async(acts[0][0])
    for i in range(24):
        k, t = i % 2, i // 2
        if i + 1 < 24:
            sets[(i + 1) % 2].step_async(acts[(i + 1) % 2][(i + 1) // 2])      # the other set's step is now in flight too
        got = sets[k].wait()
        want = refs[k].step(acts[k][t].cuda())
        assert got[0].dtype == adt and got[0].shape == (n, 2)
        assert np.array_equal(got[0], want[0].cpu().numpy()) and np.array_equal(got[1], want[1].cpu().numpy())
        assert np.array_equal(got[2], want[2].cpu().numpy()) and np.array_equal(got[3], want[3].cpu().numpy())
    for e, r in zip(sets, refs):
        assert torch.equal(e.pos, r.pos) and torch.equal(e.ep_ret, r.ep_ret)
        assert torch.equal(e.stats.cpu(), r.stats.cpu())


def test_compact_unaligned_buffers_through_the_c_abi():
    """Action / observation pointers that miss the 4- / 8-byte alignment of the vector kernel take the one-env-per-thread
    kernel; results equal the aligned call."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld, _lib
    lib = _lib.load()
    cfg = _vec_kwargs(_config2(0.8))
    n = 5000
    a = VecGridWorld(n, seed=8, wire="compact", **cfg)
    b = VecGridWorld(n, seed=8, wire="compact", **cfg)
    act_big = torch.zeros(n + 1, dtype=torch.uint8, device="cuda")
    obs_big = torch.zeros(2 * n + 2, dtype=torch.uint8, device="cuda")
    rew = torch.zeros(n, dtype=torch.float32, device="cuda")
    term = torch.zeros(n, dtype=torch.uint8, device="cuda")
    trunc = torch.zeros(n, dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(1)
    for t in range(40):
        act = torch.from_numpy(rng.integers(0, 4, n).astype(np.uint8)).cuda()
        oa = a.step(act)
        act_big[1:].copy_(act)
        st = lib.ef_gridworld_step_u8(ctypes.byref(b._desc), *b._state_ptrs(), b.ep_ret.data_ptr(),
                                      act_big.data_ptr() + 1, None, b._seed, t, 0, obs_big.data_ptr() + 2,
                                      rew.data_ptr(), term.data_ptr(), trunc.data_ptr(), None,
                                      b._stats_raw.data_ptr(), b.err.data_ptr(), None, n, 1,
                                      torch.cuda.current_stream().cuda_stream)
        assert st == 0
        assert torch.equal(oa[0], obs_big[2:].view(n, 2)) and torch.equal(oa[1], rew)
        assert torch.equal(oa[2].view(torch.uint8), term) and torch.equal(oa[3].view(torch.uint8), trunc)
    assert torch.equal(a.pos, b.pos)


def test_compact_wire_limits():
    """Grids wider than a byte are refused (constructor and C ABI); large grids within the limit use the global-table
    path."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld, _lib
    with pytest.raises(ValueError):
        VecGridWorld(8, rows=300, cols=4, wire="compact")
    with pytest.raises(ValueError):
        VecGridWorld(8, wire="compact", rng="pcg64", seed=1)
    with pytest.raises(ValueError):
        VecGridWorld(8, wire="tiny")
    wide = VecGridWorld(16, rows=300, cols=4, terminal_states={(299, 3): 1.0}, episode_length_limit=10)
    z = torch.zeros(64, dtype=torch.uint8, device="cuda")
    st = _lib.load().ef_gridworld_step_u8(ctypes.byref(wide._desc), *wide._state_ptrs(), wide.ep_ret.data_ptr(),
                                          z.data_ptr(), None, 0, 0, 0, z.data_ptr(), None, None, None, None, None, None,
                                          None, 16, 1, None)
    assert st == 3   # EF_ERR_UNSUPPORTED
    # 256 x 200: table beyond shared memory, row indices up to 255
    n = 8192
    kw = dict(rows=256, cols=200, terminal_states={(255, 199): 1.0}, start_state=(250, 190), p_success=0.9,
              episode_length_limit=30)
    a = VecGridWorld(n, seed=2, **kw)
    b = VecGridWorld(n, seed=2, wire="compact", **kw)
    rng = np.random.default_rng(5)
    for t in range(60):
        act = rng.integers(0, 4, n).astype(np.uint8)
        oa = a.step(torch.from_numpy(act.astype(np.int32)).cuda())
        ob = b.step(torch.from_numpy(act).cuda())
        assert torch.equal(oa[0], ob[0].to(torch.int32)) and torch.equal(oa[1], ob[1]) and torch.equal(oa[2], ob[2])
    assert int(b.state.max()) >= 250


def test_compact_cuda_graph_replay_equals_canonical_eager_steps():
    """The lean step() path of the narrow format (one struct argument, `wire` flag) captured in a CUDA graph with the
    device step clock: replays equal the eager canonical run."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, seed, per_graph, replays = 70_001, 31, 5, 6
    cfg = _vec_kwargs(_config2(0.8))
    eager = VecGridWorld(n, seed=seed, **cfg)
    graphed = VecGridWorld(n, seed=seed, wire="compact", **cfg)
    acts = [torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda") for _ in range(per_graph)]
    acts32 = [a.to(torch.int32) for a in acts]
    graphed.step(acts[0])
    eager.step(acts32[0])
    graphed.enable_device_clock()
    stream = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for a in acts:
                graphed.step(a)
    torch.cuda.synchronize()
    for _ in range(replays):
        g.replay()
        for a in acts32:
            out_e = eager.step(a)
    torch.cuda.synchronize()
    assert graphed.step_count == 1 + per_graph * replays == eager.step_count
    assert torch.equal(graphed.pos, eager.pos) and torch.equal(graphed.ep_len, eager.ep_len)
    assert torch.equal(graphed.ep_ret, eager.ep_ret)
    assert graphed._obs.dtype == torch.uint8 and torch.equal(graphed._obs.to(torch.int32), out_e[0])
    assert torch.equal(graphed._reward, out_e[1])
    sg, se = graphed.episode_stats(), eager.episode_stats()
    assert sg["episodes"] == se["episodes"] > 0 and sg["env_steps"] == se["env_steps"]


def test_compact_refuses_off_grid_start_states():
    from rl_envs_forge_b200 import VecGridWorld
    with pytest.raises(ValueError, match="inside the grid"):
        VecGridWorld(4, start_state=(9, 9), wire="compact")
    env = VecGridWorld(4, wire="compact")
    with pytest.raises(ValueError, match="inside the grid"):
        env.reset(new_start_state=(7, 0))
    obs, _ = env.reset(new_start_state=(2, 3))
    assert obs.cpu().tolist() == [[2, 3]] * 4 and env.start_state == (2, 3)
