"""GPU parity tests for GridWorld: the CUDA path (through the C ABI) against the reference's golden vectors and the
oracle restatement, bit-exact (integer state, fp32-cast rewards)."""
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


def _config2(p):
    cfg = dict(CONFIG2_LAYOUT, p_success=p)
    return cfg


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_golden_steps_as_batch(path):
    """Every recorded reference step becomes one env of a batch: same pre-state, action, injected draw."""
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld
    g = load(path)
    n = len(g["action"])
    env = VecGridWorld(n, autoreset=False, **_vec_kwargs(grid_kwargs(g)))
    for fs, a, ts, r in grid_runtime_specials(g):
        env.add_special_transition(fs, Action(a), ts, r)
    env.state = g["pre"]
    raised = g["raised"] > 0
    env.ep_len = torch.from_numpy((g["ep_len"] - (~raised)).astype(np.int32))
    r32 = np.floor(g["u"] * 2.0 ** 32).astype(np.uint64).astype(np.uint32)
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(g["action"].astype(np.int32)).cuda(),
                                        draws=torch.from_numpy(r32.view(np.int32)).cuda())
    obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
    ok = ~raised
    assert np.array_equal(obs[ok], g["post"][ok])
    assert np.array_equal(obs[raised], g["pre"][raised])          # the reference raises before mutating state
    assert np.array_equal(rew[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(term[ok], g["done"][ok]) and np.array_equal(trunc[ok], g["trunc"][ok])
    assert np.array_equal(env.ep_len.cpu().numpy()[ok], g["ep_len"][ok])
    err = env.err.cpu().numpy()
    assert int(err[0]) == int(raised.sum())
    if raised.any():
        with pytest.raises(ValueError):
            env.check_errors()


@pytest.mark.parametrize("name", ["default", "slip_seed42", "config2_p08", "limit10", "runtime_special"])
def test_single_env_strict_replay(name):
    """num_envs=1, autoreset off, strict: the wrapper reproduces the reference trajectory call for call, including
    the ValueError on cells the reference cannot execute."""
    from rl_envs_forge_b200 import Action, VecGridWorld
    g = load([f for f in FILES if f.endswith(f"gridworld_{name}.npz")][0])
    env = VecGridWorld(1, autoreset=False, strict=True, **_vec_kwargs(grid_kwargs(g)))
    for fs, a, ts, r in grid_runtime_specials(g):
        env.add_special_transition(fs, Action(a), ts, r)
    env.reset()
    steps = min(400, len(g["action"]))
    for i in range(steps):
        assert tuple(env.state.cpu().numpy()[0]) == tuple(g["pre"][i])
        r32 = np.uint32(np.floor(g["u"][i] * 2.0 ** 32))
        if g["raised"][i]:
            with pytest.raises(ValueError):
                env.step(np.array([g["action"][i]]), draws=np.array([r32]))
            continue
        obs, rew, term, trunc, _ = env.step(np.array([g["action"][i]]), draws=np.array([r32]))
        assert tuple(obs[0]) == tuple(g["post"][i]) and rew[0] == np.float32(g["reward"][i])
        assert bool(term[0]) == bool(g["done"][i]) and bool(trunc[0]) == bool(g["trunc"][i])
        if g["done"][i] or g["trunc"][i]:
            env.reset()


def _run_against_oracle(n, steps, p, seed, limit=200, check_every=1, env_offset=0):
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _config2(p)
    cfg["episode_length_limit"] = limit
    port = GridWorldPort(**cfg)
    tab = FlatTables(port)
    env = VecGridWorld(n, seed=seed, env_offset=env_offset, **_vec_kwargs(cfg))
    pos, ln, ret = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
    g = np.arange(n) + env_offset
    eps = lens = 0
    rets = 0.0
    errors = 0
    for t in range(steps):
        wt, wp = philox.step_words(seed, g, np.full(n, t, np.uint64))
        act = (wp >> np.uint32(30)).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        o = vec_step(tab, pos, ln, ret, act, wt, limit=limit, start_pos=0, autoreset=True)
        fin = o["finished"]
        eps += int(fin.sum()); lens += int(o["final_len"][fin].sum()); rets += float(o["final_ret"][fin].astype(np.float64).sum())
        errors += int(o["error"].sum())
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
        if t % check_every == 0 or t == steps - 1:
            assert np.array_equal(obs.cpu().numpy(), o["obs"]), t
            assert np.array_equal(rew.cpu().numpy(), o["reward"]), t
            assert np.array_equal(term.cpu().numpy(), o["terminated"]), t
            assert np.array_equal(trunc.cpu().numpy(), o["truncated"]), t
    assert np.array_equal(env.pos.cpu().numpy(), pos)
    assert np.array_equal(env.ep_len.cpu().numpy(), ln)
    assert np.array_equal(env.ep_ret.cpu().numpy(), ret)
    st = env.episode_stats()
    assert int(st["episodes"]) == eps and int(st["length_sum"]) == lens
    assert st["return_sum"] == pytest.approx(rets, rel=1e-12, abs=1e-9)
    assert int(st["env_steps"]) == n * steps - errors
    assert int(env.err.cpu().numpy()[0]) == errors
    return env, (pos, ln, ret), (eps, lens, rets)


@pytest.mark.parametrize("p", [1.0, 0.8])
def test_philox_step_matches_oracle_65536_envs(p):
    """Config 2 (8x8, walls, special transition), device-side Philox outcome draws, auto-reset: bit-exact per step.
    With p=0.8 the cell ((3,3), UP) is one the reference raises on; it is flagged and left untouched."""
    _run_against_oracle(65536, 256, p, seed=2024, check_every=1)


def test_full_size_config2_1M_envs():
    """BASELINE config 2 at full size (2^20 envs), 48 steps, both variants, bit-exact against the restatement."""
    _run_against_oracle(1 << 20, 48, 0.8, seed=5, check_every=8)
    _run_against_oracle(1 << 20, 24, 1.0, seed=6, check_every=8)


@pytest.mark.parametrize("p", [1.0, 0.8])
def test_rollout_equals_step_sequence(p):
    """Fused K-step rollout (random Philox policy) == K step() calls fed the same Philox policy words: identical state and
    identical integer statistics (size-independent property, run at 2^20 envs)."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, K, seed = 1 << 20, 64, 99
    cfg = _vec_kwargs(_config2(p))
    a = VecGridWorld(n, seed=seed, **cfg)
    b = VecGridWorld(n, seed=seed, **cfg)
    a.rollout(K)
    for t in range(K):
        b.step(None_actions(b, t))
    for x, y in ((a.pos, b.pos), (a.ep_len, b.ep_len), (a.ep_ret, b.ep_ret)):
        assert torch.equal(x, y)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] and sa["length_sum"] == sb["length_sum"] and sa["env_steps"] == sb["env_steps"]
    assert sa["return_sum"] == pytest.approx(sb["return_sum"], rel=1e-12)
    assert sa["episodes"] > 0
    # odd start time and odd K exercise the word-pairing rule
    a.rollout(7)
    a.rollout(5)
    for t in range(K, K + 12):
        b.step(None_actions(b, t))
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_ret, b.ep_ret)


def None_actions(env, t):
    """Policy word of step t for every env of `env`, as a CUDA int32 action tensor (host twin of the in-kernel policy)."""
    import torch
    g = np.arange(env.num_envs) + env.env_offset
    _, wp = philox.step_words(env._seed, g, np.full(env.num_envs, t, np.uint64))
    return torch.from_numpy((wp >> np.uint32(30)).astype(np.int32)).cuda()


def test_rollout_external_actions_and_emitted_outputs():
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, K, seed = 50_001, 33, 3
    cfg = _vec_kwargs(_config2(0.8))
    a = VecGridWorld(n, seed=seed, **cfg)
    b = VecGridWorld(n, seed=seed, **cfg)
    acts = torch.randint(0, 4, (K, n), dtype=torch.int32, device="cuda")
    rew, flg = a.rollout(K, acts, return_rewards=True)
    for t in range(K):
        _, r, te, tr, _ = b.step(acts[t])
        assert torch.equal(rew[t], r)
        assert torch.equal(flg[t], te.to(torch.uint8) | (tr.to(torch.uint8) << 1))
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    # in-kernel policy == stepping with None actions
    c = VecGridWorld(n, seed=seed, **cfg)
    d = VecGridWorld(n, seed=seed, **cfg)
    c.rollout(9)
    for t in range(9):
        d.step(None_actions(d, t))
    assert torch.equal(c.pos, d.pos)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 33, 257, 1023])
def test_ragged_sizes_and_tail(n):
    _run_against_oracle(n, 40, 0.8, seed=n, limit=7)


def test_sharding_invariance_env_offset():
    """Two shards with env_offset reproduce the single-launch result and statistics (Philox is keyed by global index)."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld, shard_range
    n, K, seed = 100_003, 50, 17
    cfg = _vec_kwargs(_config2(0.8))
    whole = VecGridWorld(n, seed=seed, **cfg)
    whole.rollout(K)
    parts = []
    for r in range(3):
        lo, hi = shard_range(n, r, 3)
        e = VecGridWorld(hi - lo, seed=seed, env_offset=lo, **cfg)
        e.rollout(K)
        parts.append(e)
    assert torch.equal(torch.cat([e.pos for e in parts]), whole.pos)
    assert torch.equal(torch.cat([e.ep_ret for e in parts]), whole.ep_ret)
    tot = sum(e.stats for e in parts)
    assert torch.equal(tot[:2], whole.stats[:2]) and tot[4] == whole.stats[4]
    assert float(tot[2]) == pytest.approx(float(whole.stats[2]), rel=1e-12)


def test_config5_full_size_64M_envs_sharding_invariance_and_conservation():
    """BASELINE configs[4] at its full size on one GPU: 2^26 GridWorld envs, fused 64-step rollout, stepped once as ONE
    batch and once as eight shards of 2^23 envs (what eight ranks own) -- identical state and identical integer statistics.
    Size-independent properties on top: env-steps executed + rejected = N * K, the finished episodes' lengths plus the
    counters of the episodes still running add up to the executed env-steps, and no counter reaches the limit."""
    import torch
    from rl_envs_forge_b200 import StatsReducer, VecGridWorld, shard_range
    n, K, seed = 1 << 26, 64, 5
    cfg = _vec_kwargs(_config2(0.8))
    whole = VecGridWorld(n, seed=seed, **cfg)
    whole.rollout(K)
    ws = whole.stats.cpu()
    rejected = int(whole.err[0].item()) & 0xFFFFFFFF
    assert int(ws[4]) + rejected == n * K and rejected > 0
    lens = whole.ep_len
    assert int(ws[1]) + int(lens.sum(dtype=torch.int64)) == int(ws[4])          # every executed step belongs to one episode
    assert int(lens.max()) < cfg["episode_length_limit"]
    pos_w, ret_w = whole.pos.clone(), whole.ep_ret.clone()
    del whole, lens
    shards, chunks = [], []
    for r in range(8):
        lo, hi = shard_range(n, r, 8)
        e = VecGridWorld(hi - lo, seed=seed, env_offset=lo, **cfg)
        e.rollout(K)
        assert torch.equal(e.pos, pos_w[lo:hi]) and torch.equal(e.ep_ret, ret_w[lo:hi])
        shards.append(e)
    tot = StatsReducer(shards).reduce().cpu()                                   # one fold kernel over the eight env sets
    assert int(tot[0]) == int(ws[0]) and int(tot[1]) == int(ws[1]) and int(tot[4]) == int(ws[4])
    assert float(tot[2]) == pytest.approx(float(ws[2]), rel=1e-12)


def test_unaligned_views_take_scalar_path():
    """State tensors that are not 16-byte aligned (views offset by one element) must still be correct."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n = 4099
    cfg = _vec_kwargs(_config2(0.8))
    a = VecGridWorld(n, seed=1, **cfg)
    b = VecGridWorld(n, seed=1, **cfg)
    for name in ("pos", "ep_len", "ep_ret"):
        big = torch.zeros(n + 1, dtype=getattr(b, name).dtype, device="cuda")
        setattr(b, name, big[1:])
    b.reset()
    for t in range(30):
        act = None_actions(a, t)
        oa = a.step(act)
        ob = b.step(act)
        assert all(torch.equal(x, y) for x, y in zip(oa[:4], ob[:4]))
    assert torch.equal(a.pos, b.pos)


def test_large_grid_uses_global_table_path():
    """200x200 grid: the outcome table (2.5 MB) exceeds shared memory, kernels read it through the read-only path."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    rng = np.random.default_rng(0)
    walls = {(int(r), int(c)) for r, c in rng.integers(1, 199, (3000, 2))}
    cfg = dict(rows=200, cols=200, start_state=(0, 0), walls=walls, terminal_states={(199, 199): 5.0},
               p_success=0.7, episode_length_limit=500)
    port = GridWorldPort(**cfg)
    tab = FlatTables(port)
    n, seed = 20_000, 8
    env = VecGridWorld(n, seed=seed, **cfg)
    pos, ln, ret = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
    g = np.arange(n)
    for t in range(60):
        wt, wp = philox.step_words(seed, g, np.full(n, t, np.uint64))
        act = (wp >> np.uint32(30)).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        o = vec_step(tab, pos, ln, ret, act, wt, limit=500, start_pos=0, autoreset=True)
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
        assert np.array_equal(obs.cpu().numpy(), o["obs"]) and np.array_equal(rew.cpu().numpy(), o["reward"])
    env.rollout(40)
    for t in range(60, 100):
        wt, wp = philox.step_words(seed, g, np.full(n, t, np.uint64))
        o = vec_step(tab, pos, ln, ret, (wp >> np.uint32(30)).astype(np.int32), wt, limit=500, start_pos=0, autoreset=True)
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
    assert np.array_equal(env.pos.cpu().numpy(), pos)


def test_reference_test_suite_on_vec_surface():
    """tests/envs/grid_world/test_grid_world.py re-expressed against VecGridWorld(num_envs=1)."""
    from rl_envs_forge_b200 import Action, VecGridWorld
    env = VecGridWorld(1, autoreset=False)
    assert (env.rows, env.cols, env.start_state) == (5, 5, (0, 0))                      # test_initialization
    obs, *_ = env.step(np.array([Action.DOWN]))
    assert tuple(obs[0]) == (1, 0)                                                        # test_actions
    env = VecGridWorld(1, autoreset=False)
    assert env.step(np.array([Action.RIGHT]))[1][0] == np.float32(-0.1)                   # test_rewards
    env.step(np.array([Action.DOWN]))
    obs, _ = env.reset()
    assert tuple(obs.cpu().numpy()[0]) == (0, 0)                                          # test_reset
    env.add_special_transition(from_state=(0, 0), action=Action.DOWN, to_state=(4, 4), reward=0.5)
    env.state = (0, 0)
    obs, rew, *_ = env.step(np.array([Action.DOWN]))
    assert tuple(obs[0]) == (4, 4) and rew[0] == 0.5                                      # test_special_transition
    env = VecGridWorld(1, autoreset=False, episode_length_limit=10)                       # test_episode_length_limit
    for i in range(9):
        assert not env.step(np.array([Action.RIGHT if i % 2 else Action.LEFT]))[3][0]
    _, _, term, trunc, _ = env.step(np.array([Action.LEFT]))
    assert trunc[0] and not term[0]
    _, _, _, trunc, _ = env.step(np.array([Action.LEFT]))
    assert trunc[0]                                                                       # truncation is sticky past the limit
    env = VecGridWorld(1, rows=5, cols=5, autoreset=False, strict=True)                   # test_reset_to_new_starting_state
    obs, _ = env.reset((5, 5))
    assert tuple(obs.cpu().numpy()[0]) == (5, 5)
    with pytest.raises(ValueError):                                                       # ...and stepping from there raises
        env.step(np.array([0]))
    with pytest.raises(ValueError):
        env.reset([1, 1])
    env = VecGridWorld(1, rows=4, cols=4, walls={(1, 1), (2, 2)}, terminal_states={(3, 3): 1.0})
    assert len(env.mdp) == (16 - 2) * 4                                                   # test_mdp_size
    env = VecGridWorld(1, rows=3, cols=3, walls={(1, 1)}, terminal_states={(2, 2): 1.0},
                       special_transitions={((0, 0), Action.RIGHT): ((0, 2), 0.0)})       # test_probability_matrix
    P = env.P
    assert P.shape == (9, 9, 4) and P[0, 2, Action.RIGHT] == 1.0 and P[4, 4, 0] == 1.0 and P[8, 8, 1] == 1.0
    env = VecGridWorld(1, autoreset=False, strict=True)
    with pytest.raises(ValueError):                                                       # Action(7) raises in the reference
        env.step(np.array([7]))
    # terminal self-loop with reward 0 and done=True when auto-reset is off (grid_world.py:263-265)
    env = VecGridWorld(1, autoreset=False)
    env.state = (3, 3)
    obs, rew, term, _, _ = env.step(np.array([Action.RIGHT]))
    assert tuple(obs[0]) == (3, 4) and rew[0] == 1.0 and term[0]
    obs, rew, term, _, _ = env.step(np.array([Action.UP]))
    assert tuple(obs[0]) == (3, 4) and rew[0] == 0.0 and term[0]


def test_autoreset_semantics_and_final_observation():
    import torch
    from rl_envs_forge_b200 import Action, VecGridWorld
    env = VecGridWorld(4, autoreset=True)
    env.state = np.array([[3, 3], [0, 0], [3, 3], [2, 2]])
    acts = torch.tensor([Action.RIGHT, Action.RIGHT, Action.UP, Action.LEFT], dtype=torch.int32, device="cuda")
    obs, rew, term, trunc, info = env.step(acts, return_final_obs=True)
    assert obs.cpu().tolist() == [[0, 0], [0, 1], [2, 3], [2, 1]]                 # env 0 finished -> start state
    assert info["final_observation"].cpu().tolist() == [[3, 4], [0, 1], [2, 3], [2, 1]]
    assert term.cpu().tolist() == [True, False, False, False]
    assert env.ep_len.cpu().tolist() == [0, 1, 1, 1] and env.ep_ret.cpu().tolist()[0] == 0.0
    st = env.episode_stats()
    assert st["episodes"] == 1 and st["length_sum"] == 1 and st["return_sum"] == 1.0
    obs, _ = env.reset(mask=torch.tensor([0, 1, 0, 0], dtype=torch.bool))
    assert env.ep_len.cpu().tolist() == [0, 0, 1, 1]


@pytest.mark.parametrize("transport", ["pipelined", "pipelined+h2d", "zerocopy", "staged"])
def test_host_buffer_path_equals_device_path(transport):
    """numpy / pinned actions in -> numpy out.  "zerocopy": the step kernel reads and writes pinned host memory itself;
    "staged": copy-engine H2D, kernel, D2H.  Both must equal the device-tensor path bit for bit."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(0.8))
    n = 10_001                                             # ragged tail included
    a = VecGridWorld(n, seed=4, **cfg)
    a.host_transport, a.host_stage_actions, a.host_chunks = transport.split("+")[0], transport.endswith("+h2d"), 3
    b = VecGridWorld(n, seed=4, **cfg)
    rng = np.random.default_rng(0)
    for i in range(20):
        act = rng.integers(0, 4, n).astype(np.int32)
        src = torch.from_numpy(act).pin_memory() if i % 2 else act   # caller-pinned buffers are used in place
        oa = a.step(src, return_final_obs=True)
        ob = b.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        assert isinstance(oa[0], np.ndarray) and oa[0].dtype == np.int32 and oa[2].dtype == np.bool_
        assert np.array_equal(oa[0], ob[0].cpu().numpy()) and np.array_equal(oa[1], ob[1].cpu().numpy())
        assert np.array_equal(oa[2], ob[2].cpu().numpy()) and np.array_equal(oa[3], ob[3].cpu().numpy())
        assert np.array_equal(oa[4]["final_observation"], ob[4]["final_observation"].cpu().numpy())
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)


# ---- stream-exact mode: rng="pcg64" -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["default", "slip_seed42", "config2_p1", "config2_p08", "p0", "random02", "random08"])
def test_pcg64_mode_reproduces_reference_trajectories_without_injection(name):
    """`VecGridWorld(1, seed=s, rng="pcg64")` fed the recorded actions reproduces the trajectory the UNMODIFIED reference
    produced with GridWorld(seed=s) and its own numpy generator: same states, rewards, flags -- no draws injected."""
    from rl_envs_forge_b200 import VecGridWorld
    g = load([f for f in FILES if f.endswith(f"gridworld_{name}.npz")][0])
    env = VecGridWorld(1, seed=int(g["seed"]), rng="pcg64", autoreset=False, strict=True, **_vec_kwargs(grid_kwargs(g)))
    env.reset()
    for i in range(min(600, len(g["action"]))):
        assert tuple(env.state.cpu().numpy()[0]) == tuple(g["pre"][i]), i
        if g["raised"][i]:
            with pytest.raises(ValueError):
                env.step(np.array([g["action"][i]]))
            continue
        obs, rew, term, trunc, _ = env.step(np.array([g["action"][i]]))
        assert tuple(obs[0]) == tuple(g["post"][i]) and rew[0] == np.float32(g["reward"][i]), i
        assert bool(term[0]) == bool(g["done"][i]) and bool(trunc[0]) == bool(g["trunc"][i])
        if g["done"][i] or g["trunc"][i]:
            env.reset()


def test_pcg64_mode_batch_matches_numpy_generators():
    """N envs with seeds s, s+1, ...: each follows the port driven by ITS numpy Generator(PCG64(SeedSequence(seed)))."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, steps, base = 512, 120, 1000
    cfg = _config2(0.8)
    env = VecGridWorld(n, seed=base, rng="pcg64", **_vec_kwargs(cfg))
    ports = [GridWorldPort(**cfg) for _ in range(n)]
    gens = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(base + i))) for i in range(n)]
    arng = np.random.default_rng(5)
    for p in ports:
        p.reset()
    for t in range(steps):
        act = arng.integers(0, 4, n).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        obs, rew, term, trunc = obs.cpu().numpy(), rew.cpu().numpy(), term.cpu().numpy(), trunc.cpu().numpy()
        for i, p in enumerate(ports):
            outs = p.mdp.get((p.state, int(act[i])), [])
            try:
                idx = gens[i].choice(len(outs), p=[o[3] for o in outs])
            except ValueError:
                assert tuple(obs[i]) == p.state and not term[i] and not trunc[i]
                continue
            nxt, r, d, _ = outs[idx]
            p.state = nxt
            p.episode_length_counter += 1
            tr = p.episode_length_counter >= p.episode_length_limit
            assert rew[i] == np.float32(r) and bool(term[i]) == d and bool(trunc[i]) == tr, (t, i)
            if d or tr:
                p.reset()                                  # device auto-reset: same-step reset observation
            assert tuple(obs[i]) == p.state, (t, i)
    assert int(env.err.cpu().numpy()[0]) > 0               # the ((3,3), UP) cell was hit and rejected without a draw


def test_cuda_graph_replay_with_device_clock_equals_eager_steps():
    """enable_device_clock(): a captured CUDA graph of step() launches can be replayed and keeps advancing the Philox time
    coordinate on the device -- state, outputs and statistics equal the eager run with host-side step counting."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, seed, per_graph, replays = 70_001, 31, 5, 6
    cfg = _vec_kwargs(_config2(0.8))
    eager = VecGridWorld(n, seed=seed, **cfg)
    graphed = VecGridWorld(n, seed=seed, **cfg)
    acts = [torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda") for _ in range(per_graph)]
    graphed.step(acts[0])                       # warm-up launch outside capture (module load), mirrored on the eager env
    eager.step(acts[0])
    graphed.enable_device_clock()
    stream = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for a in acts:
                graphed.step(a)
    torch.cuda.synchronize()
    # capture does not execute: bring the eager env to the same point (nothing yet), then replay
    for _ in range(replays):
        g.replay()
        for a in acts:
            out_e = eager.step(a)
    torch.cuda.synchronize()
    assert graphed.step_count == 1 + per_graph * replays == eager.step_count
    assert torch.equal(graphed.pos, eager.pos) and torch.equal(graphed.ep_len, eager.ep_len)
    assert torch.equal(graphed.ep_ret, eager.ep_ret)
    assert torch.equal(graphed._obs, out_e[0]) and torch.equal(graphed._reward, out_e[1])
    sg, se = graphed.episode_stats(), eager.episode_stats()
    assert sg["episodes"] == se["episodes"] > 0 and sg["length_sum"] == se["length_sum"] and sg["env_steps"] == se["env_steps"]
    graphed.disable_device_clock()
    graphed.step(acts[0])
    eager.step(acts[0])
    assert torch.equal(graphed.pos, eager.pos)


@pytest.mark.parametrize("cold", [False, True])
def test_work_ahead_of_the_wait_changes_no_result_in_any_launch_order(cold):
    """The step kernel prefetches its inputs and warms the device clock word BEFORE griddepcontrol.wait (a load, or under the
    EF_STEP_COLD_INPUTS hint a line prefetch) and reads the clock through one thread per CTA after it.  None of that may
    change a bit: a graph that chains steps of ONE env set (the clock word is under atomic fire while the next launch
    looks at it) and a graph that alternates TWO env sets, with and without the hint, against eager steps with host-side
    step counting."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, per_graph, replays = 1 << 18, 6, 5
    cfg = _vec_kwargs(_config2(0.8))
    acts = [torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda") for _ in range(per_graph)]
    eager = [VecGridWorld(n, seed=41 + i, env_offset=i * n, **cfg) for i in range(2)]
    graphed = [VecGridWorld(n, seed=41 + i, env_offset=i * n, cold_inputs=cold, **cfg) for i in range(2)]
    for e in eager + graphed:
        e.step(acts[0])
    for e in graphed:
        e.enable_device_clock()
    stream = torch.cuda.Stream()
    chain, alternate = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(chain, stream=stream):
            for a in acts:
                graphed[0].step(a)
        with torch.cuda.graph(alternate, stream=stream):
            for j, a in enumerate(acts):
                graphed[j % 2].step(a)
    torch.cuda.synchronize()
    for _ in range(replays):
        chain.replay()
        for a in acts:
            eager[0].step(a)
        alternate.replay()
        for j, a in enumerate(acts):
            eager[j % 2].step(a)
    torch.cuda.synchronize()
    for gset, eset in zip(graphed, eager):
        assert gset.step_count == eset.step_count
        assert torch.equal(gset.pos, eset.pos) and torch.equal(gset.ep_len, eset.ep_len) and torch.equal(gset.ep_ret, eset.ep_ret)
        assert torch.equal(gset._reward, eset._reward) and torch.equal(gset._obs, eset._obs)
        sg, se = gset.episode_stats(), eset.episode_stats()
        assert sg["episodes"] == se["episodes"] > 0 and sg["length_sum"] == se["length_sum"] and sg["env_steps"] == se["env_steps"]


def test_cold_inputs_hint_follows_the_live_env_sets():
    """cold_inputs=None: the hint is on once the env sets alive on the device exceed its L2 together (vec_env._Residency),
    and off again when they are gone; an explicit True / False wins; a bad value is refused."""
    import gc
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    from rl_envs_forge_b200.vec_env import RESIDENCY
    gc.collect()
    cfg = _vec_kwargs(_config2(0.8))
    small = VecGridWorld(1024, **cfg)
    l2 = RESIDENCY.l2_bytes(torch, small.device)
    before = small._cold_inputs()
    big = [VecGridWorld(l2 // 42 // 2 + 1024, env_offset=(i + 1) << 24, **cfg) for i in range(2)]   # 2 x (L2 / 2 + a bit)
    assert small._cold_inputs() and big[0]._cold_inputs()
    act = torch.zeros(1024, dtype=torch.int32, device="cuda")
    small.step(act)
    assert small._fast_args[0].wire & 0x2
    del big
    gc.collect()
    assert small._cold_inputs() == before
    if not before:
        small.step(act)
        assert not (small._fast_args[0].wire & 0x2)
    small.cold_inputs = True
    small.step(act)
    assert small._fast_args[0].wire & 0x2
    with pytest.raises(ValueError):
        VecGridWorld(8, cold_inputs="yes", **cfg)


def test_empty_batch_is_a_noop_through_the_c_abi():
    """n = 0 (empty input): every entry point returns EF_OK without touching memory."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld, _lib
    env = VecGridWorld(8, **_vec_kwargs(_config2(0.8)))
    lib = _lib.load()
    before = env.pos.clone()
    st = torch.cuda.current_stream().cuda_stream
    assert lib.ef_gridworld_step(env._desc, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None, None,
                                 1, 0, 0, None, None, None, None, None, None, None, None, 0, 1, st) == 0
    assert lib.ef_gridworld_rollout(env._desc, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None,
                                    1, 0, 0, 16, None, None, None, None, None, 0, st) == 0
    assert lib.ef_gridworld_rollout(env._desc, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None,
                                    1, 0, 0, 0, None, None, None, None, None, 8, st) == 0     # K = 0
    assert lib.ef_gridworld_reset(0, 8, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None, None, 0, st) == 0
    torch.cuda.synchronize()
    assert torch.equal(env.pos, before)
    # negative sizes / out-of-range global index are rejected, not launched
    assert lib.ef_gridworld_step(env._desc, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None, None,
                                 1, 0, 0, None, None, None, None, None, None, None, None, -1, 1, st) == 1
    assert lib.ef_gridworld_step(env._desc, env.pos.data_ptr(), env.ep_len.data_ptr(), env.ep_ret.data_ptr(), None, None,
                                 1, 0, 2 ** 32 - 4, None, None, None, None, None, None, None, None, 8, 1, st) == 1


def test_maximum_grid_size_65535_cells():
    """Largest supported grid (255 x 257 = 65535 cells, 16-bit cell index; 8 MB outcome table read through L1/L2)."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    rows, cols = 255, 257
    rng = np.random.default_rng(3)
    walls = {(int(r), int(c)) for r, c in zip(rng.integers(1, rows - 1, 4000), rng.integers(1, cols - 1, 4000))}
    cfg = dict(rows=rows, cols=cols, start_state=(0, 0), walls=walls, terminal_states={(rows - 1, cols - 1): 3.0, (0, cols - 1): -2.0},
               p_success=0.6, episode_length_limit=50)
    port = GridWorldPort(**cfg)
    tab = FlatTables(port)
    n, seed = 30_000, 12
    env = VecGridWorld(n, seed=seed, **cfg)
    # scatter the agents over the whole grid (incl. the last row / column) so large cell indices are exercised
    free = np.array([r * cols + c for r in range(rows) for c in range(cols) if (r, c) not in walls and (r, c) not in cfg["terminal_states"]])
    pos = rng.choice(free, n)
    env.state = np.stack([pos // cols, pos % cols], axis=1)
    pos = pos.astype(np.int64)
    ln, ret = np.zeros(n, np.int32), np.zeros(n, np.float32)
    g = np.arange(n)
    for t in range(25):
        wt, wp = philox.step_words(seed, g, np.full(n, t, np.uint64))
        act = (wp >> np.uint32(30)).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        o = vec_step(tab, pos, ln, ret, act, wt, limit=50, start_pos=0, autoreset=True)
        pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
        assert np.array_equal(obs.cpu().numpy(), o["obs"]) and np.array_equal(rew.cpu().numpy(), o["reward"])
        assert np.array_equal(term.cpu().numpy(), o["terminated"])
    with pytest.raises(ValueError):
        VecGridWorld(4, rows=256, cols=256)        # 65536 cells: one more than the 16-bit table index allows


# ---- packed state layout ----------------------------------------------------------------------------------------------
def test_packed_state_layout_equals_split_layout():
    """cell | episode_length_counter << 16 in one word per env (8 fewer bytes per env-step) must not change a bit:
    step (vector + scalar + pcg64), fused rollout, masked reset, state accessors, and leaving the layout."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(0.8))
    n = 20_003
    for rng_mode in ("philox", "pcg64"):
        a = VecGridWorld(n, seed=6, rng=rng_mode, **cfg)
        b = VecGridWorld(n, seed=6, rng=rng_mode, state_layout="split", **cfg)
        assert a._packed and not b._packed
        gen = np.random.default_rng(3)
        for i in range(260):                                   # past the 200-step limit: truncations + auto-resets
            act = torch.from_numpy(gen.integers(0, 4, n).astype(np.int32)).cuda()
            oa, ob = a.step(act, return_final_obs=True), b.step(act, return_final_obs=True)
            for x, y in zip(oa[:4], ob[:4]):
                assert torch.equal(x, y), i
            assert torch.equal(oa[4]["final_observation"], ob[4]["final_observation"])
            if i == 100:
                m = torch.from_numpy(gen.integers(0, 2, n).astype(np.uint8)).cuda()
                a.reset(mask=m)
                b.reset(mask=m)
        assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
        assert int(a.ep_len.max()) > 100
        if rng_mode == "philox":
            draws = torch.from_numpy(gen.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.int32)).cuda()
            act = torch.from_numpy(gen.integers(0, 4, n).astype(np.int32)).cuda()
            for x, y in zip(a.step(act, draws=draws)[:4], b.step(act, draws=draws)[:4]):   # scalar kernel
                assert torch.equal(x, y)
            a.rollout(37)
            b.rollout(37)
            ra, rb = a.rollout(5, return_rewards=True), b.rollout(5, return_rewards=True)
            assert torch.equal(ra[0], rb[0]) and torch.equal(ra[1], rb[1])
            assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
        sa, sb = a.stats.cpu().numpy(), b.stats.cpu().numpy()
        assert np.array_equal(sa[[0, 1, 2, 4]], sb[[0, 1, 2, 4]]) and torch.equal(a.err[:1], b.err[:1])
        assert sa[3] == pytest.approx(sb[3], rel=1e-12)   # sum of squared returns: inexact in binary64, atomic order varies
        ra, rb = a.running_episode_stats(), b.running_episode_stats()
        assert all(ra[k] == rb[k] for k in ("episodes", "length_sum", "return_sum"))
        # accessors: get/set state round trip, state setter, leaving the packed layout when auto-reset goes off
        st = b.get_state()
        a.set_state(st)
        assert torch.equal(a.pos, st["pos"]) and torch.equal(a.ep_len, st["ep_len"])
        a.state = (2, 3)
        assert int(a.pos.min()) == int(a.pos.max()) == 2 * 8 + 3 and torch.equal(a.ep_len, st["ep_len"])
        a.autoreset = False
        assert not a._packed and torch.equal(a.ep_len, st["ep_len"])
    with pytest.raises(ValueError):
        VecGridWorld(4, state_layout="packed")                 # no episode limit: the counter is unbounded


def test_lean_step_path_survives_reassigned_state_tensors():
    """The usual call shape caches its pointer arguments; replacing a state tensor by assignment must not leave a stale
    pointer behind."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(0.8))
    a, b = VecGridWorld(4096, seed=3, **cfg), VecGridWorld(4096, seed=3, **cfg)
    act = torch.randint(0, 4, (4096,), dtype=torch.int32, device="cuda")
    for _ in range(3):
        a.step(act)
        b.step(act)
    a.ep_ret = a.ep_ret.clone()                      # new tensor objects, same contents
    a.pos = a.pos.clone()
    for _ in range(5):
        for x, y in zip(a.step(act)[:4], b.step(act)[:4]):
            assert torch.equal(x, y)
    assert torch.equal(a.ep_ret, b.ep_ret) and torch.equal(a.pos, b.pos)


@pytest.mark.parametrize("external", [False, True])
def test_rollout_trajectory_of_observations_equals_step_sequence(external):
    """rollout(K, return_obs=True) materialises what K step() calls would have returned: obs [K,N,2], rewards, flags."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    n, K, seed = 30_001, 70, 12
    cfg = _vec_kwargs(_config2(0.8))
    cfg["episode_length_limit"] = 25
    a, b = VecGridWorld(n, seed=seed, **cfg), VecGridWorld(n, seed=seed, state_layout="split", **cfg)
    acts = torch.randint(0, 4, (K, n), dtype=torch.int32, device="cuda") if external else None
    obs, rew, flg = a.rollout(K, acts, return_obs=True)
    assert obs.shape == (K, n, 2) and obs.dtype == torch.int32
    for k in range(K):
        o = b.step(acts[k] if external else None_actions(b, k))
        assert torch.equal(obs[k], o[0]) and torch.equal(rew[k], o[1]), k
        assert torch.equal(flg[k], o[2].to(torch.uint8) | (o[3].to(torch.uint8) << 1)), k
    assert torch.equal(a.pos, b.pos) and torch.equal(a.ep_len, b.ep_len)


def test_wild_action_values_are_rejected_without_touching_the_table():
    """Actions far outside 0..3 (the reference raises ValueError in Action(action), grid_world.py:497) are counted and
    leave the env untouched; they must not be used as a table index on the way (vector and scalar kernels)."""
    import torch
    from rl_envs_forge_b200 import VecGridWorld
    cfg = _vec_kwargs(_config2(1.0))      # p_success = 1: no cell the reference raises on, every error is an action
    for n in (1 << 16, 4099, 3):
        env = VecGridWorld(n, seed=3, **cfg)
        twin = VecGridWorld(n, seed=3, **cfg)
        rng = np.random.default_rng(n)
        bad_total = 0
        for t in range(20):
            act = rng.integers(0, 4, n).astype(np.int32)
            ok = act.copy()
            bad = rng.random(n) < 0.01
            bad[0] = True
            act[bad] = rng.choice(np.array([4, -1, 1 << 20, -(1 << 31), (1 << 31) - 1, 1 << 16], np.int64), int(bad.sum())).astype(np.int32)
            bad_total += int(bad.sum())
            pre = env.pos.clone()
            obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
            o2 = twin.step(torch.from_numpy(ok).cuda())
            good = torch.from_numpy(~bad).cuda()
            assert torch.equal(obs[good], o2[0][good]) and torch.equal(rew[good], o2[1][good])
            assert torch.equal(env.pos[~good], pre[~good]) and float(rew[~good].abs().sum()) == 0.0
            twin.pos, twin.ep_len = env.pos.clone(), env.ep_len.clone()
            twin.ep_ret.copy_(env.ep_ret)
        assert int(env.err.cpu().numpy()[0]) == bad_total
