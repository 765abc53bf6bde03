"""GPU parity tests for the ACML envs: GamblersProblem and JacksCarRental through the C ABI vs
oracle/acml.py.  Integer state, rewards and flags: bit-exact (the Poisson multiplication method runs in FP64 on the device
over the same doubles the oracle builds from the same Philox words)."""
import json

import numpy as np
import pytest

from oracle import acml as OA
from oracle import philox
from tests._golden import golden_files, load

pytestmark = pytest.mark.gpu


# ---- GamblersProblem ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", golden_files("gambler"), ids=lambda p: p.split("gambler_")[-1][:-4])
def test_gambler_golden_steps_as_batch(path):
    """Every recorded step of the unmodified reference becomes one env: same capital, bet and injected draw."""
    import torch
    from rl_envs_forge_b200 import VecGamblersProblem
    g = load(path)
    kw = json.loads(str(g["config"]))
    n = len(g["action"])
    r32 = np.floor(g["u"] * 2.0 ** 32).astype(np.uint64).astype(np.uint32)
    # u -> r = floor(u * 2^32) keeps `u < p` unless u shares a 2^-32 bucket with p (probability ~2^-32 per step)
    assert np.array_equal(r32.astype(np.float64) * 2.0 ** -32 < kw["win_probability"], g["u"] < kw["win_probability"])
    env = VecGamblersProblem(n, autoreset=False, **kw)
    env.state = g["pre"].astype(np.int32)
    obs, rew, term, trunc, _ = env.step(torch.from_numpy(g["action"].astype(np.int32)).cuda(),
                                        draws=torch.from_numpy(r32.view(np.int32)).cuda())
    assert np.array_equal(obs.cpu().numpy(), g["post"])
    assert np.array_equal(rew.cpu().numpy(), g["reward"].astype(np.float32))
    assert np.array_equal(term.cpu().numpy(), g["done"]) and not bool(trunc.any())
    assert int(env.err[0]) == 0


@pytest.mark.parametrize("goal,p,n", [(100, 0.4, 65536), (16, 0.5, 4099), (30, 1.0, 1023), (30, 0.0, 7)])
def test_gambler_philox_steps_match_oracle(goal, p, n):
    import torch
    from rl_envs_forge_b200 import VecGamblersProblem
    seed, off, limit = 17, 1000, 9
    env = VecGamblersProblem(n, goal_amount=goal, win_probability=p, start_capital=3, seed=seed, env_offset=off,
                             episode_length_limit=limit)
    gidx = np.arange(off, off + n)
    state = np.full(n, 3, np.int64)
    ln = np.zeros(n, np.int64)
    ret = np.zeros(n, np.float32)
    rng = np.random.default_rng(goal)
    errors = episodes = 0
    for t in range(40):
        act = rng.integers(0, goal + 1, n).astype(np.int32)
        act[rng.random(n) < 0.01] = goal + 5                               # the reference asserts on these
        act[rng.random(n) < 0.01] = -1
        obs, rew, term, trunc, info = env.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        r32 = philox.group_word(seed, gidx, t, philox.STREAM_TRANSITION)
        s2, rw, done, invalid = OA.gambler_step_vec(state, act, r32, goal, p)
        ln = ln + ~invalid
        tr = (ln >= limit) & ~invalid
        ret = (ret + rw).astype(np.float32)
        assert np.array_equal(info["final_observation"].cpu().numpy(), s2), t
        fin = done | tr
        episodes += int(fin.sum())
        errors += int(invalid.sum())
        state = np.where(fin, 3, s2)
        ln = np.where(fin, 0, ln)
        ret = np.where(fin, np.float32(0), ret)
        assert np.array_equal(obs.cpu().numpy(), state), t
        assert np.array_equal(rew.cpu().numpy(), rw), t
        assert np.array_equal(term.cpu().numpy(), done) and np.array_equal(trunc.cpu().numpy(), tr), t
    assert np.array_equal(env.ep_len.cpu().numpy(), ln) and np.array_equal(env.ep_ret.cpu().numpy(), ret)
    st = env.episode_stats()
    assert st["episodes"] == episodes and st["env_steps"] == 40 * n - errors and int(env.err[0]) == errors
    with pytest.raises(ValueError):
        env.check_errors()


def test_gambler_scalar_path_policy_and_sharding():
    """Unaligned views take the scalar kernel; the in-kernel random policy bets legally; shards reproduce the whole."""
    import torch
    from rl_envs_forge_b200 import VecGamblersProblem
    n, seed = 10_007, 5
    whole = VecGamblersProblem(n, goal_amount=50, win_probability=0.45, start_capital=7, seed=seed)
    parts = [VecGamblersProblem(hi - lo, goal_amount=50, win_probability=0.45, start_capital=7, seed=seed, env_offset=lo)
             for lo, hi in ((0, 1001), (1001, 6002), (6002, n))]          # unaligned offsets: Philox groups straddle shards
    act = torch.randint(0, 20, (n,), dtype=torch.int32, device="cuda")
    for t in range(12):
        ow = whole.step(act)
        outs = [p.step(act[p.env_offset:p.env_offset + p.num_envs]) for p in parts]
        cat = [torch.cat([o[k] for o in outs]) for k in range(4)]
        for x, y in zip(ow[:4], cat):
            assert torch.equal(x, y), t
    assert torch.equal(whole.state, torch.cat([p.state for p in parts]))
    # random policy through the C ABI directly (action == NULL): bets stay within 0..min(s, goal - s)
    from rl_envs_forge_b200 import _lib
    before = whole.state.clone()
    whole._call(whole._lib.ef_gambler_step, whole._params, _lib.ptr(whole._state), _lib.ptr(whole.ep_len),
                _lib.ptr(whole.ep_ret), None, None, seed, 99, 0, _lib.ptr(whole._obs), _lib.ptr(whole._reward),
                _lib.ptr(whole._term), _lib.ptr(whole._trunc), None, None, _lib.ptr(whole.err), None, n, 0)
    delta = (whole.state - before).abs()
    assert bool((delta <= torch.minimum(before, 50 - before)).all()) and int(whole.err[0]) == 0


def test_gambler_reference_surface():
    """tests/envs/acml/gambler_problem/test_gambler_problem.py on the vector surface."""
    from rl_envs_forge_b200 import VecGamblersProblem
    env = VecGamblersProblem(1, goal_amount=100, win_probability=0.40, start_capital=10, autoreset=False, strict=True)
    assert int(env.state[0]) == 10                                         # test_initialization
    env.step(np.array([5]))
    obs, info = env.reset()                                                # test_reset
    assert int(obs[0]) == 10 and info == {}
    mdp = env.build_mdp()                                                  # test_mdp_build
    assert mdp == OA.gambler_build_mdp(100, 0.40) and len(mdp) > 0
    with pytest.raises(AssertionError):
        env.step(np.array([101]))                                          # assert action_space.contains(action)
    assert env.single_action_space.n == 101 and env.single_observation_space.n == 101


# ---- JacksCarRental ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(max_cars=4, max_move_cars=2, request_lambda=[2, 2], return_lambda=[2, 2]),
                                dict(max_cars=10, max_move_cars=3, request_lambda=[0, 9.5], return_lambda=[1, 0],
                                     rental_credit=7, move_car_cost=3),
                                dict(max_cars=40, max_move_cars=5, request_lambda=[12, 4], return_lambda=[15.5, 0]),
                                dict(max_cars=200, max_move_cars=20, request_lambda=[60, 10], return_lambda=[25, 133.3])],
                         ids=["default", "small", "zero_lambda", "ptrs_mixed", "ptrs_all"])
def test_car_rental_steps_match_oracle(kw):
    """Device step == reference step fed by the same Philox doubles (oracle pinned against the reference's own golden
    trajectories in tests/test_oracle_acml.py).  Exact: cars, rewards, counters."""
    import torch
    from rl_envs_forge_b200 import VecJacksCarRental
    n, seed, off = 30_001, 23, 4096
    p = OA.CarRentalParams(**kw)
    env = VecJacksCarRental(n, seed=seed, env_offset=off, **kw)
    gidx = np.arange(off, off + n)
    state = env.state.cpu().numpy().astype(np.int64)
    assert state.min() >= 0 and state.max() <= p.max_cars and len(np.unique(state)) > 1      # reset("random")
    rng = np.random.default_rng(1)
    total = np.zeros(n, np.float64)
    for t in range(6):
        act = rng.integers(0, 2 * p.max_move_cars + 1, n).astype(np.int32)
        obs, rew, term, trunc, _ = env.step(torch.from_numpy(act).cuda())
        state, want = OA.car_rental_engine_step_vec(p, seed, gidx, t, state, act)
        assert np.array_equal(obs.cpu().numpy(), state), t
        assert np.array_equal(rew.cpu().numpy(), want.astype(np.float32)), t
        assert not bool(term.any()) and not bool(trunc.any())
        total += want
    assert np.array_equal(env.ep_len.cpu().numpy(), np.full(n, 6))
    assert np.array_equal(env.ep_ret.cpu().numpy(), total.astype(np.float32))
    assert env.episode_stats()["env_steps"] == 6 * n


def test_car_rental_limit_reset_modes_and_host_path():
    import torch
    from rl_envs_forge_b200 import VecJacksCarRental
    n = 5003
    a = VecJacksCarRental(n, seed=3, episode_length_limit=4, init_state_option="equal")
    b = VecJacksCarRental(n, seed=3, episode_length_limit=4, init_state_option="equal")
    assert bool((a.state == 10).all())                                                      # max_cars // 2
    rng = np.random.default_rng(2)
    for t in range(9):
        act = rng.integers(0, 11, n).astype(np.int32)
        a.host_transport = "zerocopy" if t % 2 else "staged"
        oa = a.step(act, return_final_obs=True)                                             # numpy in -> numpy out
        ob = b.step(torch.from_numpy(act).cuda(), return_final_obs=True)
        assert isinstance(oa[0], np.ndarray) and oa[0].shape == (n, 2)
        for x, y in zip(oa[:4], ob[:4]):
            assert np.array_equal(x, y.cpu().numpy()), t
        assert np.array_equal(oa[4]["final_observation"], ob[4]["final_observation"].cpu().numpy())
        assert bool(ob[3].all()) == (t % 4 == 3)                                            # truncation every 4th step
        if t % 4 == 3:
            assert bool((ob[0] == 10).all())                                                # auto-reset to "equal"
    assert b.episode_stats()["episodes"] == 2 * n
    st, info = b.reset("random", mask=torch.arange(n, device="cuda") % 2 == 0)
    assert info == {"Reset Option": "random"} and int(b.ep_len[0]) == 0 and int(b.ep_len[1]) == 1
    before = b.state.clone()
    assert torch.equal(b.reset("bogus")[0], before)                                         # unknown option: untouched
    with pytest.raises(NotImplementedError):
        VecJacksCarRental(4, request_lambda=[2e6, 3])


def test_car_rental_reference_surface_and_mdp():
    """tests/envs/acml/car_rental/test_car_rental.py on the vector surface + build_mdp against the reference's dump."""
    from rl_envs_forge_b200 import VecJacksCarRental
    env = VecJacksCarRental(1, max_cars=4, max_move_cars=2, request_lambda=[2, 2], return_lambda=[2, 2])
    assert env.max_cars == 4 and env.max_move_cars == 2
    s = env.state.cpu().numpy()[0]
    assert 0 <= s[0] <= 4 and 0 <= s[1] <= 4
    env.step(np.array([0]))
    state, info = env.reset()
    assert tuple(state.shape) == (1, 2) and info == {"Reset Option": "random"}
    assert 0 <= env._get_probability_of_rental_requests(0, 0) <= 1
    assert 0 <= env._get_probability_of_car_returns(0, 0) <= 1
    g = load(golden_files("carrental")[0])
    mdp = env.build_mdp()
    assert len(mdp) == len(g["mdp_keys"])
    for key, nxt, rew in zip(g["mdp_keys"], g["mdp_next"], g["mdp_reward"]):
        got = mdp[((int(key[0]), int(key[1])), int(key[2]))]
        assert got[0] == tuple(int(x) for x in nxt) and got[2] is False
        assert abs(got[1] - rew) <= 1e-12 * max(1.0, abs(rew))


# ---- fused rollouts ----------------------------------------------------------------------------------------------------
def test_gambler_rollout_equals_step_sequence():
    """K fused steps == K step() calls: with external actions (rewards / flags emitted) and with the in-kernel policy
    (fed to step() through its host twin: bet = mulhi(policy word, min(s, goal - s) + 1))."""
    import torch
    from rl_envs_forge_b200 import VecGamblersProblem
    n, K, seed = 100_003, 40, 8
    kw = dict(goal_amount=60, win_probability=0.45, start_capital=9, episode_length_limit=25, seed=seed, env_offset=500)
    a, b = VecGamblersProblem(n, **kw), VecGamblersProblem(n, **kw)
    acts = torch.randint(0, 12, (K, n), dtype=torch.int32, device="cuda")
    acts[3, ::101] = 77                                              # rejected bets
    rew, flg = a.rollout(K, acts, return_rewards=True)
    for k in range(K):
        o = b.step(acts[k])
        assert torch.equal(rew[k], o[1]) and torch.equal(flg[k], o[2].to(torch.uint8) | (o[3].to(torch.uint8) << 1)), k
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    assert int(a.err[0]) == int(b.err[0]) == (n + 100) // 101
    g = np.arange(n) + 500
    a.rollout(K)                                                     # random policy, nothing emitted
    for k in range(K):
        wp = philox.group_word(seed, g, K + k, philox.STREAM_POLICY).astype(np.uint64)
        s = b.state.cpu().numpy().astype(np.int64)
        hi = np.maximum(np.minimum(s, 60 - s), 0).astype(np.uint64)
        b.step(torch.from_numpy(((wp * (hi + 1)) >> np.uint64(32)).astype(np.int32)).cuda())
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] and sa["length_sum"] == sb["length_sum"] and sa["env_steps"] == sb["env_steps"]
    assert sa["return_sum"] == sb["return_sum"]


@pytest.mark.parametrize("rates", [dict(), dict(request_lambda=[12, 4], return_lambda=[15.5, 0], max_cars=40)],
                         ids=["default", "ptrs"])
def test_car_rental_rollout_equals_step_sequence(rates):
    import torch
    from rl_envs_forge_b200 import VecJacksCarRental
    n, K, seed = 20_001, 12, 4
    kw = dict(seed=seed, env_offset=64, episode_length_limit=5, **rates)
    a, b = VecJacksCarRental(n, **kw), VecJacksCarRental(n, **kw)
    acts = torch.randint(0, 11, (K, n), dtype=torch.int32, device="cuda")
    rew, flg = a.rollout(K, acts, return_rewards=True)
    for k in range(K):
        o = b.step(acts[k])
        assert torch.equal(rew[k], o[1]) and torch.equal(flg[k], o[3].to(torch.uint8) << 1), k
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    g = np.arange(n) + 64
    a.rollout(K)
    for k in range(K):
        wp = philox.group_word(seed, g, K + k, philox.STREAM_POLICY).astype(np.uint64)
        b.step(torch.from_numpy(((wp * np.uint64(11)) >> np.uint64(32)).astype(np.int32)).cuda())
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    assert a.episode_stats()["episodes"] == b.episode_stats()["episodes"] == n * (2 * K // 5)


# ---- edge cases --------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 33, 1023, 1025])
def test_ragged_sizes(n):
    """Tiny and ragged batches (vector body + scalar tail of the gambler kernel, partial CTAs elsewhere) == the same envs
    cut out of a large batch (global env index keys every draw)."""
    import torch
    from rl_envs_forge_b200 import VecGamblersProblem, VecJacksCarRental
    off = 4096
    for cls, kw, hi in ((VecGamblersProblem, dict(goal_amount=40, win_probability=0.5, start_capital=5), 12),
                        (VecJacksCarRental, dict(init_state_option="equal"), 11)):
        big = cls(off + 2048, seed=9, **kw)
        small = cls(n, seed=9, env_offset=off, **kw)
        for t in range(6):
            act = torch.randint(0, hi, (off + 2048,), dtype=torch.int32, device="cuda")
            ob, os_ = big.step(act), small.step(act[off:off + n].clone())
            for x, y in zip(ob[:4], os_[:4]):
                assert torch.equal(x[off:off + n], y), (cls.__name__, n, t)
        big.rollout(5)
        small.rollout(5)
        assert torch.equal(big.state[off:off + n], small.state)


def test_empty_batches_are_noops_through_the_c_abi():
    import ctypes
    from rl_envs_forge_b200 import _lib
    lib = _lib.load()
    gp = _lib.GamblerParams()
    gp.goal_amount, gp.win_threshold = 10, 1 << 31
    fake = ctypes.c_void_p(0x1000)
    assert lib.ef_gambler_step(ctypes.byref(gp), fake, fake, fake, None, None, 0, 0, 0, *([None] * 8), 0, 1, None) == 0
    assert lib.ef_gambler_rollout(ctypes.byref(gp), fake, fake, fake, None, 0, 0, 0, 4, None, None, None, None, None, 0, None) == 0
    cp = _lib.CarRentalParams()
    cp.max_cars, cp.max_move_cars = 20, 5
    assert lib.ef_car_rental_step(ctypes.byref(cp), fake, fake, fake, None, 0, 0, 0, *([None] * 7), 0, 1, None) == 0
    cp.lam[1] = 12.0          # a rate in numpy's PTRS range needs its PTRS constants in the descriptor
    assert lib.ef_car_rental_step(ctypes.byref(cp), fake, fake, fake, None, 0, 0, 0, *([None] * 7), 8, 1, None) == 1
    cp.lam[1] = 2.0e6
    assert lib.ef_car_rental_step(ctypes.byref(cp), fake, fake, fake, None, 0, 0, 0, *([None] * 7), 8, 1, None) == 3   # unsupported
