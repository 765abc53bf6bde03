"""GPU parity tests for CartPole / PendulumDisk: fp32 device kernels (through the C ABI) vs the float64 reference.

Tolerance (BASELINE.json north_star): per step |device - ref| <= 1e-5 * |ref| + 1e-6 with the reference
re-seeded to the device's (fp32) state each step.  Flags and the discrete reward are threshold comparisons; they may only
differ where the compared quantity lies within the fp32 error band of the threshold -- those cases are counted."""
import json
import os

import numpy as np
import pytest

from oracle import pendulums as OP
from oracle import philox
from tests._golden import golden_files, load

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6
# accelerations are differences of terms ~100x larger than the result (divided by J = 1.91e-4 for the disk, |acc| up to
# 220): absolute floor on top of the 1e-5 relative band.  Observed on B200: CartPole never leaves the relative band,
# PendulumDisk exceeds it by at most 1.33e-5 (profiles/r02m_flips.log).
ACC_ATOL = 2e-5


def _within(dev, ref, angle_col=None):
    """|dev - ref| <= rtol*|ref| + atol; the angle component is compared on the circle (an fp32 angle one ulp below
    +pi and a float64 angle just above -pi are the same physical angle: the wrap decision sits inside the fp32 band)."""
    d = np.abs(dev.astype(np.float64) - ref)
    if angle_col is not None:
        d[:, angle_col] = np.minimum(d[:, angle_col], np.abs(2 * np.pi - d[:, angle_col]))
    return d <= RTOL * np.abs(ref) + ATOL


def _cls(name):
    import rl_envs_forge_b200 as r
    return {"cartpole": (r.VecCartPole, OP.CartPoleParams, OP.cartpole_step, 4),
            "pendulumdisk": (r.VecPendulumDisk, OP.PendulumDiskParams, OP.pendulum_disk_step, 2)}[name]


ANGLE_COL = {"cartpole": 2, "pendulumdisk": 0}


def _flag_band(kind, prm, ref_state, band=2e-6):
    """True where some threshold comparison of the reference lies within `band` (relative) of its threshold."""
    if kind == "cartpole":
        th, x = ref_state[:, 2], ref_state[:, 0]
        near = (np.abs(np.abs(x) - prm.x_threshold) < band * 3) | (np.abs(np.abs(th) - prm.theta_threshold_radians) < band)
        ang = th
    else:
        th, w = ref_state[:, 0], ref_state[:, 1]
        near = np.abs(np.abs(w) - 15 * np.pi) < band * 50
        ang = th
    if prm.angle_termination is not None:
        near |= np.abs(np.abs(ang) - prm.angle_termination) < band * 3
    if not prm.continuous_reward:
        near |= (np.abs(ang - prm.reward_angle_range[0]) < band) | (np.abs(ang - prm.reward_angle_range[1]) < band)
    return near


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_golden_steps_as_batch(kind):
    """Every recorded step of the unmodified reference becomes one env: same pre-state (rounded to fp32), same action."""
    import torch
    Vec, Params, _, dim = _cls(kind)
    for path in golden_files(kind):
        g = load(path)
        kw = json.loads(str(g["config"]))
        n = len(g["action"])
        env = Vec(n, autoreset=False, **kw)
        env.state = g["pre"].astype(np.float32)
        env.ep_len.copy_(torch.from_numpy((g["step"] - 1).astype(np.int32)))
        obs, rew, term, trunc, info = env.step(torch.from_numpy(g["action"].astype(np.float32)).cuda(), return_info=True)
        obs, rew = obs.cpu().numpy(), rew.cpu().numpy()
        assert _within(obs, g["post"], ANGLE_COL[kind]).all(), path
        prm = Params(**kw)
        near = _flag_band(kind, prm, g["post"])
        bad = (term.cpu().numpy() != g["done"]) | (trunc.cpu().numpy() != g["trunc"])
        assert not (bad & ~near).any(), path
        if prm.continuous_reward:
            assert _within(rew, g["reward"]).all()
        else:
            assert not ((rew != g["reward"]) & ~near).any()
        acc = info["x_acc"].cpu().numpy() if kind == "cartpole" else info["alpha_dot_dot"].cpu().numpy()
        excess = np.abs(acc - g["acc"][:, 0]) - 1e-5 * np.abs(g["acc"][:, 0])
        print(f"{os.path.basename(path)}: acceleration error beyond 1e-5 relative: max {excess.max():.3e} (|acc| up to "
              f"{np.abs(g['acc'][:, 0]).max():.1f})")
        assert np.all(excess <= ACC_ATOL), path


@pytest.mark.parametrize("kind,n", [("cartpole", 1 << 22), ("pendulumdisk", 1 << 22)])
def test_per_step_parity_reseeded_4M_envs(kind, n):
    """Config 4 size (4M envs): several steps, the float64 reference restarted from the device's fp32 state each step."""
    import torch
    Vec, Params, ref_step, dim = _cls(kind)
    env = Vec(n, seed=11, autoreset=False)
    prm = Params()
    if kind == "pendulumdisk":  # spread the state over the whole swing range, not just +-0.01
        rng = np.random.default_rng(0)
        env.state = rng.uniform([-np.pi, -40.0], [np.pi, 40.0], (n, 2)).astype(np.float32)
    else:
        rng = np.random.default_rng(1)
        env.state = rng.uniform([-2.5, -2, -0.25, -2], [2.5, 2, 0.25, 2], (n, 4)).astype(np.float32)
    flips = 0
    for t in range(6):
        s0 = env.state.cpu().numpy().astype(np.float64)
        steps0 = env.ep_len.cpu().numpy()
        a = torch.empty(n, device="cuda").uniform_(-3, 3)
        obs, rew, term, trunc, _ = env.step(a)
        ref = ref_step(prm, s0, a.cpu().numpy(), steps0)
        ok = _within(obs.cpu().numpy(), ref["state"], ANGLE_COL[kind])
        assert ok.all(), (t, np.abs(obs.cpu().numpy() - ref["state"])[~ok][:5], ref["state"][~ok.all(axis=1)][:5])
        near = _flag_band(kind, prm, ref["state"]) | (np.abs(np.abs(ref["state"][:, ANGLE_COL[kind]]) - np.pi) < 1e-5)
        bad = (term.cpu().numpy() != ref["terminated"]) | (rew.cpu().numpy() != ref["reward"])
        assert not (bad & ~near).any()
        flips += int(bad.sum())
    print(f"{kind}: threshold-band flag/reward flips over 6 steps x {n} envs: {flips}")
    assert flips <= 8      # 2.5e7 checked env-steps; observed 1 (CartPole) / 0 (PendulumDisk), all inside the threshold band


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_free_running_drift_K64(kind):
    """Fused 64-step rollout with external actions vs the float64 reference integrating freely from the same start:
    states drift apart only at fp32 level.  Stated drift (relative, 1e-3 floor; the inverted pendulum amplifies any
    perturbation exponentially, so the tail is set by the dynamics): median <= 2e-6, max <= 1e-2."""
    import torch
    Vec, Params, ref_step, dim = _cls(kind)
    n, K = 1 << 18, 64
    env = Vec(n, seed=5, episode_length_limit=10_000)
    prm = Params(episode_length_limit=10_000)
    s = env.state.cpu().numpy().astype(np.float64)
    init_dev = s.copy()
    acts = torch.empty((K, n), device="cuda").uniform_(-3, 3)
    rew, flg = env.rollout(K, acts, return_rewards=True)
    steps = np.zeros(n, int)
    alive = np.ones(n, bool)   # envs that have not terminated in the reference (after a reset trajectories differ by design)
    a_np = acts.cpu().numpy()
    flg_np = flg.cpu().numpy()
    for k in range(K):
        out = ref_step(prm, s, a_np[k], steps)
        s, steps = out["state"], out["current_step"]
        alive &= ~(out["terminated"] | out["truncated"]) & (flg_np[k] == 0)
    dev = env.state.cpu().numpy().astype(np.float64)
    diff = np.abs(dev - s)
    c = ANGLE_COL[kind]
    diff[:, c] = np.minimum(diff[:, c], np.abs(2 * np.pi - diff[:, c]))
    rel = (diff / np.maximum(np.abs(s), 1e-3))[alive]
    assert alive.sum() > 2000   # random +-3 N forces topple most CartPoles within 64 steps; survivors suffice
    print(f"{kind}: K=64 drift over {alive.sum()} surviving envs: median {np.median(rel):.3e}  max {rel.max():.3e}")
    assert np.median(rel) <= 2e-6 and rel.max() <= 1e-2
    # initial states came from Philox STREAM_RESET_USER: bit-exact host twin
    w = philox.draw_block(5, np.arange(n), 0, philox.STREAM_RESET_USER)
    want = np.stack([philox.uniform_f32(w[i], -0.01, 0.01) for i in range(dim)], axis=1)
    assert np.array_equal(init_dev.astype(np.float32), want)


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_rollout_equals_step_sequence_bit_exact(kind):
    """Same kernels' arithmetic: fused rollout == K step() calls, bit for bit (state, counters, emitted rewards/flags,
    statistics), including auto-reset initial states (Philox STREAM_RESET_AUTO keyed by the step index)."""
    import torch
    Vec, Params, ref_step, dim = _cls(kind)
    n, K = 300_001, 48
    kw = dict(episode_length_limit=17, angle_termination=0.05 if kind == "cartpole" else 0.02)
    a, b = Vec(n, seed=9, **kw), Vec(n, seed=9, **kw)
    acts = torch.empty((K, n), device="cuda").uniform_(-3, 3)
    rew, flg = a.rollout(K, acts, return_rewards=True)
    for k in range(K):
        _, r, te, tr, _ = b.step(acts[k])
        assert torch.equal(rew[k], r) and torch.equal(flg[k], te.to(torch.uint8) | (tr.to(torch.uint8) << 1))
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len) and torch.equal(a.ep_ret, b.ep_ret)
    sa, sb = a.episode_stats(), b.episode_stats()
    assert sa["episodes"] == sb["episodes"] > 0 and sa["length_sum"] == sb["length_sum"]
    assert sa["return_sum"] == pytest.approx(sb["return_sum"], rel=1e-12)
    # in-kernel U(-3,3) policy == host twin of the policy word
    c, d = Vec(n, seed=9, **kw), Vec(n, seed=9, **kw)
    c.rollout(11)
    g = np.arange(n)
    for t in range(11):
        _, wp = philox.step_words(9, g, np.full(n, t, np.uint64))
        d.step(torch.from_numpy(philox.uniform_f32(wp, -3.0, 3.0)).cuda())
    assert torch.equal(c.state, d.state)


def test_auto_reset_draws_match_host_twin():
    import torch
    from rl_envs_forge_b200 import VecCartPole
    n = 4096
    env = VecCartPole(n, seed=21, episode_length_limit=3)
    for t in range(3):
        obs, rew, term, trunc, info = env.step(torch.zeros(n, device="cuda"), return_final_obs=True)
    assert bool(trunc.all()) and not bool(term.any())
    w = philox.draw_block(21, np.arange(n), 2, philox.STREAM_RESET_AUTO)   # reset at the end of step index 2
    want = np.stack([philox.uniform_f32(w[i], -0.01, 0.01) for i in range(4)], axis=1)
    assert np.array_equal(obs.cpu().numpy(), want)
    assert not np.array_equal(info["final_observation"].cpu().numpy(), want)
    assert env.ep_len.cpu().numpy().max() == 0
    assert env.episode_stats()["episodes"] == n


def test_angle_wrap_matches_reference():
    """normalize_angle values of the reference tests (test_cart_pole.py:151-156) through a device step with tau*omega = angle."""
    import torch
    from rl_envs_forge_b200 import VecPendulumDisk
    angles = np.array([0.0, np.pi, -np.pi, 2 * np.pi, -2 * np.pi, 3 * np.pi, -3 * np.pi, 3.2, -3.2, 6.0, -6.0, 100.0])
    env = VecPendulumDisk(len(angles), tau=1.0, autoreset=False, episode_length_limit=100)
    st = np.zeros((len(angles), 2), np.float32)
    st[:, 1] = angles.astype(np.float32)          # alpha' = normalize(0 + 1.0 * alpha_dot)
    env.state = st
    obs, *_ = env.step(torch.zeros(len(angles), device="cuda"))
    want = OP.normalize_angle(st[:, 1].astype(np.float64))
    got = obs.cpu().numpy()[:, 0].astype(np.float64)
    assert np.all(np.abs(got - want) <= 1e-6), (got, want)
    assert got[1] < -3.14 and got[2] > 3.14       # fp32(pi) lies above pi and wraps negative; fp32(-pi) wraps positive


def test_reference_test_suite_on_vec_surface():
    """tests/envs/inverted_pendulum/{cart_pole/test_cart_pole.py, pendulum_disk/test_pendulum_disk.py} at num_envs=1."""
    import torch
    from rl_envs_forge_b200 import VecCartPole, VecPendulumDisk
    env = VecCartPole(1, autoreset=False)
    assert (env.tau, env.continuous_reward, env.episode_length_limit) == (0.02, False, 1000)
    obs, _ = env.reset()
    assert obs.shape == (1, 4) and np.all(np.abs(obs.cpu().numpy()) <= 0.01)                 # test_state_initialization
    obs, rew, term, trunc, _ = env.step(np.array([[0.5]], np.float32))
    assert rew[0] == 1.0 and not term[0] and not trunc[0]                                    # test_step
    env = VecCartPole(1, continuous_reward=True, autoreset=False)                            # test_continuous_reward
    env.state = [0.0, 0.0, 0.1, 0.0]
    obs, rew, *_ = env.step(np.array([0.0], np.float32))
    assert rew[0] == pytest.approx(1 - abs(obs[0, 2]) / 0.2, rel=1e-5)
    env = VecCartPole(1, continuous_reward=True, nonlinear_reward=True, reward_decay_rate=30.0, autoreset=False)
    env.state = [0.0, 0.0, 0.1, 0.0]
    obs, rew, *_ = env.step(np.array([0.0], np.float32))                                     # test_nonlinear_reward
    assert rew[0] == pytest.approx(1 - (1 - np.exp(-30 * abs(obs[0, 2]))) / (1 - np.exp(-30 * 0.2)), rel=1e-5)
    env = VecCartPole(1, autoreset=False)                                                    # test_discrete_reward
    env.state = [0.0, 0.0, 0.0, 0.0]
    assert env.step(np.array([0.0], np.float32))[1][0] == 1.0
    env.state = [0.0, 0.0, 0.2, 0.0]
    assert env.step(np.array([0.0], np.float32))[1][0] == 0.0
    env = VecCartPole(1, episode_length_limit=10, autoreset=False)                           # test_episode_length_limit
    for i in range(10):
        _, _, term, trunc, _ = env.step(np.array([0.0], np.float32))
        assert bool(trunc[0]) == (i == 9) and not term[0]
    env = VecCartPole(1, angle_termination=0.1, autoreset=False)                             # test_angle_termination
    env.state = [0.0, 0.0, 0.11, 0.0]
    _, _, term, trunc, _ = env.step(np.array([0.0], np.float32))
    assert term[0] and not trunc[0]
    env = VecCartPole(1, strict=True, autoreset=False)                                       # test_step_with_invalid_action
    with pytest.raises(AssertionError):
        env.step(np.array([5.0], np.float32))
    with pytest.raises(AssertionError):
        env.step(np.array([1.0], np.float64))
    env = VecCartPole(1, autoreset=False)
    obs, _ = env.reset(initial_state=[0.1, 0.1, 0.1, 0.1])                                   # SURVEY 8(c) known answer
    obs, rew, *_ = env.step(np.array([1.0], np.float32))
    assert np.allclose(obs[0], [0.102, 0.11807538, 0.102, 0.1023734], rtol=1e-5, atol=1e-6) and rew[0] == 0.0
    env = VecPendulumDisk(1, autoreset=False)
    assert (env.tau, env.reward_angle_range) == (0.005, (-0.2, 0.2))
    obs, _ = env.reset(initial_state=[0.1, 0.1])
    obs, rew, *_ = env.step(np.array([1.0], np.float32), return_info=True)
    assert np.allclose(obs[0], [0.1005, 0.306123], rtol=1e-5, atol=1e-6) and rew[0] == 1.0
    env = VecPendulumDisk(1, continuous_reward=True, autoreset=False)
    env.state = [0.1, 0.0]
    obs, rew, *_ = env.step(np.array([0.0], np.float32))
    assert rew[0] == pytest.approx(1 - abs(obs[0, 0]) / np.pi, rel=1e-5)                     # test_pendulum_disk.py:78-86
    env = VecPendulumDisk(1, autoreset=False)
    env.state = [0.25, 0.0]
    assert env.step(np.array([0.0], np.float32))[1][0] == 0.0                                # outside +-0.2
    env = VecPendulumDisk(1, angle_termination=0.1, autoreset=False)
    env.state = [0.11, 0.0]
    _, _, term, trunc, _ = env.step(np.array([0.0], np.float32))
    assert term[0] and not trunc[0]
    assert torch.cuda.is_available()


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
@pytest.mark.parametrize("transport", ["pipelined", "pipelined+h2d", "zerocopy", "staged"])
def test_host_buffer_path_equals_device_path(kind, transport):
    """numpy actions in -> numpy out; the zero-copy transport (kernel reads/writes pinned host memory) and the staged
    one (copy engines) must both equal the device-tensor path bit for bit."""
    import torch
    Vec, _, _, dim = _cls(kind)
    n = 4099
    a, b = Vec(n, seed=9, episode_length_limit=7), Vec(n, seed=9, episode_length_limit=7)
    a.host_transport, a.host_stage_actions, a.host_chunks = transport.split("+")[0], transport.endswith("+h2d"), 3
    rng = np.random.default_rng(1)
    for i in range(12):
        act = rng.uniform(-3, 3, n).astype(np.float32)
        oa = a.step(act.reshape(n, 1) if i % 2 else act, return_final_obs=True, return_info=True)
        ob = b.step(torch.from_numpy(act).cuda(), return_final_obs=True, return_info=True)
        assert isinstance(oa[0], np.ndarray) and oa[0].shape == (n, dim) and oa[2].dtype == np.bool_
        for x, y in zip(oa[:4], ob[:4]):
            assert np.array_equal(x, y.cpu().numpy())
        assert np.array_equal(oa[4]["final_observation"], ob[4]["final_observation"].cpu().numpy())
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len)


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_aliased_observation_equals_separate_observation_buffer(kind):
    """alias_obs=True (default): the state tensor doubles as the observation (the reference returns np.array(self.state));
    alias_obs=False writes a separate buffer.  Same numbers either way, over auto-resets."""
    import torch
    Vec, _, _, dim = _cls(kind)
    n = 8195
    a, b = Vec(n, seed=2, episode_length_limit=9), Vec(n, seed=2, episode_length_limit=9, alias_obs=False)
    assert a._obs is a._state and b._obs is not b._state
    for t in range(25):
        act = torch.empty(n, device="cuda").uniform_(-3, 3)
        oa, ob = a.step(act), b.step(act)
        for x, y in zip(oa[:4], ob[:4]):
            assert torch.equal(x, y), t
        assert oa[0].data_ptr() == a.state.data_ptr()
    assert torch.equal(a.state, b.state) and a.episode_stats()["episodes"] == b.episode_stats()["episodes"] > 0


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_rollout_trajectory_of_observations_equals_step_sequence(kind):
    """rollout(K, return_obs=True): obs [K,N,DIM] / rewards / flags are bit for bit what K step() calls return."""
    import torch
    Vec, _, _, dim = _cls(kind)
    n, K = 20_003, 40
    a, b = Vec(n, seed=6, episode_length_limit=11), Vec(n, seed=6, episode_length_limit=11)
    acts = torch.empty((K, n), device="cuda").uniform_(-3, 3)
    obs, rew, flg = a.rollout(K, acts, return_obs=True)
    assert obs.shape == (K, n, dim)
    for k in range(K):
        o = b.step(acts[k])
        assert torch.equal(obs[k], o[0]) and torch.equal(rew[k], o[1]), k
        assert torch.equal(flg[k], o[2].to(torch.uint8) | (o[3].to(torch.uint8) << 1)), k
    assert torch.equal(a.state, b.state) and torch.equal(a.ep_len, b.ep_len)


@pytest.mark.parametrize("kind", ["cartpole", "pendulumdisk"])
def test_invalid_actions_are_flagged_and_leave_the_env_untouched(kind):
    """The reference asserts `action_space.contains(action)` (cart_pole.py:108, pendulum_disk.py:109): 5.0, NaN and
    -3.0001 are rejected.  On the device such env-steps are counted in `err` (EF_STEP_BAD_ACTION), the env keeps its state
    and counters, reward and flags are zero; check_errors() raises the reference's AssertionError with the offender.
    The other envs of the batch step exactly as they do without the bad actions."""
    import torch
    cls = _cls(kind)[0]
    n = 4096
    env, ref = cls(n, seed=3), cls(n, seed=3)
    a = torch.empty(n, device="cuda").uniform_(-3, 3)
    for _ in range(5):                       # move away from the initial state, identical in both
        env.step(a)
        ref.step(a)
    bad = {17: 5.0, 1000: float("nan"), 4095: -3.0001}
    a_bad = a.clone()
    for i, v in bad.items():
        a_bad[i] = v
    before = {k: v.clone() for k, v in env.get_state().items() if torch.is_tensor(v)}
    steps0 = float(env.stats[4].item())
    obs, rew, term, trunc, _ = env.step(a_bad)
    robs, rrew, rterm, rtrunc, _ = ref.step(a)
    idx = torch.tensor(sorted(bad), device="cuda")
    keep = torch.ones(n, dtype=torch.bool, device="cuda")
    keep[idx] = False
    assert torch.equal(obs[keep], robs[keep]) and torch.equal(rew[keep], rrew[keep])
    assert torch.equal(term[keep], rterm[keep]) and torch.equal(trunc[keep], rtrunc[keep])
    assert torch.equal(env.state[idx], before["state"][idx])                 # untouched
    assert torch.equal(env.ep_len[idx], before["ep_len"][idx]) and torch.equal(env.ep_ret[idx], before["ep_ret"][idx])
    assert torch.all(rew[idx] == 0) and not bool(term[idx].any()) and not bool(trunc[idx].any())
    assert float(env.stats[4].item()) - steps0 == n - len(bad)               # rejected env-steps are not counted
    assert int(env.err[0].item()) == len(bad)
    with pytest.raises(AssertionError, match="first offender env 4095"):     # largest global index wins the offender word
        env.check_errors()
    assert int(env.err[0].item()) == 0
    env.check_errors()                                                       # cleared
    # the fused rollout with external actions applies the same rule, step by step
    K = 4
    acts = torch.empty((K, n), device="cuda").uniform_(-3, 3)
    acts_bad = acts.clone()
    acts_bad[2, 7] = 3.5
    acts_bad[0, 7] = float("inf")
    s0 = float(env.stats[4].item())
    r_bad, f_bad = env.rollout(K, acts_bad, return_rewards=True)
    assert int(env.err[0].item()) == 2 and float(env.stats[4].item()) - s0 == K * n - 2
    assert float(r_bad[2, 7]) == 0.0 and int(f_bad[2, 7]) == 0 and float(r_bad[0, 7]) == 0.0
    with pytest.raises(AssertionError):
        env.check_errors()
    # the boundary itself is valid, like Box(-3, 3).contains
    env.step(torch.full((n,), 3.0, device="cuda"))
    env.step(torch.full((n,), -3.0, device="cuda"))
    env.check_errors()
