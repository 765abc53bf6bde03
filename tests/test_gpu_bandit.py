"""GPU parity tests for KArmedBandit: device fp32 samplers (through the C ABI) vs the float64 legacy-numpy algorithms
fed the same Philox primitives (oracle/bandit.py).  Actions / observations / timestep: exact.  Rewards: rtol 1e-5 +
atol 1e-6, except accept/reject flips of the rejection samplers, which are counted."""
import numpy as np
import pytest

from oracle import bandit as OB
from oracle import philox
from oracle.gen_golden import BANDIT_ARMS

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-6

CASES = [("normal", (5.0, 1.0)), ("uniform", (4.0, 6.0)), ("exponential", (0.2,)), ("gamma", (9.0, 0.55)),
         ("gamma", (0.4, 2.0)), ("gamma", (1.0, 2.0)), ("beta", (2.0, 5.0)), ("beta", (0.5, 0.7)),
         ("logistic", (5.0, 0.3)), ("weibull", (2.0,)), ("weibull", (0.7,)), ("lognormal", (1.6, 0.3))]
_NAMES = {"normal": ("mean", "std"), "uniform": ("low", "high"), "exponential": ("scale",), "gamma": ("shape", "scale"),
          "beta": ("a", "b"), "logistic": ("loc", "scale"), "weibull": ("a",), "lognormal": ("mean", "sigma")}


def _close(dev, ref):
    return np.abs(dev.astype(np.float64) - ref) <= RTOL * np.abs(ref) + ATOL


@pytest.mark.parametrize("fam,args", CASES)
def test_family_parity_vs_legacy_numpy_algorithms(fam, args):
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n, seed = 1 << 18, 77
    params = dict(distribution=fam, **dict(zip(_NAMES[fam], args)))
    env = VecKArmedBandit(n, k=1, arm_params={0: params}, seed=seed)
    flips = 0
    for t in range(3):
        obs, rew, done, trunc, _ = env.step(torch.zeros(n, dtype=torch.int32, device="cuda"))
        ref, _ = OB.engine_sample_vec(seed, np.arange(n), t, fam, args[0], args[1] if len(args) > 1 else 0.0)
        ok = _close(rew.cpu().numpy(), ref)
        flips += int((~ok).sum())
        assert bool(done.all()) and not bool(trunc.any()) and int(obs.abs().sum()) == 0
    print(f"{fam}{args}: {flips} accept/reject flips in {3 * n} samples")
    # Rejection samplers (gamma, beta) may in principle flip an accept/reject decision that falls inside the fp32 rounding
    # of its comparison; for these pinned seeds the count is 0 in all 12 cases (786432 samples each, measured on B200,
    # profiles/r02m_flips.log), so 0 is what is asserted -- a changed sampler has to re-earn it.
    assert flips == 0
    assert env.timestep == 3


def test_config3_mixed_arms_with_shift_16M():
    """BASELINE config 3: 10 arms (8 families + 2 default normals), absolute parameter shift on arm 0, 2^24 bandits."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n, seed, K = 1 << 24, 123, 3
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[0]["param_functions"] = [{"function": lambda t: 0.1 * t, "target_param": "mean"},
                                  {"function": lambda t: -0.01 * t, "target_param": "std"}]
    env = VecKArmedBandit(n, k=10, arm_params=arms, seed=seed)
    sub = np.arange(0, n, 64)                      # the float64 oracle checks a 262144-env subsample
    means = env.default_means[torch.from_numpy(sub).cuda()].cpu().numpy()
    flips = 0
    for t in range(K):
        act = torch.randint(0, 10, (n,), dtype=torch.int32, device="cuda")
        obs, rew, done, trunc, _ = env.step(act)
        assert torch.equal(obs, act) and env.timestep == t + 1
        a, r = act.cpu().numpy()[sub], rew.cpu().numpy()[sub]
        ref = np.zeros(len(sub))
        for arm in range(10):
            m = a == arm
            if arm >= 8:
                mu = OB.default_arm_mean(seed, sub[m], arm)
                assert np.all(np.abs(means[m, arm] - mu) <= 2e-6 * np.abs(mu) + 1e-6)
                z, _ = OB.engine_sample_vec(seed, sub[m], t, "normal", 0.0, 1.0)
                ref[m] = mu + z
                continue
            p = dict(BANDIT_ARMS[arm])
            fam = p.pop("distribution")
            vals = [float(p[k]) for k in _NAMES[fam]]
            if arm == 0:                           # reward of step t+1 uses init + f(t) (SURVEY 2.2 ordering)
                vals = [5 + 0.1 * t, 1 - 0.01 * t]
            ref[m], _ = OB.engine_sample_vec(seed, sub[m], t, fam, vals[0], vals[1] if len(vals) > 1 else 0.0)
        flips += int((~_close(r, ref)).sum())
    print(f"config3: {flips} flips in {K * len(sub)} checked samples")
    assert flips == 0      # observed on B200 for this seed (786432 checked samples); see the note in the family test
    assert env.arms[0].current_params["mean"] == pytest.approx(5 + 0.1 * K)
    st = env.episode_stats()
    assert st["episodes"] == K * n and st["length_sum"] == K * n


def test_reset_replays_stream_and_keeps_timestep():
    """k_armed_bandit.py:202-217: reset rebuilds arms and the generator (fixed seed => identical rewards), timestep survives."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n = 10_000
    env = VecKArmedBandit(n, k=10, seed=0)
    act = torch.randint(0, 10, (n,), dtype=torch.int32, device="cuda")
    r1 = [env.step(act)[1].clone() for _ in range(3)]
    m1 = env.default_means.clone()
    obs, info = env.reset()
    assert int(obs.abs().sum()) == 0 and env.timestep == 3
    r2 = [env.step(act)[1].clone() for _ in range(3)]
    assert all(torch.equal(x, y) for x, y in zip(r1, r2)) and torch.equal(m1, env.default_means)
    env.reset(seed=1)
    assert not torch.equal(env.step(act)[1], r1[0])


def test_rollout_equals_steps_and_random_policy():
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n, K = 100_003, 9
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[0]["param_functions"] = [{"function": lambda t: 0.1 * t, "target_param": "mean"}]
    a = VecKArmedBandit(n, k=10, arm_params=arms, seed=3)
    b = VecKArmedBandit(n, k=10, arm_params=arms, seed=3)
    acts = torch.randint(0, 10, (K, n), dtype=torch.int32, device="cuda")
    obs, rew = a.rollout(K, acts)
    for j in range(K):
        assert torch.equal(b.step(acts[j])[1], rew[j])
    assert a.timestep == b.timestep == K and a.arms[0].current_params == b.arms[0].current_params
    obs, rew = a.rollout(4)                        # in-kernel uniform random arm = mulhi(policy word, k)
    for j in range(4):
        _, wp = philox.step_words(3, np.arange(n), np.full(n, K + j, np.uint64))
        want = ((wp.astype(np.uint64) * 10) >> np.uint64(32)).astype(np.int32)
        assert np.array_equal(obs[j].cpu().numpy(), want)


def test_reference_test_suite_on_vec_surface():
    """tests/envs/k_armed_bandit/test_k_armed_bandit.py at num_envs=1."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    with pytest.raises(ValueError):                                                    # test_invalid_arm_index
        VecKArmedBandit(1, k=3, arm_params={3: dict(distribution="normal", mean=0, std=1)})
    with pytest.raises(ValueError):                                                    # test_not_implemented_arm
        VecKArmedBandit(1, k=1, arm_params={0: dict(distribution="cauchy")})
    env = VecKArmedBandit(1, k=1, seed=0, arm_params={0: dict(distribution="normal", mean=0, std=1, param_functions=[
        {"function": lambda t: t, "target_param": "mean"}])})
    for _ in range(10):                                                                # test_linear_param_shift
        obs, rew, done, trunc, info = env.step(np.array([0]))
        assert obs[0] == 0 and done[0] and not trunc[0] and info == {}
    assert env.arms[0].current_params["mean"] == 10 and env.state[0] == 10
    for fam, args in CASES:                                                            # test_arms_sampling
        e = VecKArmedBandit(1, k=1, seed=0, arm_params={0: dict(distribution=fam, **dict(zip(_NAMES[fam], args)))})
        rs = [float(e.step(np.array([0]))[1][0]) for _ in range(20)]
        assert all(np.isfinite(rs))
    env = VecKArmedBandit(1, k=10, seed=0, strict=True)
    with pytest.raises(AssertionError):
        env.step(np.array([10]))
    env = VecKArmedBandit(1, k=1, arm_params={0: dict(distribution="gamma", shape=2.0)})
    with pytest.raises(ValueError):                                                    # missing required parameter
        env.step(np.array([0]))
    env = VecKArmedBandit(1, k=1, arm_params={0: dict(distribution="normal", mean=0, std=1, param_functions=[
        {"function": lambda t: -2.0 * t, "target_param": "std"}])})
    env.step(np.array([0]))
    with pytest.raises(ValueError, match="scale < 0"):                                  # numpy's error for std < 0
        env.step(np.array([0]))
    assert torch.cuda.is_available()


@pytest.mark.parametrize("transport", ["pipelined", "pipelined+h2d", "zerocopy", "staged"])
def test_host_buffer_path_equals_device_path(transport):
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n = 5003
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    a, b = VecKArmedBandit(n, k=10, arm_params=arms, seed=5), VecKArmedBandit(n, k=10, arm_params=arms, seed=5)
    a.host_transport, a.host_stage_actions, a.host_chunks = transport.split("+")[0], transport.endswith("+h2d"), 3
    rng = np.random.default_rng(2)
    for _ in range(5):
        act = rng.integers(0, 10, n).astype(np.int32)
        oa, ob = a.step(act), b.step(torch.from_numpy(act).cuda())
        assert isinstance(oa[1], np.ndarray)
        assert np.array_equal(oa[0], ob[0].cpu().numpy()) and np.array_equal(oa[1], ob[1].cpu().numpy())
        assert oa[2].all() and not oa[3].any()


def test_family_sorted_kernel_is_bit_identical_to_plain_kernel(monkeypatch):
    """The default kernel evaluates a CTA's tile in sampler-sorted order; the plain one-env-per-thread kernel
    (ENVFORGE_BANDIT_UNSORTED=1) must give the same bits: step, K-step rollout, random policy, rejected actions, ragged n."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[8] = dict(distribution="gamma", shape=0.4, scale=2.0)
    arms[9] = dict(distribution="beta", a=0.5, b=0.7)
    arms[10] = dict(distribution="gamma", shape=1.0, scale=2.0)
    arms[11] = dict(distribution="beta", a=0.5, b=3.0)
    for n in (1, 31, 2047, 2049, 4095, 4097, 100_003):         # around the 4096-env tile of the sorted kernel
        outs = []
        for unsorted in (False, True):
            if unsorted:
                monkeypatch.setenv("ENVFORGE_BANDIT_UNSORTED", "1")
            else:
                monkeypatch.delenv("ENVFORGE_BANDIT_UNSORTED", raising=False)
            env = VecKArmedBandit(n, k=14, arm_params=arms, seed=11)
            act = torch.from_numpy(np.random.default_rng(n).integers(0, 14, n).astype(np.int32)).cuda()
            act[::97] = 14                                    # rejected: "Invalid action" in the reference
            o1, r1, *_ = env.step(act)
            o2, r2 = env.rollout(3)                           # in-kernel random policy
            a3 = torch.from_numpy(np.random.default_rng(n + 1).integers(0, 14, (2, n)).astype(np.int32)).cuda()
            o3, r3 = env.rollout(2, a3)
            outs.append([x.clone() for x in (o1, r1, o2, r2, o3, r3)] + [env.err[:1].clone(), env.stats[:2].clone(), env.stats[4].clone()])   # (which offender err[2:] names is kernel-specific)
        for x, y in zip(*outs):
            assert torch.equal(x, y)
        assert int(outs[0][6][0]) == (n + 96) // 97


def test_default_mean_table_is_bit_identical_to_rederived_means(monkeypatch):
    """ef_bandit_step_means: a default arm's per-env mean looked up in the [n,k] table of ef_bandit_default_means must give
    the bits of the in-kernel re-derivation (mean_table=False), in the sorted AND the plain kernel, for step, K-step
    rollouts with given / in-kernel random actions, a non-zero env_offset, ragged n, and after reset(seed)."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}      # arms 8 and 9 stay default normals (per-env means)
    arms[0]["param_functions"] = [{"function": lambda t: 0.1 * t, "target_param": "mean"}]
    for n, off in ((1, 0), (4097, 0), (100_003, 12_345)):
        outs = []
        for table, unsorted in ((True, False), (False, False), (True, True)):
            if unsorted:
                monkeypatch.setenv("ENVFORGE_BANDIT_UNSORTED", "1")
            else:
                monkeypatch.delenv("ENVFORGE_BANDIT_UNSORTED", raising=False)
            env = VecKArmedBandit(n, k=10, arm_params=arms, seed=5, env_offset=off, mean_table=table)
            assert (env._means_ptr() is not None) == table
            act = torch.from_numpy(np.random.default_rng(n).integers(0, 10, n).astype(np.int32)).cuda()
            act[::3] = 8 + (torch.arange(0, n, 3, device="cuda") % 2).to(torch.int32)       # plenty of default-arm pulls
            _, r1, *_ = env.step(act)
            r1 = r1.clone()
            _, r2 = env.rollout(3)
            a3 = torch.from_numpy(np.random.default_rng(n + 1).integers(0, 10, (2, n)).astype(np.int32)).cuda()
            _, r3 = env.rollout(2, a3)
            env.reset(seed=6)                                 # re-keys the generator: the table must follow
            _, r4, *_ = env.step(act)
            outs.append([r1, r2.clone(), r3.clone(), r4.clone(), env.default_means.clone(), env.stats[:2].clone(), env.stats[4].clone()])
        for other in outs[1:]:
            for x, y in zip(outs[0], other):
                assert torch.equal(x, y)
        assert not torch.equal(outs[0][0], outs[0][3])        # the new seed did change the rewards
