"""GPU parity tests of the stream-exact bandit mode (rng="mt19937" -> ef_bandit_step_mt19937): every env owns numpy's
legacy RandomState(seed_i) and the device runs numpy's legacy float64 samplers on it.  Ground truth: the reward streams
the UNMODIFIED reference produced (tests/golden/bandit_*.npz) and numpy's own RandomState (a third-party dependency of the
reference that is installed on the GPU box too), with NO injected draws and no oracle in between.

Tolerance: CUDA's double log/exp/pow/sqrt may differ from glibc's in the last ulp or two, so float64 samples are compared
at rtol 1e-13 and the public fp32 rewards at rtol 1e-6; the generator STATE afterwards (624 words, position, gauss cache
flag) is compared exactly -- a single accept/reject flip would desynchronise it."""
import numpy as np
import pytest

from oracle.gen_golden import BANDIT_ARMS, bandit_shift_functions
from tests._golden import GOLDEN, load

pytestmark = pytest.mark.gpu

CASES = [("normal", dict(mean=5, std=1)), ("uniform", dict(low=4, high=6)), ("exponential", dict(scale=0.2)),
         ("gamma", dict(shape=9, scale=0.55)), ("gamma", dict(shape=0.4, scale=2.0)), ("gamma", dict(shape=1.0, scale=2.0)),
         ("beta", dict(a=2, b=5)), ("beta", dict(a=0.5, b=0.7)), ("logistic", dict(loc=5, scale=0.3)),
         ("weibull", dict(a=2)), ("weibull", dict(a=0.7)), ("lognormal", dict(mean=1.6, sigma=0.3))]
NUMPY_ARGS = {"normal": ("mean", "std"), "uniform": ("low", "high"), "exponential": ("scale",), "gamma": ("shape", "scale"),
              "beta": ("a", "b"), "logistic": ("loc", "scale"), "weibull": ("a",), "lognormal": ("mean", "sigma")}


def _replay_golden(name, arm_params):
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    g = load(f"{GOLDEN}/bandit_{name}.npz")
    env = VecKArmedBandit(1, k=int(g["k"]), arm_params=arm_params, seed=int(g["seed"]), rng="mt19937")
    resets = set(int(x) for x in g["reset_at"])
    for i, a in enumerate(g["action"]):
        if i in resets:
            env.reset()
        obs, rew, done, trunc, _ = env.step(torch.tensor([int(a)], dtype=torch.int32, device="cuda"))
        r64 = float(env.reward64[0])
        want = float(g["reward"][i])
        assert abs(r64 - want) <= 1e-13 * abs(want), (i, r64, want)
        assert float(rew[0]) == pytest.approx(np.float32(want), rel=1e-6)
        assert int(obs[0]) == int(a) and bool(done[0]) and not bool(trunc[0])
        assert env.timestep == int(g["timestep"][i])
        arm0 = env.arms[0]
        m = float(env.default_means[0, 0]) if arm0.per_env_mean else arm0.current_params.get("mean", np.nan)
        assert (np.isnan(m) and np.isnan(g["arm0_mean"][i])) or m == g["arm0_mean"][i]


def test_reference_golden_default_seed0_without_injection():
    _replay_golden("default_seed0", None)


def test_reference_golden_config3_with_shift_and_reset_without_injection():
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[0]["param_functions"] = bandit_shift_functions()
    _replay_golden("config3", arms)


def test_reference_golden_small_shapes_without_injection():
    arms = {0: dict(distribution="gamma", shape=0.4, scale=2.0), 1: dict(distribution="beta", a=0.5, b=0.7),
            2: dict(distribution="gamma", shape=1.0, scale=3.0), 3: dict(distribution="weibull", a=0.7)}
    _replay_golden("small_shapes", arms)


def test_readme_known_answer():
    """rl_envs_forge/envs/k_armed_bandit/README.md:14-19: KArmedBandit(k=10, seed=0).step(1) -> 0.5442007795281013."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    env = VecKArmedBandit(1, k=10, seed=0, rng="mt19937")
    want = np.array([1.764052, 0.400157, 0.978738, 2.240893, 1.867558, -0.977278, 0.950088, -0.151357, -0.103219, 0.410599])
    assert np.allclose(env.default_means.cpu().numpy()[0], want, atol=1e-6)
    obs, rew, done, trunc, info = env.step(torch.tensor([1], dtype=torch.int32, device="cuda"))
    assert abs(float(env.reward64[0]) - 0.5442007795281013) < 1e-15
    assert float(rew[0]) == np.float32(0.5442007795281013) and int(obs[0]) == 1 and info == {}


def test_batch_against_numpy_randomstate_all_families():
    """1536 bandits x 12 arms covering every family and both branches of the gamma / beta samplers, 450 steps (several
    regenerations of each 624-word state): samples vs numpy's own RandomState(seed + i), then the generator states."""
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n, steps, seed, off, k = 1536, 450, 1000, 40, len(CASES)
    arms = {j: dict(distribution=fam, **kw) for j, (fam, kw) in enumerate(CASES)}
    env = VecKArmedBandit(n, k=k, arm_params=arms, seed=seed, env_offset=off, rng="mt19937")
    refs = [np.random.RandomState(seed + off + i) for i in range(n)]
    means = np.stack([r.randn(k) for r in refs])
    assert np.array_equal(env.default_means.cpu().numpy(), means)
    rng = np.random.default_rng(0)
    worst = 0.0
    for t in range(steps):
        act = rng.integers(0, k, n).astype(np.int32)
        obs, rew, _, _, _ = env.step(torch.from_numpy(act).cuda())
        want = np.array([getattr(refs[i], CASES[a][0])(*[CASES[a][1][p] for p in NUMPY_ARGS[CASES[a][0]]])
                         for i, a in enumerate(act)])
        got = env.reward64.cpu().numpy()
        err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-13, (t, int(err.argmax()), got[err.argmax()], want[err.argmax()])
        assert np.allclose(rew.cpu().numpy(), want.astype(np.float32), rtol=1e-6, atol=0)
        assert np.array_equal(obs.cpu().numpy(), act)
    # generator state: every stream still in step with numpy's
    key = env._mt_key.cpu().numpy().view(np.uint32)
    pos, hg, gs = env._mt_pos.cpu().numpy(), env._mt_has_gauss.cpu().numpy(), env._mt_gauss.cpu().numpy()
    for i in range(n):
        st = refs[i].get_state()
        assert np.array_equal(key[:, i], st[1]) and pos[i] == st[2] and hg[i] == st[3], i
        assert gs[i] == pytest.approx(st[4], rel=1e-13, abs=0), i
    st = env.episode_stats()
    assert int(st["episodes"]) == n * steps and int(st["env_steps"]) == n * steps
    print("worst relative difference of a float64 sample:", worst)


def test_default_arms_shards_seed_list_reset_and_invalid_actions():
    import torch
    from rl_envs_forge_b200 import VecKArmedBandit
    n, k = 300, 5                                                    # odd k: the gauss cache is live after randn(k)
    whole = VecKArmedBandit(n, k=k, seed=77, rng="mt19937")
    lo = VecKArmedBandit(100, k=k, seed=77, env_offset=0, rng="mt19937")
    hi = VecKArmedBandit(200, k=k, seed=77, env_offset=100, rng="mt19937")
    listed = VecKArmedBandit(n, k=k, seed=[77 + i for i in range(n)], rng="mt19937")
    refs = [np.random.RandomState(77 + i) for i in range(n)]
    means = np.stack([r.randn(k) for r in refs])
    rng = np.random.default_rng(1)
    first = []
    for t in range(40):
        act = rng.integers(0, k, n).astype(np.int32)
        bad = rng.random(n) < 0.05
        sent = act.copy()
        sent[bad] = rng.choice(np.array([-1, k, 99], np.int32), int(bad.sum()))
        d = torch.from_numpy(sent).cuda()
        r_whole = whole.step(d)[1].clone()
        r_lo, r_hi = lo.step(d[:100])[1], hi.step(d[100:])[1]
        r_list = listed.step(d)[1]
        assert torch.equal(r_whole[:100], r_lo) and torch.equal(r_whole[100:], r_hi) and torch.equal(r_whole, r_list)
        want = np.array([0.0 if bad[i] else refs[i].normal(means[i, a], 1) for i, a in enumerate(act)])   # :162-165
        assert np.allclose(whole.reward64.cpu().numpy(), want, rtol=1e-13, atol=0)      # rejected: no draw consumed
        first.append(r_whole.cpu().numpy())
    assert int(whole.err.cpu().numpy()[0]) > 0
    with pytest.raises(ValueError):
        whole.check_errors()
    # reset(): RandomState(seed) is rebuilt (:215) -- same means, same reward stream; timestep keeps counting (:202-217)
    whole.reset()
    assert whole.timestep == 40 and np.array_equal(whole.default_means.cpu().numpy(), means)
    rng = np.random.default_rng(1)
    for t in range(40):
        act = rng.integers(0, k, n).astype(np.int32)
        bad = rng.random(n) < 0.05
        sent = act.copy()
        sent[bad] = rng.choice(np.array([-1, k, 99], np.int32), int(bad.sum()))
        assert np.array_equal(whole.step(torch.from_numpy(sent).cuda())[1].cpu().numpy(), first[t])
    whole.err.zero_()
    # host arrays in, numpy out; rollout = step sequence
    twin = VecKArmedBandit(n, k=k, seed=5, rng="mt19937")
    roll = VecKArmedBandit(n, k=k, seed=5, rng="mt19937")
    acts = rng.integers(0, k, (6, n)).astype(np.int32)
    obs_r, rew_r = roll.rollout(6, torch.from_numpy(acts).cuda())
    for j in range(6):
        o, r, d, tr, _ = twin.step(acts[j])
        assert isinstance(r, np.ndarray) and np.array_equal(r, rew_r[j].cpu().numpy()) and np.array_equal(o, acts[j])
    with pytest.raises(ValueError):
        VecKArmedBandit(4, seed=2 ** 32 - 2, rng="mt19937")           # seed + 3 leaves numpy's seed range
    with pytest.raises(ValueError):
        VecKArmedBandit(VecKArmedBandit.MT19937_MAX_ENVS + 1, rng="mt19937")
    with pytest.raises(ValueError):
        VecKArmedBandit(4, rng="xorshift")
    strict = VecKArmedBandit(2, k=3, seed=1, rng="mt19937", strict=True)
    with pytest.raises(AssertionError):
        strict.step(np.array([0, 3]))
