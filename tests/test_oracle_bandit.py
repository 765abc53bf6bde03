"""Pins oracle/bandit.py: legacy RandomState algorithms (bit-exact vs numpy) and the KArmedBandit restatement
against golden rewards produced by the unmodified reference."""
import numpy as np
import pytest

from oracle.bandit import BanditPort, LegacyDistributions, MTPrimitives, PhiloxPrimitives, engine_sample
from oracle.gen_golden import BANDIT_ARMS, bandit_shift_functions
from tests._golden import GOLDEN, load

CASES = [("normal", (5, 1)), ("uniform", (4, 6)), ("exponential", (0.2,)), ("gamma", (9, 0.55)),
         ("gamma", (0.4, 2.0)), ("gamma", (1.0, 2.0)), ("beta", (2, 5)), ("beta", (0.5, 0.7)),
         ("logistic", (5, 0.3)), ("weibull", (2,)), ("weibull", (0.7,)), ("lognormal", (1.6, 0.3))]


@pytest.mark.parametrize("fam,args", CASES)
def test_legacy_family_bit_exact_vs_numpy(fam, args):
    ref = np.random.RandomState(123)
    mine = LegacyDistributions(MTPrimitives(np.random.RandomState(123)))
    for _ in range(1500):
        assert getattr(ref, fam)(*args) == getattr(mine, fam)(*args)


def test_legacy_interleaved_bit_exact_vs_numpy():
    ref = np.random.RandomState(7)
    mine = LegacyDistributions(MTPrimitives(np.random.RandomState(7)))
    pick = np.random.default_rng(0)
    for _ in range(3000):
        fam, args = CASES[pick.integers(len(CASES))]
        assert getattr(ref, fam)(*args) == getattr(mine, fam)(*args)
    assert np.array_equal(ref.randn(7), mine.randn(7))


def test_readme_known_answer():
    """rl_envs_forge/envs/k_armed_bandit/README.md:14-19."""
    assert BanditPort(k=10, seed=0).step(1) == (1, 0.5442007795281013, True, False, {})


def _replay(g, arm_params):
    env = BanditPort(k=int(g["k"]), arm_params=arm_params, seed=int(g["seed"]))
    resets = set(int(x) for x in g["reset_at"])
    for i, a in enumerate(g["action"]):
        if i in resets:
            env.reset()
        obs, r, d, t, _ = env.step(int(a))
        assert obs == a and r == g["reward"][i] and d is True and t is False
        assert env.timestep == g["timestep"][i]
        m = env.arms[0].current_params.get("mean", np.nan)
        assert (np.isnan(m) and np.isnan(g["arm0_mean"][i])) or m == g["arm0_mean"][i]


def test_golden_default_seed0():
    _replay(load(f"{GOLDEN}/bandit_default_seed0.npz"), None)


def test_golden_config3_with_shift_and_reset():
    arms = {i: dict(p) for i, p in BANDIT_ARMS.items()}
    arms[0]["param_functions"] = bandit_shift_functions()
    _replay(load(f"{GOLDEN}/bandit_config3.npz"), arms)


def test_golden_small_shapes():
    arms = {0: dict(distribution="gamma", shape=0.4, scale=2.0), 1: dict(distribution="beta", a=0.5, b=0.7),
            2: dict(distribution="gamma", shape=1.0, scale=3.0), 3: dict(distribution="weibull", a=0.7)}
    _replay(load(f"{GOLDEN}/bandit_small_shapes.npz"), arms)


def test_reference_test_suite_pins():
    """tests/envs/k_armed_bandit/test_k_armed_bandit.py:43-60, :89-95, :97-129."""
    with pytest.raises(ValueError):
        BanditPort(k=3, arm_params={3: dict(distribution="normal", mean=0, std=1)})
    with pytest.raises(ValueError):
        BanditPort(k=1, arm_params={0: dict(distribution="cauchy")})
    env = BanditPort(k=1, seed=0, arm_params={0: dict(distribution="normal", mean=0, std=1, param_functions=[
        {"function": lambda t: t, "target_param": "mean"}])})
    for _ in range(10):
        env.step(0)
    assert env.arms[0].current_params["mean"] == 10
    env.reset()
    assert env.timestep == 10  # timestep survives reset (k_armed_bandit.py:202-217)


def test_engine_stream_moments():
    """The Philox-driven primitive source produces sane distributions (sanity of the injection layer)."""
    xs = np.array([engine_sample(1, g, 0, "gamma", 9.0, 0.55)[0] for g in range(4000)])
    assert abs(xs.mean() - 9 * 0.55) < 0.1 and abs(xs.var() - 9 * 0.55 ** 2) < 0.3
    xs = np.array([engine_sample(1, g, 3, "normal", 5.0, 1.0)[0] for g in range(4000)])
    assert abs(xs.mean() - 5) < 0.06 and abs(xs.std() - 1) < 0.05
    p = PhiloxPrimitives(1, 0, 0)
    assert 0 < p.u() < 1 and p.consumed == 1


@pytest.mark.parametrize("fam,args", CASES)
def test_vectorised_engine_oracle_matches_scalar(fam, args):
    from oracle.bandit import engine_sample_vec
    p0 = float(args[0])
    p1 = float(args[1]) if len(args) > 1 else 0.0
    g = np.arange(300) + 1000
    v, used = engine_sample_vec(42, g, 5, fam, p0, p1)
    for i in range(0, 300, 7):
        s, c = engine_sample(42, int(g[i]), 5, fam, p0, p1)
        assert abs(s - v[i]) <= 4e-16 * abs(s) and c == used[i]  # np.power vs math.pow may differ in the last bit
