"""Pins oracle/acml.py: numpy's legacy Poisson (bit-exact vs RandomState.poisson) and the GamblersProblem /
JacksCarRental restatements against golden trajectories produced by the unmodified reference (oracle/gen_golden.py)."""
import json

import numpy as np
import pytest

from oracle import acml as OA
from tests._golden import golden_files, load


@pytest.mark.parametrize("lam", [0.0, 0.3, 1.0, 2.0, 3.0, 4.0, 9.5, 9.999, 10.0, 10.5, 12.5, 37.0, 100.0, 999.0, 1e5])
def test_legacy_poisson_bit_exact_vs_numpy(lam):
    ref = np.random.RandomState(99)
    src = OA.MTDoubles(np.random.RandomState(99))
    for _ in range(2000):
        assert int(ref.poisson(lam)) == OA.legacy_poisson(src, lam)
    assert ref.random_sample() == src.double()          # both streams consumed the same number of doubles


def test_legacy_poisson_mixed_rates_share_one_stream():
    """Multiplication-method and PTRS draws interleaved on one generator, the way one JacksCarRental step mixes them."""
    ref = np.random.RandomState(4)
    src = OA.MTDoubles(np.random.RandomState(4))
    for _ in range(1500):
        for lam in (12.0, 4.0, 15.5, 0.0, 3.0, 250.0):
            assert int(ref.poisson(lam)) == OA.legacy_poisson(src, lam)
    assert ref.random_sample() == src.double()
    for x in (1.0, 2.0, 3.0, 6.5, 7.0, 12.0, 400.0):
        import math
        assert OA.random_loggam(x) == pytest.approx(math.lgamma(x), rel=1e-10, abs=1e-12)


def test_scipy_poisson_rvs_is_the_global_legacy_stream():
    """car_rental.py:89-92 calls scipy.stats.poisson.rvs: it must draw from np.random's global RandomState."""
    from scipy.stats import poisson
    np.random.seed(5)
    got = [int(poisson.rvs(lam)) for lam in (3, 4, 3, 2, 14, 30.5) * 50]
    src = OA.MTDoubles(np.random.RandomState(5))
    assert got == [OA.legacy_poisson(src, lam) for lam in (3, 4, 3, 2, 14, 30.5) * 50]


@pytest.mark.parametrize("path", golden_files("gambler"), ids=lambda p: p.split("gambler_")[-1][:-4])
def test_gambler_golden(path):
    g = load(path)
    kw = json.loads(str(g["config"]))
    goal, p = kw["goal_amount"], kw["win_probability"]
    assert len(g["action"]) >= 300 and g["done"].any()
    for pre, a, u, post, r, d in zip(g["pre"], g["action"], g["u"], g["post"], g["reward"], g["done"]):
        assert OA.gambler_step(int(pre), int(a), float(u), goal, p) == (int(post), int(r), bool(d))


def test_gambler_vec_equals_scalar():
    rng = np.random.default_rng(0)
    for goal, p in ((100, 0.4), (16, 0.5), (30, 1.0), (30, 0.0)):
        n = 5000
        state = rng.integers(0, goal + 1, n)
        action = rng.integers(-2, goal + 3, n)
        r32 = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
        s2, rew, done, invalid = OA.gambler_step_vec(state, action, r32, goal, p)
        for i in range(n):
            if invalid[i]:
                assert not (0 <= action[i] <= goal) and s2[i] == state[i] and rew[i] == 0 and not done[i]
            else:
                want = OA.gambler_step(int(state[i]), int(action[i]), float(r32[i]) * 2.0 ** -32, goal, p)
                assert (int(s2[i]), int(rew[i]), bool(done[i])) == want


@pytest.mark.parametrize("path", golden_files("carrental"), ids=lambda p: p.split("carrental_")[-1][:-4])
def test_car_rental_golden(path):
    g = load(path)
    p = OA.CarRentalParams(**json.loads(str(g["config"])))
    src = OA.MTDoubles(np.random.RandomState(int(g["seed"])))
    state = (p.max_cars // 2, p.max_cars // 2)                       # reset("equal"), car_rental.py:140-142
    for pre, a, post, r in zip(g["pre"], g["action"], g["post"], g["reward"]):
        assert state == tuple(int(x) for x in pre)
        state, reward = OA.car_rental_step(p, state, int(a), src)
        assert state == tuple(int(x) for x in post) and reward == int(r)


@pytest.mark.parametrize("lams", [((3, 4), (3, 2)), ((12, 4), (15.5, 0)), ((40, 10), (25, 33.3))])
def test_car_rental_engine_vec_equals_scalar(lams):
    p = OA.CarRentalParams(max_cars=20, max_move_cars=5, request_lambda=lams[0], return_lambda=lams[1])
    rng = np.random.default_rng(1)
    n = 400
    g = np.arange(1000, 1000 + n)
    state = rng.integers(0, 21, (n, 2))
    action = rng.integers(0, 11, n)
    s2, rew = OA.car_rental_engine_step_vec(p, 77, g, 5, state, action)
    for i in range(n):
        want = OA.car_rental_engine_step(p, 77, int(g[i]), 5, state[i], action[i])
        assert (tuple(int(x) for x in s2[i]), int(rew[i])) == want


def test_car_rental_build_mdp_matches_reference_dump():
    """The reference accumulates the expected reward term by term (car_rental.py:229-301); the numpy restatement may
    differ by floating-point association only."""
    g = load(golden_files("carrental")[0])
    p = OA.CarRentalParams(max_cars=4, max_move_cars=2, request_lambda=(2, 2), return_lambda=(2, 2))
    mdp = OA.car_rental_expected_reward_mdp(p)
    assert len(mdp) == len(g["mdp_keys"]) == 25 * 5
    for key, nxt, rew in zip(g["mdp_keys"], g["mdp_next"], g["mdp_reward"]):
        got = mdp[((int(key[0]), int(key[1])), int(key[2]))]
        assert got[0] == tuple(int(x) for x in nxt) and got[2] is False
        assert abs(got[1] - rew) <= 1e-12 * max(1.0, abs(rew))


@pytest.mark.needs_reference
def test_gambler_build_mdp_equals_reference():
    from oracle.ref_shim import load_reference
    R = load_reference()
    for goal, p in ((100, 0.4), (16, 0.5)):
        assert OA.gambler_build_mdp(goal, p) == R.GamblersProblem(goal_amount=goal, win_probability=p).build_mdp()
