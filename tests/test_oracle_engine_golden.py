"""The oracle vs the UNMODIFIED reference driven by the engine's own Philox streams (tests/golden/engine_*.npz, written by
`python -m oracle.gen_golden engine`): env i of a batch is one reference instance whose random draws are the engine's words
of (seed, global env index, step) and whose actions are the engine's random policy.  This is BASELINE.json's parity clause
("the reference's own step() driven with identical injected random draws and actions") on the CPU side; the GPU tests
(tests/test_gpu_engine_golden.py) hold the device to the same files."""
import json

import numpy as np

from oracle import acml as OA
from oracle import bandit as OB
from oracle import philox
from oracle.gen_golden import BANDIT_ARMS
from oracle.gridworld import FlatTables, GridWorldPort, vec_step
from tests._golden import GOLDEN, grid_kwargs, load


def _g(name):
    return load(f"{GOLDEN}/engine_{name}.npz")


def test_gridworld_vectorised_oracle_reproduces_engine_driven_reference():
    for name in ("gridworld_config2_p08", "gridworld_config2_p1", "gridworld_config2_p08_4096"):
        g = _g(name)
        cfg = grid_kwargs(g)
        seed, off = int(g["seed"]), int(g["env_offset"])
        T, n = g["reward"].shape
        port = GridWorldPort(**cfg)
        tab = FlatTables(port)
        gi = np.arange(n) + off
        pos, ln, ret = np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.float32)
        for t in range(T):
            wt, wp = philox.step_words(seed, gi, np.full(n, t, np.uint64))
            o = vec_step(tab, pos, ln, ret, (wp >> np.uint32(30)).astype(np.int64), wt, limit=cfg["episode_length_limit"],
                         start_pos=0, autoreset=True)
            pos, ln, ret = o["pos"], o["ep_len"], o["ep_ret"]
            flags = o["terminated"].astype(np.uint8) | (o["truncated"].astype(np.uint8) << 1) | (o["error"].astype(np.uint8) << 2)
            assert np.array_equal(o["obs"], g["obs"][t]), (name, t)
            assert np.array_equal(o["reward"], g["reward"][t]) and np.array_equal(flags, g["flags"][t]), (name, t)
        assert (g["flags"] & 3 > 0).sum() > 100


def test_gambler_oracle_reproduces_engine_driven_reference():
    g = _g("gambler_default")
    kw = json.loads(str(g["config"]))
    seed, off = int(g["seed"]), int(g["env_offset"])
    T, n = g["reward"].shape
    gi = np.arange(n) + off
    s = np.full(n, kw["start_capital"], np.int64)
    for t in range(T):
        wt, wp = philox.step_words(seed, gi, np.full(n, t, np.uint64))
        hi = np.maximum(np.minimum(s, kw["goal_amount"] - s), 0).astype(np.uint64)
        bet = ((wp.astype(np.uint64) * (hi + 1)) >> np.uint64(32)).astype(np.int64)
        s, rew, done, invalid = OA.gambler_step_vec(s, bet, wt, kw["goal_amount"], kw["win_probability"])
        assert not invalid.any() and np.array_equal(rew, g["reward"][t]) and np.array_equal(done, g["done"][t]), t
        s = np.where(done, kw["start_capital"], s)
        assert np.array_equal(s, g["state"][t]), t


def test_car_rental_oracle_reproduces_engine_driven_reference():
    g = _g("carrental_default")
    p = OA.CarRentalParams(**json.loads(str(g["config"])))
    seed, off = int(g["seed"]), int(g["env_offset"])
    T, n = g["reward"].shape
    gi = np.arange(n) + off
    s = np.full((n, 2), p.max_cars // 2, np.int64)
    for t in range(T):
        wp = philox.group_word(seed, gi, t, philox.STREAM_POLICY).astype(np.uint64)
        act = ((wp * np.uint64(2 * p.max_move_cars + 1)) >> np.uint64(32)).astype(np.int64)
        s, rew = OA.car_rental_engine_step_vec(p, seed, gi, t, s, act)
        assert np.array_equal(s, g["state"][t]) and np.array_equal(rew.astype(np.float32), g["reward"][t]), t


def test_bandit_oracle_reproduces_engine_driven_reference():
    g = _g("bandit_config3_custom_arms")
    seed, off, k = int(g["seed"]), int(g["env_offset"]), int(g["k"])
    T, n = g["reward"].shape
    gi = np.arange(n) + off
    names = {"normal": ("mean", "std"), "uniform": ("low", "high"), "exponential": ("scale",), "gamma": ("shape", "scale"),
             "beta": ("a", "b"), "logistic": ("loc", "scale"), "weibull": ("a",), "lognormal": ("mean", "sigma")}
    for t in range(T):
        wp = philox.group_word(seed, gi, t, philox.STREAM_POLICY).astype(np.uint64)
        act = ((wp * np.uint64(k)) >> np.uint64(32)).astype(np.int64)
        assert np.array_equal(act, g["action"][t])
        for a in range(k):
            sel = act == a
            if not sel.any():
                continue
            arm = BANDIT_ARMS[a]
            args = [float(arm[x]) for x in names[arm["distribution"]]]
            ref, _ = OB.engine_sample_vec(seed, gi[sel], t, arm["distribution"], args[0], args[1] if len(args) > 1 else 0.0)
            # float64; numpy's vectorised log/cos/pow may differ from the scalar libm calls of the reference by an ulp
            assert np.allclose(ref, g["reward"][t][sel], rtol=1e-13, atol=0), (t, a)
