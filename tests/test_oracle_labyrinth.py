"""Pins oracle/labyrinth.py against mazes generated and stepped by the unmodified reference (tests/golden/labyrinth_*.npz,
written by oracle/gen_golden.py): every step's position, reward, done flag and full observation."""
import json

import numpy as np
import pytest

from oracle import labyrinth as OL
from tests._golden import golden_files, load

FILES = golden_files("labyrinth")


def test_golden_present():
    assert len(FILES) >= 5


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("labyrinth_")[-1][:-4])
def test_scalar_port_replays_reference(path):
    g = load(path)
    kw = json.loads(str(g["config"]))
    env = OL.LabyrinthPort(g["grid"], g["start"], g["target"], kw.get("state_vision_range"))
    assert np.array_equal(env.state, g["initial_obs"])
    assert g["done"].sum() >= 5 and (g["info"] == 1).sum() >= 100
    for pre, a, post, r, d, obs, info in zip(g["pre"], g["action"], g["post"], g["reward"], g["done"], g["obs"], g["info"]):
        env.pos = tuple(int(x) for x in pre)
        o, rew, done, trunc, inf = env.step(int(a))
        assert env.pos == tuple(int(x) for x in post) and rew == r and done == bool(d) and trunc is False
        assert np.array_equal(o, obs)
        assert {"": 0, "Invalid move!": 1, "Reached the target!": 2}[inf.get("info", "")] == int(info)


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("labyrinth_")[-1][:-4])
def test_vector_port_equals_reference(path):
    """All recorded steps as one batch (each step = one env on the same maze)."""
    g = load(path)
    kw = json.loads(str(g["config"]))
    rows, cols = g["grid"].shape
    grids = g["grid"].reshape(1, -1)
    n = len(g["action"])
    mi = np.zeros(n, np.int64)
    tgt = np.array([g["target"][0] * cols + g["target"][1]])
    pre = g["pre"][:, 0] * cols + g["pre"][:, 1]
    pos, rew, done, invalid = OL.vec_step(grids, mi, tgt, pre, g["action"], cols)
    assert not invalid.any()
    assert np.array_equal(pos, g["post"][:, 0] * cols + g["post"][:, 1])
    assert np.array_equal(rew, g["reward"].astype(np.float32)) and np.array_equal(done, g["done"])
    assert np.array_equal(OL.vec_obs(grids, mi, tgt, pos, cols, kw.get("state_vision_range")), g["obs"])


def test_vector_port_invalid_actions_and_reference_test_cases():
    W, P = OL.WALL, OL.PATH
    pre = np.array([[W, W, W, W, W], [W, P, P, P, W], [W, P, W, P, W], [W, P, P, P, W], [W, W, W, W, W]], np.uint8)
    env = OL.LabyrinthPort(pre, (1, 1), (3, 3))                       # test_labyrinth.py:296-314
    _, r, d, _, _ = env.step(1)                                       # test_valid_move :166-179
    assert env.pos == (1, 2) and r == -0.01 and not d
    _, r, d, _, _ = env.step(0)                                       # test_invalid_move :181-191
    assert env.pos == (1, 2) and r == -1
    env.pos = (2, 3)
    _, r, d, _, _ = env.step(2)                                       # test_reaching_target :193-203
    assert r == 10 and d
    with pytest.raises(ValueError):
        env.step(4)
    pos, rew, done, invalid = OL.vec_step(pre.reshape(1, -1), np.zeros(3, np.int64), np.array([18]), np.array([6, 6, 13]),
                                          np.array([7, -1, 2]), 5)
    assert invalid.tolist() == [True, True, False] and pos.tolist() == [6, 6, 18] and done.tolist() == [False, False, True]
    # vision window in a corner: out-of-bounds cells read as WALL (test_state_vision_range_2_corner :368-420)
    pre2 = np.zeros((10, 10), np.uint8)
    pre2[0:3, 0:3] = P
    env = OL.LabyrinthPort(pre2, (0, 0), (2, 2), state_vision_range=2)
    s = env.state
    assert s.shape == (5, 5) and s[2, 2] == OL.PLAYER and s[4, 4] == OL.TARGET and (s[:2, :] == W).all() and (s[:, :2] == W).all()
