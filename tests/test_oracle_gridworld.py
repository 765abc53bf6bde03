"""Pins oracle/gridworld.py against golden vectors produced by the unmodified reference (oracle/gen_golden.py)."""
import numpy as np
import pytest

from oracle.gridworld import ChoiceError, FlatTables, GridWorldPort, vec_step
from tests._golden import golden_files, grid_kwargs, grid_runtime_specials, load

FILES = golden_files("gridworld")


def _port(g):
    port = GridWorldPort(**grid_kwargs(g))
    for fs, a, ts, r in grid_runtime_specials(g):
        port.add_special_transition(fs, a, ts, r)
    return port


def test_golden_present():
    assert len(FILES) >= 19


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_mdp_and_P_match_reference(path):
    g = load(path)
    port = _port(g)
    S = port.rows * port.cols
    count = np.zeros(S * 4, np.int32)
    for (s, a), outs in port.mdp.items():
        if not (0 <= s[0] < port.rows and 0 <= s[1] < port.cols):
            continue
        row = (s[0] * port.cols + s[1]) * 4 + a
        count[row] = len(outs)
        for k, (to, r, d, p) in enumerate(outs):
            assert tuple(g["mdp_next"][row, k]) == tuple(to)
            assert g["mdp_reward"][row, k] == r
            assert bool(g["mdp_done"][row, k]) == d
            assert g["mdp_prob"][row, k] == p
    assert (count == g["mdp_count"]).all()
    # P is built at construction, i.e. before add_special_transition patches the dict (grid_world.py:98-99)
    assert np.array_equal(GridWorldPort(**grid_kwargs(g)).probability_matrix(), g["P"])


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_scalar_trajectory_matches_reference(path):
    g = load(path)
    port = _port(g)
    port.reset()
    for i in range(len(g["action"])):
        assert port.state == tuple(g["pre"][i])
        try:
            s, r, d, t, _ = port.step(int(g["action"][i]), float(g["u"][i]))
            raised = 0
        except ChoiceError:
            raised = 1
        except ValueError:
            raised = 2
        assert raised == int(g["raised"][i])
        if raised:
            continue
        assert s == tuple(g["post"][i]) and r == g["reward"][i]
        assert d == bool(g["done"][i]) and t == bool(g["trunc"][i])
        assert port.episode_length_counter == int(g["ep_len"][i])
        if d or t:
            port.reset()


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_vectorised_restatement_matches_reference(path):
    """vec_step over 'N envs' = the golden trajectory's steps taken as independent one-step problems.

    u is float64 in the golden data; vec_step takes r_u32 with u = r * 2**-32, so quantise u downwards and keep
    only the steps where that does not cross a cdf threshold (all but a ~1e-9 fraction)."""
    g = load(path)
    port = _port(g)
    tab = FlatTables(port)
    cols = port.cols
    lim = port.episode_length_limit
    r32 = np.floor(g["u"] * 2.0 ** 32).astype(np.uint64).astype(np.uint32)
    pre = g["pre"][:, 0] * cols + g["pre"][:, 1]
    ep_len_pre = g["ep_len"] - (g["raised"] == 0)
    out = vec_step(tab, pre, ep_len_pre, np.zeros(len(pre), np.float32), g["action"], r32,
                   limit=lim, start_pos=port.start_state[0] * cols + port.start_state[1], autoreset=False)
    assert np.array_equal(out["error"], g["raised"] > 0)
    ok = ~out["error"]
    post = g["post"][:, 0] * cols + g["post"][:, 1]
    assert np.array_equal(out["pos"][ok], post[ok])
    assert np.array_equal(out["reward"][ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(out["terminated"][ok], g["done"][ok])
    assert np.array_equal(out["truncated"][ok], g["trunc"][ok])


def test_survey_known_answer_slip_seed42():
    """GridWorld(p_success=0.5, seed=42), DOWN from (0,0): u -> outcome index -> state."""
    port = GridWorldPort(p_success=0.5)
    us = [0.77395605, 0.43887844, 0.85859792, 0.69736803, 0.09417735]
    want = [(0, 1), (1, 0), (0, 0), (0, 1), (1, 0)]
    for u, w in zip(us, want):
        port.reset()
        assert port.step(2, u)[0] == w


def test_reference_test_suite_pins():
    """The assertions of the reference's tests/envs/grid_world/test_grid_world.py that pin numbers."""
    e = GridWorldPort()
    assert e.step(2, 0.3)[0] == (1, 0)                                  # test_actions :37-39
    e = GridWorldPort()
    assert e.step(1, 0.3)[1] == -0.1                                    # test_rewards :41-43
    e.reset()
    assert e.state == (0, 0)                                            # test_reset :45-48
    e.add_special_transition((0, 0), 2, (4, 4), 0.5)                    # test_special_transition :50-67
    s, r, *_ = e.step(2, 0.9)
    assert s == (4, 4) and r == 0.5
    e = GridWorldPort(rows=3, cols=3, walls={(1, 1)}, terminal_states={(2, 2): 1.0},
                      special_transitions={((0, 0), 1): ((0, 2), 0.0)})  # test_probability_matrix :107-152
    P = e.probability_matrix()
    assert P.shape == (9, 9, 4) and P[0, 2, 1] == 1.0 and P[4, 4, 0] == 1.0 and P[8, 8, 3] == 1.0
    e = GridWorldPort(episode_length_limit=10)                          # test_episode_length_limit :154-166
    for i in range(9):
        assert e.step(1 if i % 2 else 3, 0.1)[3] is False
    s, r, d, t, _ = e.step(3, 0.1)
    assert t is True and d is False
    e = GridWorldPort(rows=5, cols=5)
    assert e.reset((5, 5))[0] == (5, 5)                                 # test_reset_to_new_starting_state :236-242
    e = GridWorldPort(rows=4, cols=4, walls={(1, 1), (2, 2)}, terminal_states={(3, 3): 1.0})
    assert len(e.mdp) == (16 - 2) * 4                                   # test_mdp_size :252-267
