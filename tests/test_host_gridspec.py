"""Host logic of the product (no GPU): the table compiler of rl_envs_forge_b200.grid_world against the oracle's dict
MDP, the reference's golden MDP dumps / P tensors, and the reference's validation behaviour."""
import numpy as np
import pytest

from oracle.gridworld import FlatTables, GridWorldPort
from rl_envs_forge_b200 import _lib
from rl_envs_forge_b200.grid_world import Action, GridSpec, validate_args
from tests._golden import golden_files, grid_kwargs, grid_runtime_specials, load

FILES = golden_files("gridworld")


def _spec_kwargs(g):
    kw = grid_kwargs(g)
    if kw.get("special_transitions"):
        kw["special_transitions"] = {(k[0], Action(k[1])): v for k, v in kw["special_transitions"].items()}
    return kw


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_compiled_tables_match_reference_mdp(path):
    g = load(path)
    spec = GridSpec(**_spec_kwargs(g))
    assert np.array_equal(spec.probability_matrix(), g["P"])
    for fs, a, ts, r in grid_runtime_specials(g):
        spec.add_special_transition(fs, Action(a), ts, r)
    cols = spec.cols
    assert np.array_equal(spec.count, g["mdp_count"])
    for row in range(len(spec.count)):
        for k in range(int(spec.count[row])):
            assert (spec.nxt[row, k] // cols, spec.nxt[row, k] % cols) == tuple(g["mdp_next"][row, k])
            assert spec.reward[row, k] == g["mdp_reward"][row, k]
            assert spec.done[row, k] == g["mdp_done"][row, k]
            assert spec.prob_row[row, k] == g["mdp_prob"][row, k]
    # mdp dict export has the reference's shape
    mdp = spec.mdp_dict()
    assert len(mdp) == int((g["mdp_count"] > 0).sum())
    # P stays the construction-time tensor even after runtime patches (grid_world.py:98-99, 214-243)
    assert np.array_equal(spec.probability_matrix(), g["P"])


@pytest.mark.parametrize("path", FILES, ids=lambda p: p.split("gridworld_")[-1][:-4])
def test_integer_thresholds_reproduce_numpy_choice(path):
    """Decode the packed device table on the host exactly like the kernel does and replay the golden trajectory."""
    g = load(path)
    spec = GridSpec(**_spec_kwargs(g))
    for fs, a, ts, r in grid_runtime_specials(g):
        spec.add_special_transition(fs, Action(a), ts, r)
    tab = spec.device_table()
    slots = spec.n_slots
    r32 = np.floor(g["u"] * 2.0 ** 32).astype(np.uint64)
    pre = g["pre"][:, 0] * spec.cols + g["pre"][:, 1]
    row = pre * 4 + g["action"]
    if slots == 4:
        idx = sum((r32 >= t).astype(np.int64) for t in spec.thr)
        idx = np.minimum(idx, spec.idx_cap)
    else:
        idx = np.zeros(len(row), np.int64)
    w0, w1 = tab[row * slots + idx, 0], tab[row * slots + idx, 1]
    flags = w0 >> 16
    raised = (flags & (_lib.EF_GRID_NO_OUTCOMES | _lib.EF_GRID_BAD_PROBS)) != 0
    assert np.array_equal(raised, g["raised"] > 0)
    assert np.array_equal(((flags & _lib.EF_GRID_BAD_PROBS) != 0), g["raised"] == 1)
    ok = ~raised
    post = g["post"][:, 0] * spec.cols + g["post"][:, 1]
    assert np.array_equal((w0 & 0xFFFF)[ok], post[ok])
    assert np.array_equal(w1.view(np.float32)[ok], g["reward"][ok].astype(np.float32))
    assert np.array_equal(((flags & _lib.EF_GRID_DONE) != 0)[ok], g["done"][ok])


def test_thresholds_exhaustive_boundaries():
    """idx = min(sum(r >= thr_k), cap) equals searchsorted(cdf, r*2^-32, 'right') at and around every threshold."""
    for p in (0.0, 0.25, 0.5, 0.7, 0.8, 0.9, 1 - 2 ** -30, 1 - 2 ** -53, float(np.nextafter(1.0, 0.0))):
        spec = GridSpec(p_success=p)
        cdf = np.cumsum([p] + [(1 - p) / 3] * 3)
        cdf /= cdf[-1]
        rs = {0, 1, 2 ** 32 - 1, 2 ** 32 - 2, 2 ** 31}
        for t in spec.thr:
            rs |= {max(t - 1, 0), t, min(t + 1, 2 ** 32 - 1)}
        rs |= set(int(x) for x in np.random.default_rng(0).integers(0, 2 ** 32, 2000))
        for r in rs:
            want = int(np.searchsorted(cdf, r * 2.0 ** -32, side="right"))
            got = min(sum(r >= t for t in spec.thr), spec.idx_cap)
            assert got == want, (p, r)
            assert got <= 3


def test_matches_oracle_flat_tables_on_random_configs():
    from oracle.gen_golden import random_grid_cfg
    rng = np.random.default_rng(77)
    for _ in range(40):
        cfg = random_grid_cfg(rng)
        port = GridWorldPort(**cfg)
        ft = FlatTables(port)
        kw = dict(cfg)
        kw["special_transitions"] = {(k[0], Action(k[1])): v for k, v in cfg["special_transitions"].items()}
        spec = GridSpec(**kw)
        assert np.array_equal(spec.count, ft.count)
        assert np.array_equal(spec.bad, ft.raises)
        for k in range(4):
            live = ft.count > k
            assert np.array_equal(spec.nxt[live, k], ft.nxt[live, k])
            assert np.array_equal(spec.reward[live, k], ft.reward[live, k])
            assert np.array_equal(spec.done[live, k], ft.done[live, k])


def test_validation_errors_match_reference_cases():
    """The 15 constructor cases of tests/envs/grid_world/test_grid_world.py:168-224."""
    base = dict(rows=5, cols=5, start_state=(0, 0), terminal_states={(3, 4): 1.0}, walls=None,
                special_transitions=None, rewards=None, seed=None, slip_distribution=None, p_success=1.0,
                episode_length_limit=None)
    bad = [dict(rows=0), dict(rows="5"), dict(cols=-1), dict(start_state=(0,)), dict(start_state=[0, 0]),
           dict(terminal_states=[(3, 4)]), dict(terminal_states={(3, 4): "x"}), dict(walls=[(1, 1)]),
           dict(walls={(1,)}), dict(special_transitions={((0, 0), 1): ((1, 1), 0.5)}),
           dict(special_transitions=[1]), dict(rewards={1: 2}), dict(seed="a"),
           dict(slip_distribution={0: 0.1}), dict(p_success=1.5), dict(p_success="x"),
           dict(episode_length_limit=1.5)]
    for b in bad:
        with pytest.raises(ValueError):
            validate_args(**{**base, **b})
    validate_args(**{**base, "rewards": {}, "walls": set(), "special_transitions": {}, "slip_distribution": {}})
    validate_args(**{**base, "special_transitions": {((0, 0), Action.UP): ((1, 1), 0.5)}})


def test_empty_rewards_dict_uses_reference_defaults():
    """rewards={} is accepted (test_grid_world.py:211); `rewards or defaults` then restores the default dict (:83-88)."""
    spec = GridSpec(rewards={})
    assert spec.reward[0 * 4 + 0, 0] == -1 and spec.reward[0 * 4 + 1, 0] == -0.1


def test_transition_probs_export_matches_reference():
    """`transition_probs` (grid_world.py:82, :401-430 -- the second, differently normalised model next to `mdp`) against the
    dicts the unmodified reference built for 13 configurations (tests/golden/tprobs_cases.npz, `gen_golden tprobs`):
    same keys, same destinations, bit-equal probabilities."""
    import json
    import os
    from tests._golden import GOLDEN
    g = load(os.path.join(GOLDEN, "tprobs_cases.npz"))
    names = json.loads(str(g["names"]))
    assert len(names) >= 13
    for name in names:
        kw = _spec_kwargs({"config": g[name + "_config"]})
        p = kw.get("p_success", 1.0)
        slip = json.loads(str(g[name + "_slip"]))
        slip = ({Action(int(a)): w for a, w in slip.items()} if slip is not None
                else {a: (1.0 - p) / (len(Action) - 1) for a in Action})
        spec = GridSpec(**kw)
        tp = spec.transition_probs(slip)
        S = spec.n_cells
        present = np.zeros(S * 4, bool)
        dense = np.zeros((S * 4, S))
        for (st, a), outs in tp.items():
            assert isinstance(a, Action) and isinstance(st, tuple)
            row = (st[0] * spec.cols + st[1]) * 4 + int(a)
            present[row] = True
            for to, pr in outs.items():
                dense[row, to[0] * spec.cols + to[1]] = pr
        assert np.array_equal(present, g[name + "_present"]), name
        assert np.array_equal(dense, g[name + "_tp"]), name
