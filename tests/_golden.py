"""Helpers to load tests/golden/*.npz (written by oracle/gen_golden.py from the live reference)."""
import glob
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def load(path):
    return dict(np.load(path, allow_pickle=False))


def grid_kwargs(g):
    """Constructor kwargs (reference types: tuples / set / dicts) from a gridworld golden file."""
    c = json.loads(str(g["config"]))
    kw = dict(c)
    if "start_state" in kw:
        kw["start_state"] = tuple(kw["start_state"])
    if kw.get("walls") is not None:
        kw["walls"] = {tuple(w) for w in kw["walls"]}
    if kw.get("terminal_states") is not None:
        kw["terminal_states"] = {tuple(k): v for k, v in kw["terminal_states"]}
    if kw.get("special_transitions") is not None:
        kw["special_transitions"] = {(tuple(fs), a): (tuple(ts), r) for fs, a, ts, r in kw["special_transitions"]}
    return kw


def grid_runtime_specials(g):
    return [(tuple(fs), a, None if ts is None else tuple(ts), r)
            for fs, a, ts, r in json.loads(str(g["runtime_specials"]))]
