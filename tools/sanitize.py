#!/usr/bin/env python
"""Small invocation of every kernel family and code path (all C-ABI entry points, odd and tiny sizes).  Runs stand-alone
as a quick whole-library smoke; written to run under compute-sanitizer (closed on the shared B200 pool this was built on):
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
    compute-sanitizer --tool synccheck python tools/sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl_envs_forge_b200 import (Action, VecCartPole, VecGamblersProblem, VecGridWorld, VecGridWorldLayouts,  # noqa: E402
                                VecJacksCarRental, VecKArmedBandit, VecLabyrinth, VecPendulumDisk)

walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
lay = dict(start_state=(0, 0), walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0},
           special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
rng = np.random.default_rng(0)
for n in (5, 4099):
    ai = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
    af = torch.empty(n, device="cuda").uniform_(-3, 3)
    for kw in (dict(p_success=0.8, episode_length_limit=20), dict(p_success=1.0), dict(p_success=0.8, episode_length_limit=20, rng="pcg64")):
        e = VecGridWorld(n, rows=8, cols=8, seed=1, **lay, **kw)
        for _ in range(3):
            e.step(ai, return_final_obs=True)
            e.step(ai)
        e.step(ai.cpu().numpy())
        if kw.get("rng") != "pcg64":
            e.step(ai, draws=torch.randint(0, 2 ** 31, (n,), dtype=torch.int32, device="cuda"))
            e.rollout(5)
            e.rollout(3, torch.randint(0, 4, (3, n), dtype=torch.int32, device="cuda"), return_rewards=True)
        e.reset(mask=ai > 1)
        e.running_episode_stats()
    # narrow wire format (vector + scalar kernels, zero-copy / staged / pipelined host paths) with wild action bytes
    a8 = torch.randint(0, 4, (n,), dtype=torch.uint8, device="cuda")
    a8[::7] = 251
    for kw in (dict(p_success=0.8, episode_length_limit=20), dict(p_success=1.0, autoreset=False)):
        e = VecGridWorld(n, rows=8, cols=8, seed=1, wire="compact", **lay, **kw)
        for _ in range(3):
            e.step(a8, return_final_obs=True)
            e.step(a8)
        e.step(a8, draws=torch.randint(0, 2 ** 31, (n,), dtype=torch.int32, device="cuda"))
        for tr in ("zerocopy", "staged", "pipelined"):
            e.host_transport = tr
            e.step(a8.cpu().numpy(), return_final_obs=True)
    wild = ai.clone()
    wild[::5] = torch.tensor([1 << 20, -7, (1 << 31) - 1, 4], dtype=torch.int32, device="cuda").repeat(n)[: wild[::5].numel()]
    e = VecGridWorld(n, rows=8, cols=8, seed=1, p_success=0.8, **lay)
    e.step(wild)
    e.step(wild, return_final_obs=True)
    e.rollout(3, wild.repeat(3, 1), return_rewards=True)
    big = VecGridWorld(n, rows=120, cols=120, seed=1, p_success=0.9, terminal_states={(119, 119): 1.0})  # global-memory table
    big.step(ai)
    big.rollout(2)
    m = VecGridWorldLayouts(n, [lay, dict(start_state=(2, 2), terminal_states={(5, 5): 2.0})], rows=8, cols=8, p_success=0.8,
                            episode_length_limit=9, seed=2)
    for _ in range(3):
        m.step(ai)
    m.step(ai, draws=torch.randint(0, 2 ** 31, (n,), dtype=torch.int32, device="cuda"))
    m.step(wild)
    m.rollout(4)
    m.rollout(3, wild.repeat(3, 1), return_obs=True)
    m2 = VecGridWorldLayouts(n, [lay] * 3, rows=8, cols=8, p_success=1.0, episode_length_limit=9, seed=2,
                             layout_index=rng.integers(0, 3, n), state_layout="split")
    m2.rollout(4, return_rewards=True)
    for cls in (VecCartPole, VecPendulumDisk):
        e = cls(n, seed=3, episode_length_limit=6)
        for _ in range(3):
            e.step(af)
            e.step(af, return_final_obs=True, return_info=True)
        e.step(af.cpu().numpy())
        e.rollout(8)
        e.rollout(4, torch.empty((4, n), device="cuda").uniform_(-3, 3), return_rewards=True)
        e.reset(mask=ai > 1)
    arms = {0: dict(distribution="normal", mean=5, std=1), 1: dict(distribution="gamma", shape=9, scale=0.55),
            2: dict(distribution="beta", a=2, b=5), 3: dict(distribution="gamma", shape=0.4, scale=2.0),
            4: dict(distribution="beta", a=0.5, b=0.7), 5: dict(distribution="weibull", a=2)}
    b = VecKArmedBandit(n, k=8, arm_params=arms, seed=4)
    ab = torch.randint(0, 9, (n,), dtype=torch.int32, device="cuda")     # includes rejected actions
    b.step(ab)
    b.rollout(3)
    b.step(ab.clamp(max=7).cpu().numpy())
    bm = VecKArmedBandit(min(n, 700), k=8, arm_params=arms, seed=4, rng="mt19937")     # stream-exact mode (MT19937 per env)
    for _ in range(120):                                                              # past one 624-word regeneration
        bm.step(ab[:bm.num_envs])
    os.environ["ENVFORGE_BANDIT_UNSORTED"] = "1"
    b.step(ab)
    del os.environ["ENVFORGE_BANDIT_UNSORTED"]
    g = VecGamblersProblem(n, goal_amount=30, win_probability=0.45, start_capital=4, episode_length_limit=7, seed=5)
    ag = torch.randint(0, 9, (n,), dtype=torch.int32, device="cuda")
    g.step(ag, return_final_obs=True)
    g.step(ag[1:n].contiguous() if n > 5 else ag) if False else None
    g.rollout(6)
    g.rollout(3, torch.randint(0, 9, (3, n), dtype=torch.int32, device="cuda"), return_rewards=True)
    c = VecJacksCarRental(n, seed=6, episode_length_limit=4)
    ac = torch.randint(0, 11, (n,), dtype=torch.int32, device="cuda")
    for _ in range(5):
        c.step(ac)
    c.rollout(5, return_rewards=True)
    grids = (rng.random((7, 12, 10)) > 0.3).astype(np.uint8)
    grids[:, 0, 0] = grids[:, 11, 9] = 1
    for vis in (None, 2, 9):
        lb = VecLabyrinth(n, grids, np.zeros((7, 2), np.int64), np.tile([11, 9], (7, 1)), state_vision_range=vis, seed=7,
                          episode_length_limit=9)
        for _ in range(3):
            lb.step(ai, return_final_obs=True)
        lb.rollout(6, return_rewards=True)
        lb.state
torch.cuda.synchronize()
print("sanitize workload done")
