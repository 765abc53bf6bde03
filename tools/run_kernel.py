#!/usr/bin/env python
"""Run one env family's kernel a few times (for ncu / quick timing).
    python tools/run_kernel.py --env cartpole|pendulum|gridworld|bandit|gambler|carrental|labyrinth|labyrinth_vision --mode step|rollout|rollout_ext --n N --reps R [--K 64]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import numpy as np  # noqa: E402

from rl_envs_forge_b200 import (Action, VecCartPole, VecGamblersProblem, VecGridWorld, VecJacksCarRental,  # noqa: E402
                                VecKArmedBandit, VecLabyrinth, VecPendulumDisk)

ap = argparse.ArgumentParser()
ap.add_argument("--env", default="cartpole")
ap.add_argument("--mode", default="step")
ap.add_argument("--n", type=int, default=1 << 22)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--K", type=int, default=64)
ap.add_argument("--sets", type=int, default=4)
ap.add_argument("--no-mean-table", action="store_true", help="bandit: re-derive default-arm means in the kernel")
a = ap.parse_args()

n, K = a.n, a.K
envs, acts = [], []
for s in range(a.sets):
    if a.env == "cartpole":
        e = VecCartPole(n, seed=1, env_offset=s * n)
        act = torch.empty(n, device="cuda").uniform_(-3, 3)
        ext = torch.empty((K, n), device="cuda").uniform_(-3, 3) if a.mode == "rollout_ext" else None
    elif a.env == "pendulum":
        e = VecPendulumDisk(n, seed=1, env_offset=s * n)
        act = torch.empty(n, device="cuda").uniform_(-3, 3)
        ext = torch.empty((K, n), device="cuda").uniform_(-3, 3) if a.mode == "rollout_ext" else None
    elif a.env == "gridworld":
        walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
        e = VecGridWorld(n, rows=8, cols=8, walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0},
                         special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)}, p_success=0.8,
                         episode_length_limit=200, seed=1, env_offset=s * n)
        act = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
        ext = torch.randint(0, 4, (K, n), dtype=torch.int32, device="cuda") if a.mode == "rollout_ext" else None
    elif a.env == "gambler":
        e = VecGamblersProblem(n, goal_amount=100, win_probability=0.4, start_capital=10, seed=1, env_offset=s * n)
        act, ext = torch.randint(0, 11, (n,), dtype=torch.int32, device="cuda"), None
    elif a.env == "carrental":
        e = VecJacksCarRental(n, seed=1, env_offset=s * n)
        act, ext = torch.randint(0, 11, (n,), dtype=torch.int32, device="cuda"), None
    elif a.env in ("labyrinth", "labyrinth_vision"):
        rng = np.random.default_rng(5)
        grids = (rng.random((64, 32, 32)) > 0.3).astype(np.uint8)
        grids[:, 0, 0] = grids[:, 31, 31] = 1
        e = VecLabyrinth(n, grids, np.zeros((64, 2), np.int64), np.full((64, 2), 31), seed=1, env_offset=s * n,
                         state_vision_range=3 if a.env == "labyrinth_vision" else None)
        act, ext = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda"), None
    else:
        arms = {0: dict(distribution="normal", mean=5, std=1), 1: dict(distribution="uniform", low=4, high=6),
                2: dict(distribution="exponential", scale=0.2), 3: dict(distribution="gamma", shape=9, scale=0.55),
                4: dict(distribution="logistic", loc=5, scale=0.3), 5: dict(distribution="lognormal", mean=1.6, sigma=0.3),
                6: dict(distribution="weibull", a=2), 7: dict(distribution="beta", a=2, b=5)}
        e = VecKArmedBandit(n, k=10, arm_params=arms, seed=1, env_offset=s * n, mean_table=not a.no_mean_table)
        act = torch.randint(0, 10, (n,), dtype=torch.int32, device="cuda")
        ext = None
    envs.append(e)
    acts.append((act, ext))


def once(i):
    e, (act, ext) = envs[i % a.sets], acts[i % a.sets]
    if a.mode == "step":
        e.step(act)
    elif a.mode == "rollout":
        e.rollout(K)
    else:
        e.rollout(K, ext, return_rewards=True)


for i in range(a.sets):
    once(i)
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(a.reps):
    once(i)
t.record()
torch.cuda.synchronize()
us = s.elapsed_time(t) * 1e3 / a.reps
steps = n * (K if a.mode != "step" else 1)
print(f"{a.env} {a.mode} n={n} K={K}: {us:.2f} us/launch  {steps / us * 1e6:.3e} env-steps/s")
