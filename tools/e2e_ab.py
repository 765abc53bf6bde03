#!/usr/bin/env python
"""Time the host-buffer step() of each env family with both host transports (run on the GPU box).
    python tools/e2e_ab.py [--n 1048576] [--steps 40]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rl_envs_forge_b200 import Action, VecCartPole, VecGridWorld, VecKArmedBandit, VecPendulumDisk  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--only", default="")
a = ap.parse_args()
n = a.n
walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
mk = {
    "gridworld": (lambda: VecGridWorld(n, rows=8, cols=8, walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0},
                                       special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)}, p_success=0.8,
                                       episode_length_limit=200, seed=1),
                  lambda r: r.integers(0, 4, n).astype(np.int32), 18),
    "gridworld_u8": (lambda: VecGridWorld(n, rows=8, cols=8, walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0},
                                          special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)}, p_success=0.8,
                                          episode_length_limit=200, seed=1, wire="compact"),
                     lambda r: r.integers(0, 4, n).astype(np.uint8), 9),
    "cartpole": (lambda: VecCartPole(n, seed=1), lambda r: r.uniform(-3, 3, n).astype(np.float32), 26),
    "pendulum": (lambda: VecPendulumDisk(n, seed=1), lambda r: r.uniform(-3, 3, n).astype(np.float32), 18),
    "bandit": (lambda: VecKArmedBandit(n, k=10, seed=1), lambda r: r.integers(0, 10, n).astype(np.int32), 12),
}
for name, (make, gen, bytes_per_env) in mk.items():
    if a.only and name not in a.only.split(","):
        continue
    for transport in ("staged", "zerocopy", "pipelined:1", "pipelined:2", "pipelined:4", "pipelined:8", "pipelined+h2d:4", "pipelined+h2d:8"):
        env = make()
        env.host_transport = transport.split(":")[0].split("+")[0]
        env.host_stage_actions = "+h2d" in transport
        env.host_chunks = int(transport.split(":")[1]) if ":" in transport else 1
        rng = np.random.default_rng(0)
        acts = [torch.from_numpy(gen(rng)).pin_memory() for _ in range(4)]
        for i in range(3):
            env.step(acts[i % 4])
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            for i in range(a.steps):
                out = env.step(acts[i % 4])
            torch.cuda.synchronize()
            best = min(best, (time.perf_counter() - t0) / a.steps)
        print(f"{name:10s} {transport:16s} {best * 1e6:8.1f} us/step  {n / best:.3e} env-steps/s  "
              f"{bytes_per_env * n / best / 1e9:.1f} GB/s over PCIe", flush=True)
        del env
