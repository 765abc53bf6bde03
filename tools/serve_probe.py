#!/usr/bin/env python
"""Step-server probe: us per served step (ef_gridworld_serve + the stand-in policy kernel) at 2^20 envs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import CONFIG2, SPECIAL  # noqa: E402
from rl_envs_forge_b200 import Action, VecGridWorld  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
spec = {(SPECIAL[0], Action(SPECIAL[1])): (SPECIAL[2], SPECIAL[3])}
env = VecGridWorld(n, seed=1, special_transitions=spec, **CONFIG2)
env.serve_with_demo_policy(64)
best = 1e9
for _ in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.serve_with_demo_policy(2000)
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / 2000)
print(f"served step, n={n}: {best * 1e6:.2f} us/step  {n / best:.3e} env-steps/s")
