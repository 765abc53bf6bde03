import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rl_envs_forge_b200 import VecGridWorld, Action
walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
lay = dict(start_state=(0, 0), walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0}, special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
for n in (1 << 20, 1 << 22):
    envs = [VecGridWorld(n, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1 + s * n, rng="pcg64", **lay) for s in range(2)]
    a = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
    for e in envs: e.step(a)
    torch.cuda.synchronize(); s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s0.record()
    for i in range(20): envs[i % 2].step(a)
    s1.record(); torch.cuda.synchronize()
    us = s0.elapsed_time(s1) / 20 * 1e3
    print(f"pcg64 step n={n}: {us:.1f} us/launch  {n / us * 1e6:.3e} env-steps/s  {(34 + 48) * n / us / 1e3:.0f} GB/s moved", flush=True)
