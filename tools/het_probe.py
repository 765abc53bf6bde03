import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from rl_envs_forge_b200 import VecGridWorldLayouts, VecGridWorld, Action
walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
lay = dict(start_state=(0,0), walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0}, special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
n = 1 << 24
a = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
def run(envs, tag):
    for e in envs: e.step(a)
    torch.cuda.synchronize(); s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s0.record()
    for i in range(20): envs[i % len(envs)].step(a)
    s1.record(); torch.cuda.synchronize(); print(tag, sys.argv[1:], s0.elapsed_time(s1) / 20 * 1e3, "us/launch", flush=True)
run([VecGridWorldLayouts(n, [lay] * 8, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n) for s in range(2)], "8 layouts")
run([VecGridWorldLayouts(n, [lay] * 1, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n) for s in range(2)], "1 layout via layouts API")
run([VecGridWorld(n, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n, state_layout="split", **lay) for s in range(2)], "single split")
