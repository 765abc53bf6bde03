import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from rl_envs_forge_b200 import VecGridWorldLayouts, VecGridWorld, Action
walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
lay = dict(start_state=(0,0), walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0}, special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
a = torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda")
SETS = 4 if n <= 1 << 22 else 2
def run(envs, tag):
    # CUDA graph replay (device step clock): a Python-driven step() call costs more than the kernel at 2^20 envs
    for e in envs: e.enable_device_clock(); e.step(a)
    st = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for i in range(20): envs[i % len(envs)].step(a)
    g.replay(); torch.cuda.synchronize(); s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s0.record()
    for _ in range(5): g.replay()
    s1.record(); torch.cuda.synchronize(); print(tag, sys.argv[1:], s0.elapsed_time(s1) / 100 * 1e3, "us/launch", flush=True)
run([VecGridWorldLayouts(n, [lay] * 8, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n) for s in range(SETS)], "8 layouts")
run([VecGridWorldLayouts(n, [lay] * 1, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n) for s in range(SETS)], "1 layout via layouts API")
if len(sys.argv) <= 2:
  run([VecGridWorld(n, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, env_offset=s * n, state_layout="split", **lay) for s in range(SETS)], "single split")
