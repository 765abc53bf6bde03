import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rl_envs_forge_b200 import VecGridWorld, Action
walls = {(1, 1), (1, 2), (2, 5), (3, 5), (4, 5), (6, 2), (6, 3), (5, 6)}
lay = dict(start_state=(0, 0), walls=walls, terminal_states={(7, 7): 1.0, (0, 7): -1.0}, special_transitions={((3, 3), Action.UP): ((6, 6), 0.5)})
n = 1 << 20
for transport in ("staged", "zerocopy", "pipelined"):
    e = VecGridWorld(n, rows=8, cols=8, p_success=0.8, episode_length_limit=200, seed=1, **lay)
    e.host_transport = transport
    e.host_stage_actions = True
    acts = [torch.from_numpy(np.random.default_rng(i).integers(0, 4, n).astype(np.int32)).pin_memory() for i in range(4)]
    for i in range(5): e.step(acts[i % 4])
    ts = []
    for i in range(200):
        t0 = time.perf_counter(); e.step(acts[i % 4]); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e6
    print(transport, "us/step: min %.0f p10 %.0f median %.0f p90 %.0f max %.0f mean %.0f" % (ts.min(), np.percentile(ts, 10), np.median(ts), np.percentile(ts, 90), ts.max(), ts.mean()), flush=True)
# raw copies for reference
h = torch.empty(14680064, dtype=torch.uint8).pin_memory(); d = torch.empty(14680064, dtype=torch.uint8, device="cuda")
ha = torch.empty(4194304, dtype=torch.uint8).pin_memory(); da = torch.empty(4194304, dtype=torch.uint8, device="cuda")
for name, dst, src in (("D2H 14.7MB", h, d), ("H2D 4.2MB", da, ha)):
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
    print(name, "%.0f us  %.1f GB/s" % (dt * 1e6, src.numel() / dt / 1e9))
