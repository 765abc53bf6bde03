#!/usr/bin/env python
"""us per launch of a CUDA-graph chain of VecGridWorld.step at 2^20 envs: ONE env set stepped repeatedly (its working set
stays in L2; every launch waits for a step of the same set) and R sets in rotation (cold L2; the headline's loop).
    python tools/chain_probe.py [--n 1048576] [--sets 1 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import CONFIG2, SPECIAL  # noqa: E402
from rl_envs_forge_b200 import Action, VecGridWorld  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--sets", type=int, nargs="+", default=[1, 8])
ap.add_argument("--replays", type=int, default=300)
ap.add_argument("--cold", default="auto", choices=["auto", "0", "1"], help="EF_STEP_COLD_INPUTS hint: automatic, off, on")
a = ap.parse_args()
COLD = None if a.cold == "auto" else a.cold == "1"
n, A = a.n, 7
spec = {(SPECIAL[0], Action(SPECIAL[1])): (SPECIAL[2], SPECIAL[3])}
acts = [torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda") for _ in range(A)]
for R in a.sets:
    sets = [VecGridWorld(n, seed=1, special_transitions=spec, env_offset=r * n, cold_inputs=COLD, **CONFIG2) for r in range(R)]
    G = R * A
    for i in range(G):
        sets[i % R].step(acts[i % A])
    for e in sets:
        e.enable_device_clock()
    stream = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for i in range(G):
                sets[i % R].step(acts[i % A])
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.replays):
        g.replay()
    t.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(t) * 1e3 / (a.replays * G)
    st = sum(e.episode_stats()["env_steps"] for e in sets)
    print(f"cold={a.cold} {R} set(s) x {n} envs: {us:.2f} us/launch  {n / us * 1e6:.3e} env-steps/s  (env_steps {st})")
    del sets, g
