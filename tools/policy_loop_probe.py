#!/usr/bin/env python
"""us per step of the launch-per-step RL loop on the device: [stand-in policy kernel -> VecGridWorld.step] x 56 in a CUDA
graph (one env set, actions computed from the previous step's observations), at 2^20 envs.  A/B target for step-kernel
variants: here the kernel ahead of a step launch is NOT a step (the device clock is quiescent, the inputs are L2-resident)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import CONFIG2, SPECIAL  # noqa: E402
from rl_envs_forge_b200 import Action, VecGridWorld, _lib  # noqa: E402

n = 1 << 20
COLD = {"auto": None, "0": False, "1": True}[sys.argv[1] if len(sys.argv) > 1 else "auto"]
spec = {(SPECIAL[0], Action(SPECIAL[1])): (SPECIAL[2], SPECIAL[3])}
env = VecGridWorld(n, seed=1, special_transitions=spec, cold_inputs=COLD, **CONFIG2)
lib = _lib.load()
act = torch.zeros(n, dtype=torch.int32, device="cuda")


def policy_then_step(k):
    _lib.check(lib.ef_policy_demo_step(env._obs.data_ptr(), act.data_ptr(), k, n, torch.cuda.current_stream().cuda_stream))
    env.step(act)


policy_then_step(0)
env.enable_device_clock()
stream = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
with torch.cuda.stream(stream):
    with torch.cuda.graph(g, stream=stream):
        for k in range(56):
            policy_then_step(k)
for _ in range(10):
    g.replay()
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(100):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    best = min(best, s.elapsed_time(e) * 1e3 / 5600)
print(f"cold={COLD} policy kernel + step, n={n}: {best:.2f} us/step  {n / best * 1e6:.3e} env-steps/s")
