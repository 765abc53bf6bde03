#!/usr/bin/env python
"""e2e probe: host-buffer GridWorld steps with S env sets in flight (step_async / wait), per transport and wire format.
    python tools/e2e_async_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import CONFIG2, SPECIAL  # noqa: E402
from rl_envs_forge_b200 import Action, VecGridWorld  # noqa: E402

n = 1 << 20
spec = {(SPECIAL[0], Action(SPECIAL[1])): (SPECIAL[2], SPECIAL[3])}
for wire in ("compact", "canonical"):
    adt = np.uint8 if wire == "compact" else np.int32
    h = [torch.from_numpy(np.random.default_rng(i).integers(0, 4, n).astype(adt)).pin_memory() for i in range(4)]
    for transport in ("zerocopy", "staged"):
        for S in (1, 2, 3, 4):
            envs = [VecGridWorld(n, seed=1, special_transitions=spec, env_offset=k * n, wire=wire, **CONFIG2) for k in range(S)]
            for e in envs:
                e.host_transport = transport
            for i in range(32):
                envs[i % S].step_async(h[i % 4])
                envs[i % S].wait()
            best = 0.0
            for _ in range(8):
                torch.cuda.synchronize()
                K = 120
                t0 = time.perf_counter()
                for i in range(min(S - 1, K)):
                    envs[i % S].step_async(h[i % 4])
                for i in range(K):
                    j = i + S - 1
                    if j < K:
                        envs[j % S].step_async(h[j % 4])
                    out = envs[i % S].wait()
                    _ = float(out[1][0])
                torch.cuda.synchronize()
                best = max(best, K * n / (time.perf_counter() - t0))
            print(f"{wire:9s} {transport:8s} sets in flight {S}: best window {best:.3e} env-steps/s "
                  f"({(9 if wire == 'compact' else 18) * best / 1e9:.1f} GB/s over PCIe)", flush=True)
            del envs
