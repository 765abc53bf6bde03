#!/usr/bin/env python
"""Timeline of the GridWorld step kernels inside the bench's graph-replay loop (tools only).

Needs a library built with the trace hooks:  python -m rl_envs_forge_b200._build -DEF_TRACE=1 --tag=_trace
    ENVFORGE_LIB=$PWD/rl_envs_forge_b200/csrc/libenvforge_b200_trace.so python tools/step_trace.py [--n 1048576]
Thread 0 of every CTA stores %globaltimer and clock64 at: 0 entry, 1 released by the previous launch (griddepcontrol.wait),
2 loads issued, 3+2s inputs of stage s landed, 4+2s stage s stepped and stored, 13 before the CTA epilogue, 14 done.
Prints, per launch of the last rounds, the distribution over CTAs of each point relative to the launch's first release."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import CONFIG2, SPECIAL  # noqa: E402
from rl_envs_forge_b200 import Action, VecGridWorld, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--sets", type=int, default=8, help="env sets in rotation (1 = one set stepped repeatedly: L2-resident chain)")
ap.add_argument("--policy", action="store_true", help="one env set, the stand-in policy kernel ahead of every step launch")
a = ap.parse_args()
if a.policy:
    a.sets = 1
n, R, A = a.n, a.sets, 7
lib = _lib.load()
lib.ef_debug_set_trace.argtypes = [C.c_void_p]
buf = torch.zeros(8 * 4 * 1024 * 32, dtype=torch.int64, device="cuda")
assert lib.ef_debug_set_trace(buf.data_ptr()) == 0
spec = {(SPECIAL[0], Action(SPECIAL[1])): (SPECIAL[2], SPECIAL[3])}
sets = [VecGridWorld(n, seed=1, special_transitions=spec, env_offset=r * n, **CONFIG2) for r in range(R)]
for e in sets:
    e.enable_device_clock()
acts = [torch.randint(0, 4, (n,), dtype=torch.int32, device="cuda") for _ in range(A)]
G = R * A
for i in range(G):
    sets[i % R].step(acts[i % A])
torch.cuda.synchronize()
stream = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
act_p = torch.zeros(n, dtype=torch.int32, device="cuda")


def one(i):
    if a.policy:
        _lib.check(lib.ef_policy_demo_step(sets[0]._obs.data_ptr(), act_p.data_ptr(), i, n, torch.cuda.current_stream().cuda_stream))
        sets[0].step(act_p)
    else:
        sets[i % R].step(acts[i % A])


with torch.cuda.stream(stream):
    with torch.cuda.graph(g, stream=stream):
        for i in range(G):
            one(i)
torch.cuda.synchronize()
for _ in range(a.rounds):
    g.replay()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    g.replay()
ev1.record()
torch.cuda.synchronize()
print(f"graph replay: {ev0.elapsed_time(ev1) * 1e3 / (20 * G):.2f} us per launch (with trace stores)")
tr = buf.cpu().numpy().reshape(8, 4, 1024, 16, 2).astype(np.int64)
launches = []
for s in range(8):
    for k in range(4):
        rec = tr[s, k]
        used = rec[:, 0, 0] != 0
        if used.any():
            launches.append((rec[used, 1, 0].min(), s, k, rec[used]))
launches.sort(key=lambda x: x[0])
names = {0: "entry", 1: "released", 2: "loads issued", 12: "pre-wait work done", 13: "pre-epilogue", 14: "done"}
prev_rel = launches[-6][0] if len(launches) > 5 else None
for rel0, s, k, rec in launches[-5:]:
    ncta = rec.shape[0]
    head = f"set {s} slot {k}: {ncta} CTAs"
    if prev_rel is not None:
        head += f"   period since previous launch's first release: {(rel0 - prev_rel) / 1e3:.2f} us"
    prev_rel = rel0
    print(head)
    for p in range(16):
        col = rec[:, p, 0]
        ok = col != 0
        if not ok.any():
            continue
        d = (col[ok] - rel0) / 1e3
        label = names.get(p, f"stage {(p - 3) // 2} " + ("landed" if (p - 3) % 2 == 0 else "stepped"))
        print(f"   {p:2d} {label:16s} min {d.min():7.2f}  p50 {np.median(d):7.2f}  max {d.max():7.2f} us   ({ok.sum()} CTAs)")
    # per-CTA durations in SM cycles (clock64 is per SM: only differences inside a CTA mean anything)
    c = rec[:, :, 1]
    def span(p, q):
        ok = (c[:, p] != 0) & (c[:, q] != 0)
        return np.median(c[ok, q] - c[ok, p]) if ok.any() else float("nan")
    print(f"   cycles (median per CTA): entry->released {span(0, 1):.0f}, released->issued {span(1, 2):.0f}, "
          f"issued->stage0 landed {span(2, 3):.0f}, stage0 step {span(3, 4):.0f}, pre-epilogue->done {span(13, 14):.0f}, "
          f"released->done {span(1, 14):.0f}")
