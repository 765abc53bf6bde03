#!/usr/bin/env python
"""Calibration: how fast can ANY kernel move the step kernel's traffic (44 MB per launch, rotating buffers > L2,
CUDA-graph replay)?  Uses torch's vectorised elementwise add as the yardstick.  Prints us/launch and GB/s."""
import sys
import torch

def run(n_bytes_each, sets=8, reps=2000):
    n = n_bytes_each // 4
    xs = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(sets)]
    ys = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(sets)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(sets):
            torch.add(xs[i], 1, out=ys[i])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(sets):
                torch.add(xs[i], 1, out=ys[i])
    torch.cuda.synchronize()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps // sets):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps // sets * sets)
    print(f"elementwise add  {2 * n_bytes_each / 1e6:8.1f} MB/launch  {us:8.2f} us/launch  {2 * n_bytes_each / us / 1e3:8.1f} GB/s")

for mb in (22, 44, 88, 176, 704):
    run(mb * 1000 * 1000 // 2 * 2 // 2 if False else mb * 1_000_000 // 2 * 1, reps=2000 if mb < 200 else 400)
