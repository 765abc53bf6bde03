#!/usr/bin/env python
"""Probe: does backing the pinned host buffers with transparent huge pages change the host-buffer (PCIe) step rate?
In a VM every DMA goes through the IOMMU; 4 KB pages mean an IOTLB entry per 4 KB of the 9 MB a step moves.
    python tools/hugepage_probe.py"""
import ctypes
import mmap
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip() if os.path.exists(
    "/sys/kernel/mm/transparent_hugepage/enabled") else "n/a")
n_bytes = 64 << 20
rt = torch.cuda.cudart()
dev = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")


def bench(host_tensor, label):
    torch.cuda.synchronize()
    for direction in ("d2h", "h2d"):
        best = 0.0
        for _ in range(5):
            t0 = time.perf_counter()
            for _ in range(10):
                if direction == "d2h":
                    host_tensor.copy_(dev, non_blocking=True)
                else:
                    dev.copy_(host_tensor, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, 10 * n_bytes / (time.perf_counter() - t0) / 1e9)
        print(f"{label:28s} {direction}: {best:6.1f} GB/s", flush=True)


bench(torch.empty(n_bytes, dtype=torch.uint8).pin_memory(), "torch pin_memory (4 KB pages)")
# 2 MB-aligned anonymous mapping, MADV_HUGEPAGE, touched, then registered with CUDA
m = mmap.mmap(-1, n_bytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
aligned = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
libc = ctypes.CDLL("libc.so.6", use_errno=True)
rc = libc.madvise(ctypes.c_void_p(aligned), ctypes.c_size_t(n_bytes), 14)   # MADV_HUGEPAGE
print("madvise rc", rc)
arr = np.frombuffer((ctypes.c_uint8 * n_bytes).from_address(aligned), dtype=np.uint8)
arr[:] = 1                                                                   # touch: fault the huge pages in
err = rt.cudaHostRegister(aligned, n_bytes, 0)
print("cudaHostRegister:", err)
t = torch.from_numpy(arr)
print("is_pinned:", t.is_pinned())
bench(t, "THP-backed, cudaHostRegister")
try:
    with open("/proc/meminfo") as f:
        print([ln.strip() for ln in f if "AnonHugePages" in ln or "HugePages_Total" in ln])
except OSError:
    pass
