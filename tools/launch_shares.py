#!/usr/bin/env python
"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/gpu_profile.sh).
    python tools/launch_shares.py gpurun_out/launches_<tag>.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14]
hdr = rows[0]
k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    name = r[k].split("(")[0][:110]
    tot[name] += float(r[v]) / 1e3
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total:.1f} us over {sum(cnt.values())} launches")
for name in sorted(tot, key=lambda n: -tot[n]):
    print(f"{100 * tot[name] / total:5.1f}%  {cnt[name]:4d} launches  {tot[name] / cnt[name]:9.2f} us avg  {name}")
