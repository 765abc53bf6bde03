#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` dump: warp-stall samples by reason and the instructions that collect them.
    ncu -i prof.ncu-rep --page source --csv --kernel-id :::1 > src.csv; python tools/ncu_stalls.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
idx = {k: i for i, k in enumerate(hdr)}
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]


def num(r, k):
    try:
        return int(r[idx[k]])
    except ValueError:
        return 0


total = sum(num(r, "# Samples") for r in data)
print(rows[0][1][:120] if rows[0] else "")
print("total samples", total, " warp instructions", sum(num(r, "Instructions Executed") for r in data), " SASS lines", len(data))
for k, v in sorted(((k, sum(num(r, k) for r in data)) for k in stalls), key=lambda x: -x[1])[:10]:
    print(f"  {k:26s} {v:7d} {100 * v / max(total, 1):5.1f}%")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:top_n]:
    dom = max(stalls, key=lambda k: num(r, k))
    print(f"{num(r, '# Samples'):6d} {num(r, 'Instructions Executed'):8d} {dom:18s} {r[idx['Source']].strip()[:100]}")
