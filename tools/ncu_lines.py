#!/usr/bin/env python
"""Per-source-line executed-instruction table from an ncu report captured with --import-source on.
    python tools/ncu_lines.py report.ncu-rep [top_n] [kernel-substring]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = None
hdr = None
agg = {}
seen_fn = set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and (want in (cur_fn or "")):
        i = hdr.index("Instructions Executed")
        try:
            n = int(r[i])
        except ValueError:
            continue
        key = (cur_fn, cur_file, int(r[0]), r[1].strip()[:100])
        agg[key] = max(agg.get(key, 0), n)          # the same kernel can appear once per captured launch
fns = sorted({k[0] for k in agg})
for fn in fns[:1]:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot = sum(v for _, v in items)
    print(fn, "warp-instr:", tot)
    for (f, fl, ln, src), v in sorted(items, key=lambda kv: -kv[1])[:top]:
        print(f"{100 * v / tot:5.1f}% {v:10d} {fl}:{ln}: {src}")
