#!/usr/bin/env python
"""gpurun_out/instr_<tag>.csv (tools/gpu_instr_counts.sh) -> JSON: executed thread-instructions per env-step per kernel.
    python tools/instr_summary.py gpurun_out/instr_r02.csv > profiles/kernel_instr.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
acc = {}
for r in rows:
    label, n, k = r[0], int(r[1]), int(r[2])
    name, metric, unit, value = r[7], r[15], r[16], float(r[17].replace(",", ""))   # ncu --csv: Kernel Name, Metric Name, Unit, Value
    if metric == "gpu__time_duration.sum":   # normalise to microseconds
        value *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(unit, 1.0)
    d = acc.setdefault(label, {"kernel": name.split("(")[0], "envs": n, "steps_per_launch": k, "launches": {}})
    d["launches"].setdefault(r[3], {})[metric] = value
out = {"_note": "executed instructions per launch from `ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,"
                "gpu__time_duration.sum --clock-control none` (tools/gpu_instr_counts.sh), mean of the profiled launches; "
                "per env-step = per launch / (envs * steps_per_launch)", "_source": sys.argv[1]}
for label, d in acc.items():
    ls = list(d["launches"].values())
    ti = sum(x["smsp__thread_inst_executed.sum"] for x in ls) / len(ls)
    wi = sum(x["smsp__inst_executed.sum"] for x in ls) / len(ls)
    us = sum(x["gpu__time_duration.sum"] for x in ls) / len(ls)
    steps = d["envs"] * d["steps_per_launch"]
    out[label] = {"kernel": d["kernel"], "envs": d["envs"], "steps_per_launch": d["steps_per_launch"],
                  "thread_instr_per_env_step": ti / steps, "warp_instr_per_launch": wi,
                  "active_lanes_per_warp_instr": ti / wi, "ncu_duration_us": us}
print(json.dumps(out, indent=1))
