#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh "<tag or empty>[:ENV=VAL]" ...
# (variants are built with `python -m rl_envs_forge_b200._build -DNAME=VALUE --tag=_x`; bench.py needs --allow-variant)
for spec in "$@"; do
  tag="${spec%%:*}"; envs=""; [[ "$spec" == *:* ]] && envs="${spec#*:}"
  [ "$tag" == "default" ] && tag=""
  line=$(env $envs ENVFORGE_LIB=$PWD/rl_envs_forge_b200/csrc/libenvforge_b200$tag.so python bench.py --allow-variant --steps 20000 --warmup 500 --windows 3 --no-details --no-cpu-baseline --e2e-steps 5 --e2e-windows 1 2>&1 | tail -1)
  echo "$spec => $(echo "$line" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("%.3e steps/s  %.2f us/launch  frac %.3f  episodes %d env_steps %d" % (d["value"], d["roofline"]["avg_launch_us"], d["roofline"]["frac"], d["sharding_invariant"]["episodes"], d["sharding_invariant"]["env_steps"]))' 2>&1 | tail -1)"
done
