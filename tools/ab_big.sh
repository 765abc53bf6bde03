#!/bin/bash
# A/B timing of library variants on the large-batch GridWorld step (and any run_kernel.py config).
# Usage: tools/ab_big.sh "<run_kernel.py args>" tag1 tag2 ...     (tag "default" = the in-tree build)
ARGS="$1"; shift
for tag in "$@"; do
  t="$tag"; [ "$tag" == "default" ] && t=""
  echo "$tag => $(ENVFORGE_LIB=$PWD/rl_envs_forge_b200/csrc/libenvforge_b200$t.so python tools/run_kernel.py $ARGS 2>&1 | tail -1)"
done
