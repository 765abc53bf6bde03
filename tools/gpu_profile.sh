#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full ncu capture of the dominant kernel of bench.py.
# Usage: tools/gpu_profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-gridworld_step_kernel}
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 5 --no-details --no-cpu-baseline"   # the driver's command, minus the details / CPU legs
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 120 -c 3 -f -o gpurun_out/prof_${TAG} \
    $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
