#!/bin/bash
# Executed thread-instructions per launch of the issue-bound kernels at their bench sizes (one gpurun call; light ncu pass:
# three metrics, no --set full).  Output: gpurun_out/instr_<tag>.csv, one line per kernel launch profiled.
#   tools/gpu_instr_counts.sh <tag>   ->   python tools/instr_summary.py gpurun_out/instr_<tag>.csv > profiles/kernel_instr.json
TAG=${1:-r02}
OUT=gpurun_out/instr_${TAG}.csv
: > $OUT
count() {  # label, kernel regex, envs per launch, steps per launch, run_kernel.py args...
  local label=$1 regex=$2 n=$3 k=$4; shift 4
  python tools/run_kernel.py "$@" > gpurun_out/plain_${label}_${TAG}.log 2>&1 &&
  ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none \
      -k regex:${regex} -s 3 -c 2 --csv python tools/run_kernel.py "$@" 2> gpurun_out/ncu_${label}_${TAG}.err |
      grep -E '^"[0-9]+"' | sed "s/^/\"${label}\",\"${n}\",\"${k}\",/" >> $OUT
  echo "$label rc=$? $(tail -1 gpurun_out/plain_${label}_${TAG}.log)"
}
count gridworld_rollout gridworld_rollout_kernel 1048576 64 --env gridworld --mode rollout --reps 6 --n 1048576
count cartpole_rollout_ext pendulum_rollout_kernel 4194304 64 --env cartpole --mode rollout_ext --reps 5 --n 4194304 --sets 2
count cartpole_rollout_policy pendulum_rollout_kernel 4194304 64 --env cartpole --mode rollout --reps 5 --n 4194304 --sets 2
count pendulum_rollout_policy pendulum_rollout_kernel 4194304 64 --env pendulum --mode rollout --reps 5 --n 4194304 --sets 2
count bandit_sorted bandit_step_sorted_kernel 16777216 1 --env bandit --mode step --reps 5 --n 16777216 --sets 2
count carrental_step car_rental_step_kernel 4194304 1 --env carrental --mode step --reps 5 --n 4194304 --sets 2
count gambler_rollout gambler_rollout_kernel 4194304 64 --env gambler --mode rollout --reps 5 --n 4194304 --sets 2
count gridworld_step gridworld_step_kernel 1048576 1 --env gridworld --mode step --reps 8 --n 1048576 --sets 4
count cartpole_step pendulum_step_kernel 4194304 1 --env cartpole --mode step --reps 8 --n 4194304 --sets 2
