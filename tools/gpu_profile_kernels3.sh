#!/bin/bash
# ncu --set full captures of the final round-1 kernels (one gpurun call).  Usage: tools/gpu_profile_kernels3.sh <tag>
TAG=${1:-r01h}
mkdir -p gpurun_out
prof() {  # name, kernel regex, run_kernel.py args...
  local name=$1 regex=$2; shift 2
  python tools/run_kernel.py "$@" > gpurun_out/plain_${name}_${TAG}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${regex} -s 5 -c 2 -f -o gpurun_out/prof_${name}_${TAG} \
      python tools/run_kernel.py "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1
  echo "$name rc=$? $(cat gpurun_out/plain_${name}_${TAG}.log)"
  ncu -i gpurun_out/prof_${name}_${TAG}.ncu-rep --page raw --csv > gpurun_out/raw_${name}_${TAG}.csv 2>/dev/null
}
prof cartpole_step pendulum_step_kernel --env cartpole --mode step --reps 10 --n 4194304
prof pendulum_step pendulum_step_kernel --env pendulum --mode step --reps 10 --n 4194304
prof cartpole_rollout pendulum_rollout_kernel --env cartpole --mode rollout_ext --reps 6 --n 1048576
prof gridworld_rollout gridworld_rollout_kernel --env gridworld --mode rollout --reps 8 --n 1048576
prof gridworld_step_16m gridworld_step_kernel --env gridworld --mode step --reps 8 --n 16777216 --sets 2
prof bandit_sorted bandit_step_sorted_kernel --env bandit --mode step --reps 6 --n 4194304 --sets 2
prof carrental_step car_rental_step_kernel --env carrental --mode step --reps 8 --n 4194304 --sets 2
prof labyrinth_vision labyrinth_kernel --env labyrinth_vision --mode step --reps 8 --n 1048576 --sets 2
prof gambler_step gambler_step_kernel --env gambler --mode step --reps 10 --n 16777216 --sets 2
