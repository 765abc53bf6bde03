#!/bin/bash
# ncu --set full captures of the secondary kernels (one gpurun call).  Usage: tools/gpu_profile_kernels.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
prof() {  # name, kernel regex, run_kernel.py args...
  local name=$1 regex=$2; shift 2
  python tools/run_kernel.py "$@" > gpurun_out/plain_${name}_${TAG}.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:${regex} -s 5 -c 2 -f -o gpurun_out/prof_${name}_${TAG} \
      python tools/run_kernel.py "$@" > gpurun_out/ncu_${name}_${TAG}.log 2>&1
  echo "$name rc=$? $(cat gpurun_out/plain_${name}_${TAG}.log)"
}
prof cartpole_step pendulum_step_kernel --env cartpole --mode step --reps 10
prof pendulum_step pendulum_step_kernel --env pendulum --mode step --reps 10
prof cartpole_rollout pendulum_rollout_kernel --env cartpole --mode rollout_ext --reps 6 --n 1048576
prof gridworld_rollout gridworld_rollout_kernel --env gridworld --mode rollout --reps 8 --n 1048576
prof gridworld_step_16m gridworld_step_kernel --env gridworld --mode step --reps 8 --n 16777216 --sets 2
prof bandit_step bandit_step_kernel --env bandit --mode step --reps 6 --n 4194304 --sets 2
