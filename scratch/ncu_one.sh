#!/bin/bash
# full ncu capture of one kernel inside the timed step: scratch/ncu_one.sh <kernel regex> <tag> [bench args]
K=$1; TAG=$2
shift; shift
export RBQ_CUDA_PROFILER=1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$K" -c 1 -f -o gpurun_out/k_${TAG} \
    python bench.py --nprobe 16 --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/k_${TAG}.log 2>&1
ls -la gpurun_out/k_${TAG}.ncu-rep
