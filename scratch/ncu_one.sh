#!/bin/bash
# full ncu capture of one kernel (regex $1), tag $2, skip $3 matching launches
K=$1; TAG=$2; SKIP=${3:-4}
shift; shift; shift
ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c 1 -f -o gpurun_out/k_${TAG} \
    python bench.py --nprobe 16 --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/k_${TAG}.log 2>&1
ls -la gpurun_out/k_${TAG}.ncu-rep
