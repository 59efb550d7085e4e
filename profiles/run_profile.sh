#!/bin/bash
# Captures the ncu evidence for one round (run under gpurun, 1 GPU).  Usage: profiles/run_profile.sh <tag> [workload] [extra bench args]
# 1) launch list of OUR kernels with device time (cold-cache, serialised: compare shares, not absolutes)
# 2) one --set full capture of every kernel of one search step (prep, probe select, scan head, tail, scan replay)
TAG=${1:-r01}
WL=${2:-gist1m}
shift; shift
mkdir -p gpurun_out
KREGEX='regex:query_prep|coarse_|probe_select|scan_kernel|tail_|merge_kernel|split_bf16'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 120 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload $WL --nprobe 16 --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launches_${TAG}.log 2>&1
# matching kernels per search: prep, probe_select_tc, scan (head), tail, scan (replay) = 5; 1 recall search + 3 warm-ups precede the step
ncu --set full --clock-control none --import-source on -k 'regex:query_prep_kernel|probe_select_tc_kernel|scan_kernel|tail_kernel' -s 20 -c 5 -f -o gpurun_out/step_${TAG} \
    python bench.py --workload $WL --nprobe 16 --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/step_${TAG}.log 2>&1
ls -la gpurun_out/
