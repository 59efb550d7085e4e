#!/bin/bash
# Captures the ncu evidence for one round (run under gpurun, 1 GPU).  Usage: profiles/run_profile.sh <tag> [workload]
# 1) launch list of OUR kernels with device time (cold-cache, serialised: compare shares, not absolutes)
# 2) one --set full capture of the scan kernel (dominant kernel) with source correlation
TAG=${1:-r01}
WL=${2:-gist1m}
mkdir -p gpurun_out
KREGEX='regex:query_prep_kernel|coarse_|probe_select_kernel|scan_kernel|merge_kernel|rescore'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 60 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_kernel -s 9 -c 1 -f -o gpurun_out/scan_${TAG} \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/scan_${TAG}.log 2>&1
ls -la gpurun_out/
