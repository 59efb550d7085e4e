#!/bin/bash
# Captures the ncu evidence for one round (run under gpurun, 1 GPU).  Usage: profiles/run_profile.sh <tag> [workload] [extra bench args]
# bench.py brackets its timed device steps with cudaProfilerStart/Stop when RBQ_CUDA_PROFILER=1, so
# `--profile-from-start off` captures exactly the kernels of the timed region.
# 1) launch list of OUR kernels with device time (cold-cache, serialised: compare shares, not absolutes), 2 steps
# 2) one --set full capture of every kernel of one search step
TAG=${1:-r01}
WL=${2:-gist1m}
shift; shift
mkdir -p gpurun_out
export RBQ_CUDA_PROFILER=1
KREGEX='regex:query_prep|coarse_|probe_|sample_threshold|scan_kernel|tail_|merge_kernel|split_bf16|head_scan|head_compact|resolve_|refine_|fallback_|xr_|empty_rank|fill_u'
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k "$KREGEX" -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --workload $WL --nprobe 16 --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k "$KREGEX" -c 40 -f -o gpurun_out/step_${TAG} \
    python bench.py --workload $WL --nprobe 16 --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/step_${TAG}.log 2>&1
ls -la gpurun_out/
