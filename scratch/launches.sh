#!/bin/bash
# quick launch list of one bench run (cold-cache per-kernel durations)
TAG=${1:-x}
shift
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread --clock-control none -k 'regex:query_prep|coarse_|probe_select|scan_kernel|tail_|merge_kernel|split_bf16|head_scan|resolve_|refine_' -s 32 -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --nprobe 16 --steps 1 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/launches_${TAG}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_${TAG}.csv')) if len(r)>10]
hdr=rows[0]; ci={h:i for i,h in enumerate(hdr)}
agg={}
order=[]
for r in rows[1:]:
    key=(r[ci['ID']], r[ci['Kernel Name']][:60])
    if key not in agg: agg[key]={}; order.append(key)
    agg[key][r[ci['Metric Name']]]=r[ci['Metric Value']]
for k in order[-20:]:
    m=agg[k]
    print(k[0].rjust(4), k[1].ljust(60), ' '.join(f"{n.split('.')[0][-22:]}={v}" for n,v in m.items()))
PY
