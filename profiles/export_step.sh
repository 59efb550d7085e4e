#!/bin/bash
# Exports per-kernel summaries of a step capture: profiles/export_step.sh <tag>  (reads gpurun_out/step_<tag>.ncu-rep)
TAG=$1
REP=gpurun_out/step_${TAG}.ncu-rep
ncu -i $REP --page raw --csv 2>/dev/null > /tmp/step_${TAG}_raw.csv
python - "$TAG" <<'PY'
import csv, sys, subprocess, re
tag = sys.argv[1]
rows = list(csv.reader(open(f'/tmp/step_{tag}_raw.csv')))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum']
print('units:', [rows[1][ci[w]] for w in want])
for n, r in enumerate(rows[2:]):
    name = re.sub(r'^void ', '', r[ci['Kernel Name']]).split('(')[0]
    print(n, name[:34].ljust(34), ' | '.join(r[ci[w]][:10] for w in want))
    short = re.sub(r'<.*', '', name)
    if short in ('tail_tc_kernel', 'resolve_head_kernel', 'resolve_lazy_kernel', 'head_scan_kernel', 'probe_select_fast_kernel', 'coarse_gemm_kernel', 'query_prep_fht_kernel', 'refine_kernel'):
        for page in ('raw', 'source'):
            open(f'/tmp/{page}_{short}.csv', 'w').write(subprocess.run(['ncu', '-i', f'gpurun_out/step_{tag}.ncu-rep', '--page', page, '--csv', '--kernel-id', f':::{n + 1}'], capture_output=True, text=True).stdout)
        out = subprocess.run(['python', 'profiles/ncu_summary.py', f'/tmp/raw_{short}.csv', f'/tmp/src_{short}.csv' if False else f'/tmp/source_{short}.csv', '14'], capture_output=True, text=True)
        open(f'profiles/{short}_{tag}_ncu_summary.txt', 'w').write(out.stdout + out.stderr[-300:])
PY
