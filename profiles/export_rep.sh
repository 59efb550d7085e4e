#!/bin/bash
# Per-kernel summaries of any ncu capture: profiles/export_rep.sh <file.ncu-rep> <tag> [outdir=profiles]
# (raw + source pages -> ncu_summary.py; one <kernel>_<tag>_ncu_summary.txt per distinct kernel, first launch of each)
REP=$1; TAG=$2; OUT=${3:-profiles}
ncu -i $REP --page raw --csv 2>/dev/null > /tmp/rep_${TAG}_raw.csv
python - "$REP" "$TAG" "$OUT" <<'PY'
import csv, sys, subprocess, re
rep, tag, outdir = sys.argv[1:4]
rows = list(csv.reader(open(f'/tmp/rep_{tag}_raw.csv')))
hdr = rows[0]; ci = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum']
print('units:', [rows[1][ci[w]] for w in want])
seen = set()
for n, r in enumerate(rows[2:]):
    name = re.sub(r'^void ', '', r[ci['Kernel Name']]).split('(')[0]
    print(n, name[:44].ljust(44), ' | '.join(r[ci[w]][:10] for w in want))
    short = re.sub(r'^rbq::', '', name)
    short = re.sub(r'[<>, ]+', '_', short).strip('_')
    if short in seen:
        continue
    seen.add(short)
    for page in ('raw', 'source'):
        open(f'/tmp/{page}_{short}.csv', 'w').write(subprocess.run(['ncu', '-i', rep, '--page', page, '--csv', '--kernel-id', f':::{n + 1}'], capture_output=True, text=True).stdout)
    out = subprocess.run(['python', 'profiles/ncu_summary.py', f'/tmp/raw_{short}.csv', f'/tmp/source_{short}.csv', '14'], capture_output=True, text=True)
    open(f'{outdir}/{short}_{tag}_ncu_summary.txt', 'w').write(out.stdout + out.stderr[-300:])
PY
