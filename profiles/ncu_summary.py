#!/usr/bin/env python
"""Summarise an ncu capture: key raw metrics + SASS-level opcode mix and top stalled instructions.
Usage: ncu_summary.py <raw.csv> <source.csv> [top_n]   (CSV pages exported with `ncu -i X.ncu-rep --page raw|source --csv`)"""
import collections
import csv
import re
import sys

raw, src = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 16
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.per_cycle_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.avg', 'smsp__cycles_active.avg',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg']
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{h} [{units[i]}] = {vals[i]}")
# tensor pipe / tensor memory activity (tcgen05 kernels): every exported metric that names them
for i, h in enumerate(hdr):
    if any(t in h for t in ('pipe_tensor', 'tmem', 'pipe_uniform', 'lts__t_bytes.sum', 'lts__throughput', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'sm__inst_executed_pipe_tma')):
        print(f"{h} [{units[i]}] = {vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        try:
            if float(vals[i]) >= 0.05:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {float(vals[i]):.3f}")
        except ValueError:
            pass
rows = list(csv.reader(open(src)))
# the page may hold several sections (one per source view): take the first SASS table
heads = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
hdr = rows[heads[0]]
ci = {h: i for i, h in enumerate(hdr)}
end = heads[1] - 1 if len(heads) > 1 else len(rows)
data = [r for r in rows[heads[0] + 1:end] if len(r) == len(hdr)]
tot_s = sum(int(r[ci['# Samples']]) for r in data)
tot_i = sum(int(r[ci['Instructions Executed']]) for r in data)
print(f"SASS lines {len(data)}, warp-instructions {tot_i}, samples {tot_s}")
ops, ops_s = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?PT?\d*\s+)?([A-Z][A-Z0-9_.]+)', r[ci['Source']])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(r[ci['Instructions Executed']])
    ops_s[op] += int(r[ci['# Samples']])
print("opcode      inst%  samples%")
for op, c in ops.most_common(18):
    print(f"  {op:10s} {100 * c / tot_i:6.2f} {100 * ops_s[op] / tot_s:6.2f}")
print("top stalled instructions:")
for r in sorted(data, key=lambda r: -int(r[ci['# Samples']]))[:topn]:
    st = {k[6:]: int(r[ci[k]]) for k in hdr if k.startswith('stall_') and '(' not in k and int(r[ci[k]]) > 0}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  [{data.index(r):5d}] {r[ci['Source']].strip()[:64]:64s} samples {int(r[ci['# Samples']]):6d} ({100 * int(r[ci['# Samples']]) / tot_s:4.1f}%) exec {r[ci['Instructions Executed']]:>9s} {main}")
