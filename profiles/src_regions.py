#!/usr/bin/env python
"""Region breakdown of an ncu source page (CSV): contiguous SASS ranges with equal execution counts, with their share of
executed instructions and stall samples.  Usage: src_regions.py <source.csv> [section_index]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
heads = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
hi = heads[want]
end = heads[want + 1] - 1 if want + 1 < len(heads) else len(rows)
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:end] if len(r) == len(hdr)]
tot_s = sum(int(r[ci['# Samples']]) for r in data)
tot_i = sum(int(r[ci['Instructions Executed']]) for r in data)
print(f"sections {len(heads)}; SASS lines {len(data)}, samples {tot_s}, warp-instructions {tot_i}")
stall_cols = [h for h in hdr if h.startswith('stall_') and '(' not in h]
agg = collections.Counter()
for r in data:
    for h in stall_cols:
        agg[h] += int(r[ci[h]] or 0)
print({k[6:]: v for k, v in agg.most_common(10)})
regions, cur = [], None
for idx, r in enumerate(data):
    e, s = int(r[ci['Instructions Executed']]), int(r[ci['# Samples']])
    if cur and abs(cur['e'] - e) <= max(0.05 * cur['e'], 50):
        cur['n'] += 1; cur['s'] += s; cur['i'] += e; cur['end'] = idx
    else:
        cur = {'start': idx, 'end': idx, 'e': e, 'n': 1, 's': s, 'i': e}
        regions.append(cur)
for g in regions:
    if g['s'] > tot_s * 0.01 or g['i'] > tot_i * 0.01:
        print(f"lines {g['start']:5d}-{g['end']:5d} exec/line {g['e']:9d} lines {g['n']:4d} inst {100 * g['i'] / tot_i:5.1f}% "
              f"samples {100 * g['s'] / tot_s:5.1f}%  first: {data[g['start']][ci['Source']].strip()[:60]}")
if len(sys.argv) > 3:
    for idx, r in enumerate(data):
        s = int(r[ci['# Samples']])
        if s >= tot_s * float(sys.argv[3]):
            st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]] or 0) > 0}
            main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(idx, r[ci['Source']].strip()[:70].ljust(70), s, r[ci['Instructions Executed']], main)
