#!/usr/bin/env python
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split('(')[0]
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('| kernel | launches | total us | share |\n|---|---:|---:|---:|')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |')
print(f'\nTotal {tot:.0f} us over {sum(v[0] for v in agg.values())} launches.')
