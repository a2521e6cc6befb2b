#!/usr/bin/env python
"""Like sass_by_line.py, but sums instructions / stall samples over source-line ranges.
usage: sass_by_range.py <ncu_source.csv> <nvdisasm.sass> <kernel-substring> [launch] name:lo-hi [name:lo-hi ...]"""
import csv, re, sys, collections
ncu_csv, sass, kern = sys.argv[1:4]
launch = 0
if ':' not in sys.argv[4]: launch = int(sys.argv[4]); del sys.argv[4]      # optional: which launch of the report (default the first)
ranges = []
for a in sys.argv[4:]:
    n, r = a.split(':'); lo, hi = r.split('-'); ranges.append((n, int(lo), int(hi)))
lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if '.text.' in l and kern in l and l.strip().startswith('.section'))
cur_line, idx2line, idx = None, [], 0
for l in lines[start + 1:]:
    if l.strip().startswith('.section') or l.startswith('//----'):
        if idx: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        idx2line.append(cur_line); idx += 1
rows = list(csv.reader(open(ncu_csv)))
h = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][launch]
H = rows[h]
ie, ss = H.index('Instructions Executed'), H.index('# Samples')
body = []
for r in rows[h + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'): break
    if len(r) > ie and (r[0].startswith('0x') or r[0].isdigit()): body.append(r)
agg = collections.defaultdict(lambda: [0, 0])
for k, r in enumerate(body):
    ln = idx2line[k] if k < len(idx2line) else None
    name = 'other'
    if ln and ln[0].endswith('.cu'):
        for n, lo, hi in ranges:
            if lo <= ln[1] <= hi: name = n; break
    elif ln:
        name = 'hdr:' + ln[0]
    agg[name][0] += int(float(r[ie] or 0)); agg[name][1] += int(float(r[ss] or 0))
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f'instructions {ti:.3e} samples {ts}')
for n, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{n:28s} {100*v[0]/ti:5.1f}% instr {100*v[1]/max(ts,1):5.1f}% stall-samples')
