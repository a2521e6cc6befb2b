#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS view) dump by CUDA source line, using the line
table of `nvdisasm -g -c <cubin>`.  usage: sass_by_line.py <ncu_source.csv> <nvdisasm.sass> <kernel-substring>"""
import csv, re, sys, collections
ncu_csv, sass, kern = sys.argv[1:4]
# 1) instruction index -> source line, from nvdisasm
lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if '.text.' in l and kern in l and l.strip().startswith('.section'))
cur_line, idx2line, idx = None, [], 0
for l in lines[start + 1:]:
    if l.strip().startswith('.section') or l.startswith('//----'):
        if idx: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        idx2line.append(cur_line); idx += 1
# 2) per-instruction metrics from ncu
rows = list(csv.reader(open(ncu_csv)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
H = rows[h]
ie, ss = H.index('Instructions Executed'), H.index('# Samples')
agg = collections.defaultdict(lambda: [0, 0])
body = []
for r in rows[h + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        break      # next launch in the report
    if len(r) > ie and (r[0].startswith('0x') or r[0].isdigit()):
        body.append(r)
for k, r in enumerate(body):
    ln = idx2line[k] if k < len(idx2line) else None
    try:
        agg[ln][0] += int(float(r[ie] or 0)); agg[ln][1] += int(float(r[ss] or 0))
    except ValueError:
        pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f'instructions {tot_i:.3e}  samples {tot_s}  (sass instrs {len(body)} / lineinfo {len(idx2line)})')
src_cache = {}
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[5]) if len(sys.argv) > 5 else 0])[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]:
    text = ''
    if ln:
        f = ln[0]
        if f not in src_cache:
            import glob
            c = glob.glob(f'/root/repo/wfsim_b200/csrc/{f}')
            src_cache[f] = open(c[0]).read().splitlines() if c else []
        if 0 < ln[1] <= len(src_cache[f]): text = src_cache[f][ln[1] - 1].strip()[:90]
    print(f'{100*v[0]/tot_i:5.1f}% instr {100*v[1]/max(tot_s,1):5.1f}% stall-samples  {ln}  {text}')
