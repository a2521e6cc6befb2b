#!/usr/bin/env python
"""Top source lines of one launch of an `ncu --page source --csv` dump: instructions, thread utilisation, samples,
barrier share.  usage: sass_top_lines.py <ncu_source.csv> <nvdisasm.sass> <kernel-substring> <launch> <source.cu> [n]"""
import csv, re, sys, collections
ncu_csv, sass, kern, launch, src = sys.argv[1:6]
n_top = int(sys.argv[6]) if len(sys.argv) > 6 else 40
launch = int(launch)
rows = list(csv.reader(open(ncu_csv)))
hs = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
# the launch's own kernel: 'wfs::k_x(...)' -> the mangled '<len>k_x' + 'E' picks the SASS section (k_x vs k_x_small)
kname = rows[hs[launch] - 1][1].split('(')[0].split('::')[-1] if hs[launch] > 0 and rows[hs[launch] - 1][0] == 'Kernel Name' else kern
if kern in kname: kern = '%d%sE' % (len(kname), kname)
print('kernel', kname)
lines = open(sass).read().splitlines()
start = next(i for i, l in enumerate(lines) if '.text.' in l and kern in l and l.strip().startswith('.section'))
cur, idx2 = None, []
for l in lines[start + 1:]:
    if l.strip().startswith('.section') or l.startswith('//----'):
        if idx2: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): idx2.append(cur)
H = rows[hs[launch]]
ie, ss, te, sb = H.index('Instructions Executed'), H.index('# Samples'), H.index('Thread Instructions Executed'), H.index('stall_barrier')
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
k = 0
for r in rows[hs[launch] + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'): break
    if len(r) > ie and (r[0].startswith('0x') or r[0].isdigit()):
        ln = idx2[k] if k < len(idx2) else None
        k += 1
        a = agg[ln]
        a[0] += int(float(r[ie] or 0)); a[1] += int(float(r[ss] or 0)); a[2] += int(float(r[te] or 0)); a[3] += int(float(r[sb] or 0))
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
text = open(src).read().splitlines()
print(f'instructions {ti:.3e} samples {ts}')
for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n_top]:
    t = text[ln[1] - 1].strip()[:90] if ln and ln[0].endswith('.cu') and ln[1] <= len(text) else ''
    print(f'{str(ln):22s} {100*v[0]/ti:5.1f}% instr  {v[2]/max(v[0],1):4.1f} thr  {100*v[1]/ts:5.1f}% smp  {100*v[3]/ts:4.1f}% bar | {t}')
