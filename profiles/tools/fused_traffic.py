#!/usr/bin/env python
"""profiles/r2_fused_ncu.json (read by bench.py for roofline.traffic) from an `ncu --set full` capture of the fused back
end's launches of ONE device batch.  usage: ncu -i x.ncu-rep --page raw --csv | fused_traffic.py <capture name> <out.json>"""
import csv, json, sys
name, out = sys.argv[1:3]
rows = list(csv.reader(sys.stdin))
H, U = rows[0], rows[1]
def col(r, k):
    i = H.index(k); v = float(r[i].replace(',', '')); u = U[i]
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6}.get(u, 1.0)
launches = []
for r in rows[2:]:
    launches.append({'kernel': r[H.index('Kernel Name')].split('(')[0], 'grid': int(float(r[H.index('launch__grid_size')])),
                     'block': int(float(r[H.index('launch__block_size')])), 'us': col(r, 'gpu__time_duration.sum'),
                     'dram_read': col(r, 'dram__bytes_read.sum'), 'dram_write': col(r, 'dram__bytes_write.sum')})
# The lanes' batches interleave in the capture, so one batch is put together from class averages: every size class
# (kernel, threads per CTA, full grid or not) and k_group_records launch once per device batch.
classes = {}
for l in launches:
    key = (l['kernel'], l['block'], l['grid'] >= 148)
    classes.setdefault(key, []).append(l)
launches = []
for key, ls in sorted(classes.items()):
    n = len(ls)
    launches.append({'kernel': key[0], 'block': key[1], 'grid': int(sum(l['grid'] for l in ls) / n), 'captured': n,
                     'us': sum(l['us'] for l in ls) / n, 'dram_read': sum(l['dram_read'] for l in ls) / n,
                     'dram_write': sum(l['dram_write'] for l in ls) / n})
tot = sum(l['dram_read'] + l['dram_write'] for l in launches)
json.dump({'source': f'{name}: ncu --set full --clock-control none of one device batch of the C1 workload (4000 events, 5.4e6 '
                     'photons, 4.98e6 records): dram__bytes_read.sum + dram__bytes_write.sum summed over the size-class '
                     'launches of k_group_analyse / k_group_analyse_small and k_group_records (averages per class over the captured launches)',
           'dram_bytes_per_batch': tot, 'launches': launches}, open(out, 'w'), indent=1)
print(f'{len(launches)} launches, {tot / 1e9:.3f} GB per batch')
