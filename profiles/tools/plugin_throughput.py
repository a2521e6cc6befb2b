"""Throughput of the plugin-facing generator (ChunkRawRecords) on the C1 workload: chunks are consumed
and dropped one by one, as strax does."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from wfsim_b200.resource import Resource
from wfsim_b200.strax_interface import ChunkRawRecords
n_events = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cfg = dict(bench.load_config(), chunk_size=25)           # 25 s of 1 kHz data = 2.5e4 events per chunk
uniq, row = bench.spe_tables()
sim = ChunkRawRecords(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row), seed=1)
inst = bench.workload(n_events, seed=100)
t0 = time.perf_counter(); tl = t0
n_rec = 0
for i, chunk in enumerate(sim(inst)):
    n_rec += len(chunk['raw_records'])
    now = time.perf_counter()
    print(f'chunk {i}: {len(chunk["raw_records"])} records, {len(chunk["truth"])} truth rows, {now - tl:.2f} s', flush=True)
    tl = now
    del chunk
dt = time.perf_counter() - t0
print(f'{n_events} events, {n_rec} records in {dt:.2f} s: {n_events / dt:.0f} events/s, {n_rec * 244 / dt / 1e9:.1f} GB/s of raw_records; '
      f'arenas {len(sim._arena_pool.arenas)}')
