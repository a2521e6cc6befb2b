"""e2e time of Simulator.simulate into a pinned vs an ordinary (THP) destination, same box, alternating."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from wfsim_b200.resource import Resource
from wfsim_b200.simulator import Simulator
from wfsim_b200.dtypes import raw_record_dtype
cfg = bench.load_config(); uniq, row = bench.spe_tables()
sim = Simulator(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row))
inst = bench.workload(100000, seed=100)
sim.stage(inst); c = sim.run_staged(seed=1)
cap = int(c['n_records_total'] * 1.02) + 1024
buf = bench.host_array(cap, raw_record_dtype())
for rep in range(4):
    for kind in ('pinned', 'pageable'):
        t0 = time.perf_counter()
        if kind == 'pinned':
            sim.simulate(inst, seed=1, cap_records=cap, pinned=True)
        else:
            sim.simulate(inst, seed=1, cap_records=cap, records_out=buf)
        print(rep, kind, f'{1e3 * (time.perf_counter() - t0):.1f} ms', flush=True)
