"""simulate() into a fresh output array each call (what the plugin does), 1e5 C1 events."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from wfsim_b200.resource import Resource
from wfsim_b200.simulator import Simulator
cfg = bench.load_config(); uniq, row = bench.spe_tables()
sim = Simulator(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row))
inst = bench.workload(int(sys.argv[1]) if len(sys.argv) > 1 else 100000, seed=100)
for rep in range(4):
    t0 = time.perf_counter()
    out = sim.simulate(inst, seed=1, cap_records=int(1.3e8) if len(inst) > 150000 else None)
    dt = time.perf_counter() - t0
    print(rep, f'{1e3 * dt:.0f} ms, {len(out["raw_records"])} records', flush=True)
    del out
