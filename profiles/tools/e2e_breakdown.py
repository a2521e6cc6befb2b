"""Where does the end-to-end time of Simulator.simulate go?  (run on the GPU box)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from wfsim_b200.resource import Resource, evaluate_instruction_maps
from wfsim_b200.simulator import Simulator

cfg = bench.load_config()
uniq, row = bench.spe_tables()
res = Resource(cfg, spe_ppf=uniq, spe_row=row)
sim = Simulator(cfg, resource=res)
inst = bench.workload(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, seed=100)
sim.stage(inst)
c = sim.run_staged(seed=1)
cap = int(c['n_records_total'] * 1.02) + 1024
for k in range(4):
    t0 = time.perf_counter()
    maps = evaluate_instruction_maps(cfg, res, inst)
    t1 = time.perf_counter()
    out = sim.simulate(inst, seed=1, cap_records=cap, pinned=True, maps=maps)
    t2 = time.perf_counter()
    lc = sim.last_counts
    print(f'maps {1e3*(t1-t0):.1f} ms  simulate {1e3*(t2-t1):.1f} ms  (library ms_total {lc["ms_total"]:.1f}, '
          f'phases {[round(x,1) for x in lc["ms_phase"][:8]]})  records {lc["n_records_total"]}')
