"""cProfile of Simulator.simulate (1e5 C1 events, reused destination): what is outside the library call?"""
import os, sys, time, cProfile, pstats
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
for _ in range(2):
    sim.simulate(inst, seed=1, cap_records=cap, records_out=buf)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    sim.simulate(inst, seed=1, cap_records=cap, records_out=buf)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
print('library ms_total', sim.last_counts['ms_total'])
