import sys, numpy as np
sys.path.insert(0, '/root/repo')
import bench
from wfsim_b200.resource import Resource
from wfsim_b200.simulator import Simulator
cfg = bench.load_config(); uniq, row = bench.spe_tables()
res = Resource(cfg, spe_ppf=uniq, spe_row=row)
sim = Simulator(cfg, resource=res)
inst = bench.workload(20000, seed=100)
sim.stage(inst)
for k in range(3):
    c = sim.run_staged(seed=1)
    print('batches', c['n_batches'], 'seg photon batches', c['ms_phase'][10], 'phases', [round(x, 2) for x in c['ms_phase'][:8]], 'total', round(c['ms_total'], 2))
