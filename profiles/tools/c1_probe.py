"""Device-resident C1 run at a given size (one lane by default): per-phase times of the library.
usage: c1_probe.py [events] [steps]   (env WFS_LANES, WFS_FUSED, ...)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault('WFS_LANES', '1')
import numpy as np
from bench import load_config, spe_tables, workload
from wfsim_b200.resource import Resource
from wfsim_b200.simulator import Simulator

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg = load_config()
uniq, row = spe_tables()
sim = Simulator(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row), device=0)
sim.stage(workload(n, seed=100))
for k in range(steps):
    c = sim.run_staged(seed=1)
    names = ['front', 'sort', 'win', 'digi', 'zle', 'rsort', 'pack', 'host']
    print(f"step {k}: {c['ms_total']:.1f} ms, batches {c['n_batches']} fused {c['n_fused_batches']}, photons {c['n_photons']:.3g}, "
          f"records {c['n_records_total']:.3g}; " + ' '.join(f'{a} {b:.1f}' for a, b in zip(names, c['ms_phase'][:8])), flush=True)
sim.close()
