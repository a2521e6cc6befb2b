import sys, time, os
sys.path.insert(0,'/root/repo')
import numpy as np
from tests.conftest import load_c0_config, GOLDEN
from tests.golden.synth_instructions import c0_like
from wfsim_b200.simulator import Simulator
from wfsim_b200.resource import Resource
cfg = load_c0_config()
z = np.load(os.path.join(GOLDEN,'c0_tables.npz'))
res = Resource(cfg, spe_ppf=z['spe_unique'], spe_row=z['spe_row'][:494])
sim = Simulator(cfg, resource=res)
inst = c0_like(int(sys.argv[1]) if len(sys.argv)>1 else 10, seed=1)
t0=time.time(); out = sim.simulate(inst, seed=5); t1=time.time()
print({k:(len(v) if hasattr(v,'__len__') else v) for k,v in out.items() if k!='_pinned'})
print(sim.last_counts, t1-t0)
tr = out['truth']
print(tr[['type','amp','n_photon','n_pe','n_electron','raw_area','t_first_photon','t_sigma_photon','t_mean_electron']][:6])
rr=out['raw_records']; print(rr['data'].sum(), np.all(np.diff(rr['time'])>=0))
out2 = sim.simulate(inst, seed=5)
print('reproducible', out2['raw_records'].tobytes()==rr.tobytes(), out2['truth'].tobytes()==tr.tobytes())
