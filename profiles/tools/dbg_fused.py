import os, sys
sys.path.insert(0, '/root/repo')
os.environ['WFS_DEBUG_SEG'] = '1'
os.environ['WFS_BATCH_INSTRUCTIONS'] = '10'
from tests.test_gpu_afterpulse_plugin import make_sim
from tests.golden.synth_instructions import c0_like
sim, cfg = make_sim(enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
inst = c0_like(40, seed=11, e_range=(1, 6))
out = sim.simulate(inst, seed=2)
print(sim.last_counts['n_batches'], sim.last_counts['n_fused_batches'])
