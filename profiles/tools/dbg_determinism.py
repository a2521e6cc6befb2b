"""Repeats one call and prints a hash of the records: the result must not change from call to call."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from tests.golden.synth_instructions import c1_like
from tests.test_gpu_afterpulse_plugin import make_sim
ap = int(sys.argv[1]) if len(sys.argv) > 1 else 1
sim, cfg = make_sim(enable_pmt_afterpulses=bool(ap))
inst = c1_like(400, seed=9)
for fused in ('1', '0', '1'):
    os.environ['WFS_FUSED'] = fused
    for batch in (None, '100'):
        if batch: os.environ['WFS_BATCH_INSTRUCTIONS'] = batch
        else: os.environ.pop('WFS_BATCH_INSTRUCTIONS', None)
        for rep in range(3):
            out = sim.simulate(inst, seed=5)
            ph = sim.sample_stage(inst, stage=0, seed=5)
            r = out['raw_records']
            print('fused', fused, 'batch', batch, 'rep', rep, len(r), hashlib.md5(r.tobytes()).hexdigest()[:10],
                  'photons', len(ph), hashlib.md5(np.sort(ph, order=['instruction', 't', 'channel']).tobytes()).hexdigest()[:10],
                  hashlib.md5(out['truth'].tobytes()).hexdigest()[:8], flush=True)
