import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from tests.golden.synth_instructions import c1_like
from tests.test_gpu_afterpulse_plugin import make_sim
from wfsim_b200.dtypes import raw_record_dtype
sim, cfg = make_sim(enable_pmt_afterpulses=True)
inst = c1_like(400, seed=9)
ref = sim.simulate(inst, seed=5)
n = len(ref['raw_records'])
dest = np.empty(n + 5000, raw_record_dtype())
dest.view(np.uint8)[:] = 0xAB
assert sim.pin(dest)
os.environ['WFS_BATCH_INSTRUCTIONS'] = '100'
for frac in sys.argv[1:]:
    os.environ['WFS_PLAIN_FRACTION'] = frac
    for rep in range(3):
        dest.view(np.uint8)[:] = 0xAB
        out = sim.simulate(inst, seed=5, records_out=dest)
        c = sim.last_counts
        a = out['raw_records'].view(np.uint8).reshape(-1, 244); b = ref['raw_records'].view(np.uint8).reshape(-1, 244)
        print('lens', len(a), len(b), c['n_records_total'], c['n_records'], len(out['raw_records_he']), sim.last_counts['n_fused_batches'])
        m = min(len(a), len(b)); a, b = a[:m], b[:m]
        bad = np.flatnonzero((a != b).any(axis=1))
        print(frac, rep, 'n', n, len(a), 'batches', c['n_batches'], 'plain', c['n_plain_records'], 'bad', len(bad), bad[:10], bad[-5:])
        if len(bad):
            i = bad[0]; print(' first bad cols', np.flatnonzero(a[i] != b[i])[:12], a[i][:24], b[i][:24])
