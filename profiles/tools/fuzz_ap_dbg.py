"""Replays one iteration of fuzz_simulate.py (FUZZ_AP=1) and reports where the records are out of order.
usage: fuzz_ap_dbg.py <seed> <iteration>"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_gpu_configs import make_sim as make_sim_res
from tests.golden.synth_tables import EleApHist, pmt_ap_tables
from wfsim_b200.dtypes import instruction_dtype
sim, cfg = make_sim_res(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                        enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
rng = np.random.default_rng(int(sys.argv[1]))
target = int(sys.argv[2])
for it in range(target + 1):
    n = int(rng.integers(0, 40)) if it % 4 else int(rng.integers(100, 600))
    inst = np.zeros(n, instruction_dtype)
    inst['type'] = rng.choice([1, 2], n)
    inst['time'] = np.sort(rng.integers(0, int(10 ** rng.uniform(4, 9)), n)) if rng.random() < 0.7 else rng.integers(0, 10 ** 8, n)
    r = np.sqrt(rng.uniform(0, 55 ** 2, n)); th = rng.uniform(-np.pi, np.pi, n)
    inst['x'], inst['y'] = r * np.cos(th), r * np.sin(th)
    inst['z'] = rng.uniform(-110, 5, n)
    inst['amp'] = np.where(rng.random(n) < 0.2, rng.integers(1, 4, n), (10 ** rng.uniform(0, 3.7 if n < 100 else 2.5, n)).astype(int))
    inst['recoil'] = 7
    inst['local_field'] = 82.0
    inst['event_number'] = np.arange(n)
    inst = inst[inst['amp'] > 0]
print('iteration', target, 'instructions', len(inst), 'time span', inst['time'].min(), inst['time'].max())
res = {}
for fused in ('1', '0'):
    os.environ['WFS_FUSED'] = fused
    out = sim.simulate(inst, seed=target)
    c = sim.last_counts
    rr = out['raw_records']
    key = rr['time'].astype(np.int64) * 1024 + rr['channel']
    bad = np.flatnonzero(np.diff(key) < 0)
    print('fused', fused, 'batches', c['n_batches'], 'fused batches', c['n_fused_batches'], 'records', len(rr), 'groups', len(out['groups']),
          'out of order at', bad[:10], 'n', len(bad))
    g = out['groups']
    if len(bad):
        for b in bad[:3]:
            print('  ', rr['time'][b - 1:b + 3], rr['channel'][b - 1:b + 3], rr['record_i'][b - 1:b + 3], rr['pulse_length'][b - 1:b + 3])
            t = rr['time'][b] // 10
            gi = np.flatnonzero((g['left'] <= t) & (g['right'] >= t))
            print('   groups containing it:', gi, g[gi] if len(gi) else None)
    ov = np.flatnonzero(g['left'][1:] <= g['right'][:-1])
    print('  overlapping consecutive groups:', ov[:10], [(int(g['left'][i]), int(g['right'][i]), int(g['left'][i+1]), int(g['right'][i+1])) for i in ov[:3]])
    res[fused] = out
a, b = res['1'], res['0']
print('fused == multi-pass:', a['raw_records'].tobytes() == b['raw_records'].tobytes(), len(a['raw_records']), len(b['raw_records']))
