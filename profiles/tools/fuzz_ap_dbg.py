import os, sys
import numpy as np
sys.path.insert(0, '/root/repo')
os.environ['FUZZ_AP'] = '1'
from tests.test_gpu_configs import make_sim as make_sim_res
from tests.golden.synth_tables import EleApHist, pmt_ap_tables
from wfsim_b200.dtypes import instruction_dtype
sim, cfg = make_sim_res(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                        enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
rng = np.random.default_rng(5)
nbad = 0
for it in range(160):
    n = int(rng.integers(0, 40)) if it % 4 else int(rng.integers(100, 600))
    inst = np.zeros(n, instruction_dtype)
    inst['type'] = rng.choice([1, 2], n)
    inst['time'] = np.sort(rng.integers(0, int(10 ** rng.uniform(4, 9)), n)) if rng.random() < 0.7 else rng.integers(0, 10 ** 8, n)
    r = np.sqrt(rng.uniform(0, 55 ** 2, n)); th = rng.uniform(-np.pi, np.pi, n)
    inst['x'], inst['y'] = r * np.cos(th), r * np.sin(th)
    inst['z'] = rng.uniform(-110, 5, n)
    inst['amp'] = np.where(rng.random(n) < 0.2, rng.integers(1, 4, n), (10 ** rng.uniform(0, 3.7 if n < 100 else 2.5, n)).astype(int))
    inst['recoil'] = 7; inst['local_field'] = 82.0; inst['event_number'] = np.arange(n)
    inst = inst[inst['amp'] > 0]
    bi = str(int(rng.choice([3, 7, 400000])))
    os.environ['WFS_BATCH_INSTRUCTIONS'] = bi
    res = {}
    for mode in ('0', '1'):
        os.environ['WFS_SEGMENT_SORT'] = mode
        out = sim.simulate(inst, seed=it)
        rr = np.array(out['raw_records'])
        key = rr['time'].astype(np.int64) * 1024 + rr['channel']
        res[mode] = (rr, bool((np.diff(key) >= 0).all()), sim.last_counts['ms_phase'][10], sim.last_counts['ms_phase'][11], sim.last_counts['n_batches'])
    same = res['0'][0].tobytes() == res['1'][0].tobytes()
    if not (res['0'][1] and res['1'][1] and same):
        nbad += 1
        print('iter', it, 'n', len(inst), 'batch_instr', bi, 'sorted radix/seg', res['0'][1], res['1'][1], 'same', same,
              'seg batches', res['1'][2], res['1'][3], 'batches', res['1'][4], 'records', len(res['0'][0]), len(res['1'][0]), flush=True)
        if nbad < 3 and not res['0'][1]:
            rr = res['0'][0]; key = rr['time'].astype(np.int64) * 1024 + rr['channel']
            bad = np.flatnonzero(np.diff(key) < 0)[:3]
            for b in bad:
                print('   at', b, rr['time'][b - 1:b + 3], rr['channel'][b - 1:b + 3], rr['record_i'][b - 1:b + 3])
print('bad', nbad)
