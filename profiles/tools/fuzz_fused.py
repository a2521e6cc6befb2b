"""Randomised A/B of the two back ends: low-energy instruction sets (every group fits a CTA) through wfs_simulate with
the fused back end and with the multi-pass back end -- records, truth rows, groups and counters must be identical --
under random device-batch sizes, lane counts, destinations (pageable / page-locked with a random plain-row share) and
configurations (afterpulses, photo-ionisation, merged truth).  Not a test (run on a GPU box).
usage: fuzz_fused.py [seed] [iterations]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_gpu_afterpulse_plugin import make_sim
from wfsim_b200.dtypes import instruction_dtype, raw_record_dtype

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(seed0)
configs = [dict(), dict(enable_pmt_afterpulses=True), dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True),
           dict(enable_pmt_afterpulses=True, save_full_truth=False), dict(zle_threshold=40)]
sims = [make_sim(**c) for c in configs]
dest = np.empty(3_000_000, raw_record_dtype())
sims[0][0].pin(dest)
bad = 0
for it in range(n_iter):
    k = int(rng.integers(len(sims)))
    sim, cfg = sims[k]
    n = int(rng.integers(1, 60)) if it % 3 else int(rng.integers(200, 1500))
    inst = np.zeros(n, instruction_dtype)
    inst['type'] = rng.choice([1, 2], n)
    span = int(10 ** rng.uniform(4.5, 9.5))
    inst['time'] = np.sort(rng.integers(0, span, n))
    r = np.sqrt(rng.uniform(0, 50 ** 2, n)); th = rng.uniform(-np.pi, np.pi, n)
    inst['x'], inst['y'] = r * np.cos(th), r * np.sin(th)
    inst['z'] = rng.uniform(-97, 0, n)
    inst['amp'] = np.where(inst['type'] == 1, (10 ** rng.uniform(0, 3.0, n)).astype(int), (10 ** rng.uniform(0, 2.2, n)).astype(int))
    inst['recoil'], inst['local_field'], inst['event_number'] = 7, 82.0, np.arange(n)
    inst = inst[inst['amp'] > 0]
    os.environ['WFS_BATCH_INSTRUCTIONS'] = str(int(rng.choice([5, 37, 400000])))
    os.environ['WFS_LANES'] = str(int(rng.choice([1, 2, 4])))
    pinned = rng.random() < 0.4
    if pinned:
        os.environ['WFS_PLAIN_FRACTION'] = str(float(rng.choice([0, 0.3, 0.77, 1])))
    else:
        os.environ.pop('WFS_PLAIN_FRACTION', None)
    only = os.environ.get('FUZZ_ONLY')
    if only is not None and int(only) != it:
        continue
    if only is not None:
        for batch in ('37', '5', '400000'):
            os.environ['WFS_BATCH_INSTRUCTIONS'] = batch
            res = {}
            for fused in ('1', '0'):
                os.environ['WFS_FUSED'] = fused
                o = sim.simulate(inst, seed=it)
                res[fused] = (o['truth'].copy(), dict(sim.last_counts))
            ta, tb = res['1'][0], res['0'][0]
            d = np.flatnonzero(ta['n_pe_trigger'] != tb['n_pe_trigger'])
            print('batch', batch, 'batches', res['1'][1]['n_batches'], 'fused', res['1'][1]['n_fused_batches'], 'rows differing', d[:12],
                  'type', ta['type'][d][:12], 'n_photon', ta['n_photon'][d][:12], 'fused', ta['n_pe_trigger'][d][:12], 'multi', tb['n_pe_trigger'][d][:12],
                  'n_pe', ta['n_pe'][d][:12], 'n_photon_trigger', ta['n_photon_trigger'][d][:12], tb['n_photon_trigger'][d][:12])
        continue
    try:
        os.environ['WFS_FUSED'] = '1'
        a = sim.simulate(inst, seed=it, records_out=dest if pinned else None)
        ca = dict(sim.last_counts)
        a = {k2: np.array(v) for k2, v in a.items() if isinstance(v, np.ndarray)}
        os.environ['WFS_FUSED'] = '0'
        b = sim.simulate(inst, seed=it)
        cb = dict(sim.last_counts)
        for key in ('raw_records', 'raw_records_he', 'truth', 'groups'):
            if a[key].tobytes() != b[key].tobytes():
                detail = ''
                if len(a[key]) == len(b[key]) and a[key].dtype.names:
                    for f in a[key].dtype.names:
                        x, y = a[key][f], b[key][f]
                        neq = ~((x == y) | ((x != x) & (y != y)))
                        if neq.any():
                            i = int(np.flatnonzero(neq)[0])
                            detail += f' {f}[{i}]: {x[i]} vs {y[i]} ({int(neq.sum())} rows);'
                raise AssertionError(f'{key} differ: {len(a[key])} vs {len(b[key])};{detail}')
        for key in ('n_records_total', 'n_truth', 'n_photons', 'n_pe', 'n_pulses', 'n_intervals', 'n_samples', 'n_groups'):
            if ca[key] != cb[key]:
                raise AssertionError(f'count {key}: {ca[key]} vs {cb[key]}')
        rr = a['raw_records']
        if len(rr) and (np.diff(rr['time'].astype(np.int64) * 1024 + rr['channel']) < 0).any() and ca['n_batches'] == 1:
            raise AssertionError('records out of order')
    except AssertionError as e:
        if 'Pulse cache too long' in str(e):        # a group of 1e6 samples or more (rawdata.py:219): both back ends must say so
            os.environ['WFS_FUSED'] = '0'
            try:
                sim.simulate(inst, seed=it)
                bad += 1
                print('PROBLEM at iteration', it, ': only the fused back end found the pulse cache too long', flush=True)
            except AssertionError:
                pass
            continue
        bad += 1
        print('PROBLEM at iteration', it, 'config', k, 'n', len(inst), 'batch', os.environ['WFS_BATCH_INSTRUCTIONS'],
              'lanes', os.environ['WFS_LANES'], 'pinned', pinned, os.environ.get('WFS_PLAIN_FRACTION'), repr(e)[:600], flush=True)
    except Exception as e:      # noqa
        bad += 1
        print('PROBLEM at iteration', it, 'config', k, 'n', len(inst), 'batch', os.environ['WFS_BATCH_INSTRUCTIONS'],
              'lanes', os.environ['WFS_LANES'], 'pinned', pinned, os.environ.get('WFS_PLAIN_FRACTION'), repr(e)[:300], flush=True)
    if it % 25 == 24:
        print('iteration', it + 1, 'fused batches so far ok; last counts', ca['n_batches'], ca['n_fused_batches'], flush=True)
print('done', n_iter, 'iterations,', bad, 'problems')
