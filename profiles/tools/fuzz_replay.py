"""Randomised differential test of the FULL path against the oracle's replay of the reference scheduler, Pulse calls,
digitiser, ZLE, record packer and truth (oracle.wfsim_oracle_sim.ReplayOracle, pinned to the unmodified reference by
tests/test_oracle_replay.py): random small instruction sets incl. pile-up, random configurations; records and groups
byte-identical, truth integers exact.  Not a test (run on a GPU box).  usage: fuzz_replay.py [seed] [iterations]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.wfsim_oracle_sim import ReplayOracle
from tests.golden.synth_instructions import c0_like
from tests.test_gpu_afterpulse_plugin import make_sim
from tests.test_gpu_replay import presets_from_dumps
from wfsim_b200.dtypes import truth_dtype

seed0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(seed0)
configs = [dict(), dict(enable_pmt_afterpulses=True), dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True),
           dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True, save_full_truth=False),
           dict(enable_gate_afterpulses=True, photoelectric_p=0.004)]
sims = [make_sim(**c) for c in configs]
bad = 0
for it in range(n_iter):
    k = int(rng.integers(len(sims)))
    sim, cfg = sims[k]
    n_ev = int(rng.integers(1, 14))
    rate = float(10 ** rng.uniform(0, 4.6))                 # 1 Hz .. 40 kHz: separate events to heavy pile-up
    emax = float(10 ** rng.uniform(-0.5, 1.3))
    inst = c0_like(n_ev, seed=int(rng.integers(1 << 30)), event_rate=rate, e_range=(0.1, max(emax, 0.2)))
    os.environ['WFS_FUSED'] = str(int(rng.integers(2)))
    # a FORCED batch cut (no quiet gap within twice the budget) inside a pile-up of events with delayed secondaries is a
    # documented deviation (the secondaries of the first half cannot merge with groups of the
    # second); tiny budgets provoke it, so they are only drawn for configurations without secondaries
    secondaries = cfg.get('enable_electron_afterpulses') or cfg.get('enable_gate_afterpulses')
    os.environ['WFS_BATCH_INSTRUCTIONS'] = '400000' if secondaries else str(int(rng.choice([4, 400000])))
    try:
        out = sim.simulate(inst, seed=1000 + it)
        c = dict(sim.last_counts)
        photons, electrons, sec_rows = presets_from_dumps(sim, inst, 1000 + it)
        want = ReplayOracle(cfg, photons, electrons, sec_rows).simulate(inst, truth_dtype=truth_dtype())
        g = out['groups']
        got_groups = np.stack([g['left'], g['right'], g['n_intervals']], axis=1) if len(g) else np.zeros((0, 3), np.int64)
        if not np.array_equal(got_groups, np.array(want['groups'], np.int64).reshape(-1, 3)):
            raise AssertionError(f'groups differ: {len(got_groups)} vs {len(want["groups"])}')
        he0 = cfg['channel_map']['he'][0]
        rec = want['records']
        if out['raw_records'].tobytes() != rec[rec['channel'] < he0].tobytes():
            raise AssertionError(f'records differ: {len(out["raw_records"])} vs {(rec["channel"] < he0).sum()}')
        if len(out['truth']) != len(want['truth']):
            raise AssertionError(f'truth rows: {len(out["truth"])} vs {len(want["truth"])}')
        for f in want['truth'].dtype.names:
            if want['truth'].dtype[f].kind in 'iu':
                if not np.array_equal(out['truth'][f], want['truth'][f]):
                    i = int(np.flatnonzero(out['truth'][f] != want['truth'][f])[0])
                    raise AssertionError(f'truth {f}[{i}]: {out["truth"][f][i]} vs {want["truth"][f][i]}')
    except Exception as e:      # noqa
        if 'Pulse cache too long' in str(e):
            continue
        bad += 1
        print('PROBLEM at iteration', it, 'config', k, 'events', n_ev, 'rate', round(rate), 'emax', round(emax, 2), 'fused',
              os.environ['WFS_FUSED'], 'batch', os.environ['WFS_BATCH_INSTRUCTIONS'], repr(e)[:400], flush=True)
    if it % 10 == 9:
        print('iteration', it + 1, 'last: groups', len(out['groups']), 'records', len(out['raw_records']), 'fused batches', c['n_fused_batches'], flush=True)
print('done', n_iter, 'iterations,', bad, 'problems')
