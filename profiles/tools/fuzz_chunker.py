"""Randomised check of the chunker: ChunkRawRecords with random chunk sizes, piece sizes, record buffers and event
rates; all chunks together must equal one un-chunked call (records byte-identical, one truth row per Pulse call) and
the chunk bounds must be what chunk_boundaries computes from the groups of that call (the emulation of the reference's
bookkeeping that tests/golden/chunks.json pins to the reference class).  Not a test.  usage: fuzz_chunker.py [seed] [n]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.conftest import load_c0_config
from tests.golden.synth_instructions import c0_like
from tests.golden.synth_tables import EleApHist, pmt_ap_tables
from tests.test_gpu_afterpulse_plugin import spe
from wfsim_b200.strax_interface import ChunkRawRecords, chunk_boundaries

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 20
uniq, row = spe()
bad = 0
for it in range(n_iter):
    ap = bool(rng.integers(2))
    extra = dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True) if ap else {}
    cfg = load_c0_config(**extra)
    cfg = dict(cfg, chunk_size=float(rng.choice([0.05, 0.5, 2, 20])), b200_piece_instructions=int(rng.choice([3, 10, 40000])))
    buf = int(rng.choice([0, 2500, 20000]))
    if buf:
        cfg['b200_record_buffer'] = buf
    kw = dict(uniform_to_pmt_ap=pmt_ap_tables(), uniform_to_ele_ap=EleApHist()) if ap else {}
    n_ev = int(rng.integers(2, 30))
    inst = c0_like(n_ev, seed=int(rng.integers(1 << 30)), event_rate=float(10 ** rng.uniform(0, 2.5)), e_range=(0.5, 8))
    try:
        crr = ChunkRawRecords(cfg, spe_ppf=uniq, spe_row=row, seed=int(it), **kw)
        chunks, bounds = [], []
        for res in crr(inst):
            chunks.append({k: np.array(v) for k, v in res.items()})
            bounds.append((int(crr.chunk_time_pre), int(crr.chunk_time)))
        one = crr.simulator.simulate(inst, seed=int(it))
        rr = np.concatenate([c['raw_records'] for c in chunks])
        if rr.tobytes() != one['raw_records'].tobytes():
            raise AssertionError(f'records differ: {len(rr)} vs {len(one["raw_records"])}')
        if sum(len(c['truth']) for c in chunks) != len(one['truth']):
            raise AssertionError('truth rows')
        for (pre, ct), c in zip(bounds, chunks):
            r = c['raw_records']
            if len(r) and not (r['time'].min() > pre and r['time'].max() <= ct):
                raise AssertionError(f'records outside their chunk {pre} {ct} {r["time"].min()} {r["time"].max()}')
        g = one['groups']
        t = np.sort(np.concatenate([one[k]['time'] for k in ('raw_records', 'raw_records_he', 'raw_records_aqmon')]))
        n_rec = np.searchsorted(t, (g['right'] + 1) * cfg['sample_duration']) - np.searchsorted(t, g['left'] * cfg['sample_duration'])
        want = chunk_boundaries(cfg, inst['time'].min(), g, n_records=n_rec, record_buffer=buf or None)
        if bounds != want:
            raise AssertionError(f'bounds {bounds[:4]} vs {want[:4]} ({len(bounds)} / {len(want)})')
        crr.simulator.close()
    except ValueError as e:
        if 'insufficient record buffer' in str(e):      # a single group larger than the buffer: refused by design
            continue
        bad += 1
        print('PROBLEM at iteration', it, repr(e)[:400], flush=True)
    except Exception as e:      # noqa
        bad += 1
        print('PROBLEM at iteration', it, 'ap', ap, 'chunk', cfg['chunk_size'], 'piece', cfg['b200_piece_instructions'], 'buf', buf, 'events', n_ev,
              repr(e)[:400], flush=True)
print('done', n_iter, 'iterations,', bad, 'problems')
