"""GPU: the error paths of the hot path, with the reference's exception types and texts.

  * "Pulse cache too long" (rawdata.py:219): a digitisation group of 1e6 samples or more;
  * WFS_E_CAPACITY: a record buffer that is too small makes the C entry return the need and write nothing the
    caller may use; the repeated call with that capacity returns the same bytes as a roomy first call;
  * the record-buffer rule of the chunker (strax_interface.py:409-418): a chunk whose records do not fit the
    buffer is closed early at the end of the previous digitisation group (bounds pinned to the unmodified reference
    by tests/golden/chunks.json, tests/test_host_interface.py)."""
import ctypes as C

import numpy as np
import pytest

from tests.golden.synth_instructions import c0_like, c1_like
from tests.test_gpu_afterpulse_plugin import make_sim
from wfsim_b200 import lib as wlib
from wfsim_b200.dtypes import raw_record_dtype, truth_dtype
from wfsim_b200.simulator import _ptr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('fused', ['1', '0'])
def test_pulse_cache_too_long(fused, monkeypatch):
    """S1s 90 us apart chain into one digitisation group (gap < right_raw_extension = 100 us); 130 of them
    span more than 1e6 samples of 10 ns."""
    monkeypatch.setenv('WFS_FUSED', fused)
    sim, cfg = make_sim()
    inst = c0_like(130, seed=3, e_range=(1, 3))
    inst = inst[inst['type'] == 1]
    inst['time'] = 1_000_000 + 90_000 * np.arange(len(inst))
    assert (inst['time'][-1] - inst['time'][0]) / cfg['sample_duration'] > 1_000_000
    with pytest.raises(AssertionError, match='Pulse cache too long'):
        sim.simulate(inst, seed=1)
    # the handle is still usable, and a shorter chain is fine
    out = sim.simulate(inst[:100], seed=1)
    assert len(out['groups']) == 1 and len(out['raw_records']) > 0
    sim.close()


def call_simulate(sim, inst, seed, cap_records):
    """wfs_simulate through ctypes with a given record capacity (what Simulator.simulate wraps in its retry loop)."""
    n = len(inst)
    m, keep = sim._maps_struct(inst, None, None, seed=seed)
    rec = np.zeros(max(cap_records, 1), raw_record_dtype())
    truth = np.zeros(2 * n + 64, truth_dtype())
    groups = np.zeros(n + 64, np.dtype([('left', np.int64), ('right', np.int64), ('n_intervals', np.int64)]))
    batch_records = np.zeros((4096, 3), np.int64)
    out = wlib.Outputs(_ptr(rec), cap_records, _ptr(truth), len(truth), _ptr(groups), len(groups),
                       _ptr(batch_records), 4096, None, None)
    counts = wlib.Counts()
    rc = sim.lib.wfs_simulate(sim.handle, _ptr(inst.view(np.uint8)), n, C.byref(m), int(seed), C.byref(out), C.byref(counts))
    return rc, counts, rec, truth


@pytest.mark.parametrize('fused', ['1', '0'])
def test_capacity_error_and_repeated_call(fused, monkeypatch):
    monkeypatch.setenv('WFS_FUSED', fused)
    monkeypatch.setenv('WFS_BATCH_INSTRUCTIONS', '60')      # several device batches: the need is the sum over all of them
    sim, cfg = make_sim(enable_pmt_afterpulses=True)
    inst = c1_like(150, seed=4)
    want = sim.simulate(inst, seed=9)
    n = len(want['raw_records'])
    assert sim.last_counts['n_batches'] > 2
    rc, counts, rec, truth = call_simulate(sim, inst, 9, cap_records=n // 3)
    assert rc == wlib.E_CAPACITY
    assert counts.need_records == n and counts.need_records > n // 3
    # the repeated call with the capacity asked for
    rc, counts, rec, truth = call_simulate(sim, inst, 9, cap_records=int(counts.need_records))
    assert rc == 0 and counts.n_records_total == n
    assert rec[:n].tobytes() == want['raw_records'].tobytes()
    assert truth[:counts.n_truth].tobytes() == want['truth'].tobytes()
    # exactly one row too few is still an error, exactly enough is not
    rc, counts, _, _ = call_simulate(sim, inst, 9, cap_records=n - 1)
    assert rc == wlib.E_CAPACITY and counts.need_records == n
    sim.close()


def test_chunker_closes_a_chunk_early_when_the_record_buffer_is_full():
    """ChunkRawRecords with a record buffer of 4000 records (config b200_record_buffer; the reference's is 5e6):
    chunks are closed at the end of the last group that fits, the warning text is the reference's, every record
    is delivered exactly once and in order, and the bounds are what the chunk clock computes from the groups."""
    import logging
    from wfsim_b200.strax_interface import ChunkRawRecords, chunk_boundaries
    sim, cfg = make_sim()
    sim.close()
    cfg = dict(cfg, chunk_size=5, b200_record_buffer=4000, b200_piece_instructions=8)
    from tests.test_gpu_afterpulse_plugin import spe
    uniq, row = spe()
    inst = c0_like(16, seed=8, event_rate=8.0, e_range=(1, 8))         # 2 s of data: one chunk by the clock alone
    crr = ChunkRawRecords(cfg, spe_ppf=uniq, spe_row=row, seed=5)
    messages = []
    handler = logging.Handler()
    handler.emit = lambda r: messages.append(r.getMessage())
    logging.getLogger('wfsim_b200.interface').addHandler(handler)
    try:
        chunks, bounds = [], []
        for res in crr(inst):
            chunks.append(res)
            bounds.append((crr.chunk_time_pre, crr.chunk_time))
    finally:
        logging.getLogger('wfsim_b200.interface').removeHandler(handler)
    assert len(chunks) > 2 and any('insufficient record buffer' in m for m in messages)
    one = crr.simulator.simulate(inst, seed=5)
    rr = np.concatenate([c['raw_records'] for c in chunks])
    assert rr.tobytes() == one['raw_records'].tobytes()
    for (pre, ct), c in zip(bounds, chunks):
        r = c['raw_records']
        assert len(r) <= 4000
        if len(r):
            assert r['time'].min() > pre and r['time'].max() <= ct
    g = one['groups']
    t = one['raw_records']['time']
    n_rec = np.searchsorted(t, (g['right'] + 1) * cfg['sample_duration']) - np.searchsorted(t, g['left'] * cfg['sample_duration'])
    want = chunk_boundaries(cfg, inst['time'].min(), g, n_records=n_rec, record_buffer=4000)
    assert bounds == want
    assert len(np.concatenate([c['truth'] for c in chunks])) == len(inst)
