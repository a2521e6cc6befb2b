"""Pin the CPU oracle (oracle/) against the golden vectors produced by the unmodified reference.

Bit-exact: record headers and ADC data must be identical to what WFSim's own
Pulse.__call__/add_current -> digitize_pulse_cache -> ZLE -> ChunkRawRecords produced.
"""
import os

import numpy as np
import pytest

from tests.conftest import DET_CASES, GOLDEN, load_c0_config, load_det_case
from oracle import wfsim_oracle as orc


def test_templates_match_reference():
    cfg = load_c0_config()
    z = np.load(os.path.join(GOLDEN, 'c0_tables.npz'))
    tm = orc.pmt_current_templates(cfg)
    assert tm.shape == (10, 22)
    np.testing.assert_array_equal(tm, z['templates'])
    # reference invariant (pulse.py:180-181): every template integrates to one pe
    np.testing.assert_allclose(tm.sum(axis=1) * cfg['sample_duration'], 1.0, rtol=1e-14)
    assert orc.current_2_adc(cfg) == float(z['current_2_adc'])


@pytest.mark.parametrize('name', DET_CASES)
def test_oracle_reproduces_reference_records(name):
    c = load_det_case(name)
    out = orc.simulate_photons(c['cfg'], c['pcall'], c['channel'], c['t'], c['gain'],
                               c['group_of'], noise=c['noise'],
                               ix_rand=c['ix_rand'] if c['noise'] is not None else None)
    for got, want in ((out['raw_records'], c['rr']), (out['raw_records_he'], c['rr_he'])):
        assert len(got) == len(want)
        for f in ('time', 'length', 'dt', 'channel', 'pulse_length', 'record_i', 'baseline'):
            np.testing.assert_array_equal(got[f], want[f], err_msg=f)
        np.testing.assert_array_equal(got['data'], want['data'])
    assert len(out['raw_records_aqmon']) == 0


def test_find_intervals_merge_rule():
    """utils.py:13-58: strict '<', two flagged samples merge iff index distance <= holdoff."""
    import ctypes
    L = orc.lib()

    def run(w, thr=10, hold=101):
        w = np.asarray(w, np.int64)
        out = np.zeros((16, 2), np.int64)
        n = L.orc_find_intervals(orc._p(w), ctypes.c_int64(len(w)), ctypes.c_int64(thr),
                                 ctypes.c_int64(hold), orc._p(out), ctypes.c_int64(16))
        return out[:n].tolist()
    base = np.full(400, 20)
    w = base.copy(); w[[50, 151]] = 0
    assert run(w) == [[50, 151]]
    w = base.copy(); w[[50, 152]] = 0
    assert run(w) == [[50, 50], [152, 152]]
    w = base.copy(); w[50] = 10          # equal to threshold is NOT below
    assert run(w) == []
    w = base.copy(); w[398] = 0          # closes at the last sample
    assert run(w) == [[398, 398]]
    assert run([]) == []
