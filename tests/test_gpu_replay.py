"""GPU: scheduler, truth rows and records of the FULL path (wfs_simulate) against the oracle on the
photons the GPU generated itself -- with nothing of the expectation taken from the GPU's own grouping.

The sampling front end is dumped per instruction (wfs_sample_stage: photons with their instruction /
secondary identity and double-pe / afterpulse flags, electrons, the secondary instructions every S2
spawned).  oracle.wfsim_oracle_sim.ReplayOracle then runs the reference's scheduler (rawdata.py:38-157),
Pulse calls, digitiser, ZLE, record packer and get_truth / add_truth (rawdata.py:313-390, pulse.py:229-271)
on those presets.  ReplayOracle itself is pinned to the UNMODIFIED reference on preset stage outputs by
tests/test_oracle_replay.py (tests/golden/sched.npz).  Required here: records byte-identical, digitisation
groups (left, right, n_intervals) identical, every truth field equal (integers exactly; time moments
against exact integer arithmetic at 1e-12; areas at 1e-9).  SURVEY.md rows a4, a5, a29, a34."""
import os
from fractions import Fraction

import numpy as np
import pytest

from tests.golden.make_golden_sched import EL_DT, PH_DT
from tests.golden.synth_instructions import c0_like
from tests.golden.synth_tables import noise_sample
from tests.test_gpu_afterpulse_plugin import make_sim
from wfsim_b200.dtypes import instruction_dtype, truth_dtype

pytestmark = pytest.mark.gpu
IDT = np.dtype(instruction_dtype)


def presets_from_dumps(sim, inst, seed):
    n_prim = len(inst)
    ph = sim.sample_stage(inst, stage=0, seed=seed)
    el = sim.sample_stage(inst, stage=1, seed=seed)
    sec = sim.sample_secondaries(inst, seed=seed)
    photons = np.zeros(len(ph), PH_DT)
    is_sec = (ph['flags'] & 4) != 0
    photons['id'] = np.where(is_sec, n_prim + ph['secondary'], ph['instruction'])
    photons['t'], photons['channel'], photons['gain'] = ph['t'], ph['channel'], ph['gain']
    photons['dpe'] = ph['flags'] & 1
    photons['ap'] = (ph['flags'] >> 1) & 1
    # electrons: the emitter dump lists S1 vertices too (one per S1 with hits) -- electrons are S2-like only
    is_sec_e = (el['flags'] & 4) != 0
    e_id = np.where(is_sec_e, n_prim + el['secondary'], el['instruction'])
    s2like = is_sec_e | (inst['type'][np.clip(el['instruction'], 0, n_prim - 1)] != 1)
    electrons = np.zeros(int(s2like.sum()), EL_DT)
    electrons['id'], electrons['t'] = e_id[s2like], el['t'][s2like]
    # secondary rows as the reference builds them: a copy of the parent's row with type, time, position
    # and amp replaced (afterpulse.py:54-61, 125-133)
    sdt = np.dtype([(n, IDT[n]) for n in IDT.names] + [('_id', np.int64), ('_parent', np.int64)])
    rows = np.zeros(len(sec), sdt)
    for n in IDT.names:
        rows[n] = inst[n][sec['parent']]
    for n in ('time', 'x', 'y', 'z', 'amp', 'type'):
        rows[n] = sec[n]
    rows['_id'] = n_prim + np.arange(len(sec))
    rows['_parent'] = sec['parent']
    return photons, electrons, rows


def exact_moments(t):
    """mean and population standard deviation of integer times in exact arithmetic."""
    t = [int(x) for x in t]
    n = len(t)
    s1, s2 = sum(t), sum(x * x for x in t)
    var = Fraction(n * s2 - s1 * s1, n * n)
    return float(Fraction(s1, n)), float(var) ** 0.5


def check_truth(got, want, photons, electrons, runs):
    assert len(got) == len(want)
    exact_fields = [n for n in want.dtype.names if want.dtype[n].kind in 'iu' or n in (
        'x', 'y', 'z', 'e_dep', 'tot_e', 'local_field', 'x_pri', 'y_pri', 'z_pri',
        't_first_photon', 't_last_photon', 't_first_electron', 't_last_electron')]
    for n in want.dtype.names:
        a, b = got[n], want[n]
        if want.dtype[n].kind == 'f':
            assert np.array_equal(np.isnan(a), np.isnan(b)), n
            ok = ~np.isnan(b)
            a, b = a[ok], b[ok]
        if n in exact_fields:
            assert np.array_equal(a, b), n
        elif n.startswith('raw_area'):
            assert np.allclose(a, b, rtol=1e-9, atol=1e-9), n
        else:     # t_mean_*, t_sigma_*: the oracle's np.mean / np.std carry float error at |t| ~ 1e10
            assert np.allclose(a, b, rtol=1e-6, atol=1e-4), n
    # the time moments against exact integer arithmetic
    rows_with_truth = [r for r in runs if r[3]]
    assert len(rows_with_truth) == len(got)
    for row, (typ, ids, grp, _) in zip(got, rows_with_truth):
        for q, arr in (('photon', photons[photons['ap'] == 0]), ('electron', electrons)):
            t = arr['t'][np.isin(arr['id'], ids)]
            if q == 'electron' and typ == 1:
                t = t[:0]
            if len(t):
                m, s = exact_moments(t)
                assert row[f't_mean_{q}'] == pytest.approx(m, rel=1e-14)
                assert row[f't_sigma_{q}'] == pytest.approx(s, rel=1e-12, abs=1e-9)


CASES = {
    'plain': dict(cfg={}, n=14, seed=3, kw={}),
    'afterpulses': dict(cfg=dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True), n=12, seed=4,
                        kw=dict(e_range=(5, 60))),
    'pile_up_merged_truth': dict(cfg=dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True,
                                          save_full_truth=False), n=16, seed=6, kw=dict(event_rate=4000.0, e_range=(2, 30))),
    'gate': dict(cfg=dict(enable_gate_afterpulses=True, photoelectric_p=0.004, enable_pmt_afterpulses=True), n=10,
                 seed=8, kw={}),
    'noise': dict(cfg=dict(enable_noise=True, enable_pmt_afterpulses=True, enable_electron_afterpulses=True), n=8,
                  seed=9, kw=dict(e_range=(5, 40)), noise=True),
}


@pytest.mark.parametrize('name', list(CASES))
def test_full_path_equals_reference_scheduler_and_truth_on_gpu_photons(name):
    from oracle.wfsim_oracle_sim import ReplayOracle
    case = CASES[name]
    noise = noise_sample(length=1 << 15) if case.get('noise') else None
    sim, cfg = make_sim(noise=noise, **case['cfg'])
    inst = c0_like(case['n'], seed=case['seed'], **case['kw'])
    seed = 100 + case['seed']
    out = sim.simulate(inst, seed=seed)
    photons, electrons, sec_rows = presets_from_dumps(sim, inst, seed)
    if case['cfg'].get('enable_electron_afterpulses') or case['cfg'].get('enable_gate_afterpulses'):
        assert len(sec_rows) > 0
    orc = ReplayOracle(cfg, photons, electrons, sec_rows, noise=noise, noise_seed=seed)
    want = orc.simulate(inst, truth_dtype=truth_dtype())
    # Pulse calls and groups the reference scheduler forms vs what the library formed
    g = out['groups']
    got_groups = np.stack([g['left'], g['right'], g['n_intervals']], axis=1) if len(g) else np.zeros((0, 3), np.int64)
    assert np.array_equal(got_groups, np.array(want['groups'], np.int64).reshape(-1, 3))
    he0 = cfg['channel_map']['he'][0]
    rec = want['records']
    assert out['raw_records'].tobytes() == rec[rec['channel'] < he0].tobytes()
    assert out['raw_records_he'].tobytes() == rec[rec['channel'] >= he0].tobytes()
    assert len(out['raw_records_aqmon']) == 0
    # truth rows, in execution order
    runs = []
    for typ, ids, grp in want['runs']:
        n_ph = int(np.isin(photons['id'], ids).sum() - (np.isin(photons['id'], ids) & (photons['ap'] == 1)).sum())
        runs.append((typ, ids, grp, typ in (1, 2) or n_ph > 0))
    check_truth(out['truth'], want['truth'], photons, electrons, runs)
    tr = out['truth']
    assert (tr['n_pe'] >= tr['n_photon']).all() and (tr['n_pe_trigger'] >= tr['n_photon_trigger']).all()
    if name != 'plain':
        assert set(np.unique(tr['type'])) - {1, 2}, 'secondary truth rows expected'
    sim.close()


def test_replay_is_independent_of_device_batching():
    """The same comparison with the run cut into several device batches (and lanes)."""
    from oracle.wfsim_oracle_sim import ReplayOracle
    sim, cfg = make_sim(enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    inst = c0_like(24, seed=12, e_range=(3, 40))
    os.environ['WFS_BATCH_INSTRUCTIONS'] = '8'
    try:
        out = sim.simulate(inst, seed=77)
        assert sim.last_counts['n_batches'] > 1
        photons, electrons, sec_rows = presets_from_dumps(sim, inst, 77)
    finally:
        del os.environ['WFS_BATCH_INSTRUCTIONS']
    want = ReplayOracle(cfg, photons, electrons, sec_rows).simulate(inst, truth_dtype=truth_dtype())
    he0 = cfg['channel_map']['he'][0]
    rec = want['records']
    assert out['raw_records'].tobytes() == rec[rec['channel'] < he0].tobytes()
    g = out['groups']
    assert np.array_equal(np.stack([g['left'], g['right'], g['n_intervals']], axis=1),
                          np.array(want['groups'], np.int64).reshape(-1, 3))
    assert len(out['truth']) == len(want['truth'])
    for n in ('type', 'n_photon', 'n_pe', 'n_pe_trigger', 'n_electron', 't_first_photon', 'endtime'):
        a, b = out['truth'][n], want['truth'][n]
        assert np.array_equal(a[~np.isnan(b.astype(float))], b[~np.isnan(b.astype(float))]), n
    sim.close()
