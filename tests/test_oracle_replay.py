"""CPU: the scheduler / truth rows against the UNMODIFIED reference on identical stage outputs.

tests/golden/sched.npz holds what the reference's own RawData.__call__ / sim_data / get_truth /
digitize_pulse_cache / ZLE / ChunkRawRecords made of preset photons, electrons and secondary instructions
(tests/golden/make_golden_sched.py).  Here the same presets go through
  * oracle.wfsim_oracle_sim.ReplayOracle -- records, truth rows, Pulse calls and digitisation groups must
    be identical (this pins the oracle the GPU tests compare with, tests/test_gpu_replay.py), and
  * the library's host scheduler (wfs_schedule, the function wfs_simulate runs on the device's numbers) --
    instruction -> Pulse call -> digitisation group must be identical.
SURVEY.md rows a4, a5, a29, a34."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_c0_config
from tests.golden.make_golden_sched import CASES, EL_DT, PH_DT
from wfsim_b200.dtypes import instruction_dtype, raw_record_dtype, truth_dtype

IDT = np.dtype(instruction_dtype)
INT64_MIN = np.iinfo(np.int64).min


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'sched.npz'))


def load_case(gold, name):
    cfg = load_c0_config(**json.loads(str(gold[f'{name}_cfg'])))
    gains = cfg['gains'].copy()
    gains[[3, 100, 300]] = 0
    cfg['gains'] = gains
    c = dict(cfg=cfg, prim=gold[f'{name}_prim'].view(IDT), sec=gold[f'{name}_sec'].view(IDT),
             sec_parent=gold[f'{name}_sec_parent'], photons=gold[f'{name}_photons'].view(PH_DT),
             electrons=gold[f'{name}_electrons'].view(EL_DT))
    lens = gold[f'{name}_run_len']
    ids = np.split(gold[f'{name}_run_ids'], np.cumsum(lens)[:-1])
    c['runs'] = list(zip(gold[f'{name}_run_type'].tolist(), ids, gold[f'{name}_run_group'].tolist()))
    c['groups'] = gold[f'{name}_groups']
    c['truth'] = gold[f'{name}_truth'].view(truth_dtype())
    c['rr'] = gold[f'{name}_rr'].view(raw_record_dtype())
    c['rr_he'] = gold[f'{name}_rr_he'].view(raw_record_dtype())
    c['chunks'] = gold[f'{name}_chunks']
    return c


def secondaries_with_ids(c):
    sdt = np.dtype([(n, IDT[n]) for n in IDT.names] + [('_id', np.int64), ('_parent', np.int64)])
    sec = np.zeros(len(c['sec']), sdt)
    for n in IDT.names:
        sec[n] = c['sec'][n]
    sec['_id'] = c['sec']['g4id']
    sec['_parent'] = c['sec_parent']
    return sec


def assert_truth_equal(got, want, float_rtol=0.0):
    assert len(got) == len(want)
    for name in want.dtype.names:
        a, b = got[name], want[name]
        if a.dtype.kind == 'f':
            assert np.array_equal(np.isnan(a), np.isnan(b)), name
            ok = ~np.isnan(b)
            if float_rtol:
                assert np.allclose(a[ok], b[ok], rtol=float_rtol, atol=0), name
            else:
                assert np.array_equal(a[ok], b[ok]), name
        else:
            assert np.array_equal(a, b), name


@pytest.mark.parametrize('name', list(CASES))
def test_replay_oracle_equals_reference_on_preset_stage_outputs(gold, name):
    from oracle.wfsim_oracle_sim import ReplayOracle
    c = load_case(gold, name)
    orc = ReplayOracle(c['cfg'], c['photons'], c['electrons'], secondaries_with_ids(c))
    out = orc.simulate(c['prim'], truth_dtype=truth_dtype(), ids=c['prim']['g4id'])
    # Pulse calls in execution order: type, instruction identities, digitisation group
    assert len(out['runs']) == len(c['runs'])
    for (t0, ids0, g0), (t1, ids1, g1) in zip(out['runs'], c['runs']):
        assert t0 == t1 and g0 == g1 and np.array_equal(ids0, ids1)
    # digitisation groups: (left, right, number of ZLE intervals)
    assert np.array_equal(np.array(out['groups'], np.int64).reshape(-1, 3), c['groups'])
    # records: byte for byte
    he0 = c['cfg']['channel_map']['he'][0]
    rec = out['records']
    assert rec[rec['channel'] < he0].tobytes() == c['rr'].tobytes()
    assert rec[rec['channel'] >= he0].tobytes() == c['rr_he'].tobytes()
    # truth rows in execution order; raw areas (pulse.py:249-250) and the mean / standard deviation of
    # the times (rawdata.py:327,330) are float reductions over the photons in another order
    got = out['truth']
    loose = [n for n in got.dtype.names if n.startswith(('raw_area', 't_mean', 't_sigma'))]
    exact = [n for n in got.dtype.names if n not in loose]
    assert_truth_equal(got[exact], c['truth'][exact])
    assert_truth_equal(got[loose], c['truth'][loose], float_rtol=1e-12)


def pulse_ends(c):
    """max(right) * dt per instruction over its pulses incl. PMT afterpulses (rawdata.py:186-190)."""
    cfg = c['cfg']
    dt = cfg['sample_duration']
    right_margin = int(cfg['samples_to_store_after']) + cfg.get('samples_after_pulse_center', 20)
    n_tot = len(c['prim']) + len(c['sec'])
    end = np.full(n_tot, INT64_MIN, np.int64)
    ph = c['photons']
    if not cfg.get('enable_pmt_afterpulses', True):
        ph = ph[ph['ap'] == 0]
    live = np.asarray(cfg['gains'])[ph['channel']] != 0
    for i in range(n_tot):
        t = ph['t'][live & (ph['id'] == i)]
        if len(t):
            end[i] = (int(t.max()) // dt + right_margin) * dt
    return end


@pytest.mark.parametrize('name', list(CASES))
def test_library_scheduler_equals_reference_on_preset_pulse_ends(gold, name):
    """wfs_schedule (frontend.cu:schedule, host code: runs without a GPU) against the Pulse calls and
    digitisation groups the reference's RawData.__call__ formed."""
    from wfsim_b200 import lib as wlib
    L = wlib.load()
    c = load_case(gold, name)
    cfg, prim, sec = c['cfg'], c['prim'], c['sec']
    assert np.array_equal(prim['g4id'], np.arange(len(prim))) and \
        np.array_equal(sec['g4id'], len(prim) + np.arange(len(sec)))
    end = pulse_ends(c)
    # the secondaries the configuration lets an S2 spawn (rawdata.py:193-201)
    on = ((sec['type'] == 4) & bool(cfg.get('enable_electron_afterpulses', True))) | \
        ((sec['type'] == 6) & bool(cfg.get('enable_gate_afterpulses', False)))
    keep = np.concatenate([np.ones(len(prim), bool), on])
    old_index = np.flatnonzero(keep)
    sec, end = sec[on], end[keep]
    c['sec_parent'] = c['sec_parent'][on]
    n_tot = len(prim) + len(sec)
    run_of = np.zeros(n_tot, np.int32)
    cap = n_tot + 8
    run_type, run_group = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
    n_runs, n_groups = C.c_int64(), C.c_int64()

    def p(a):
        return np.ascontiguousarray(a).ctypes.data
    t, z, ty = (np.ascontiguousarray(prim['time'], np.int64), np.ascontiguousarray(prim['z'], np.float32),
                np.ascontiguousarray(prim['type'], np.int8))
    st, sz, sty = (np.ascontiguousarray(sec['time'], np.int64), np.ascontiguousarray(sec['z'], np.float32),
                   np.ascontiguousarray(sec['type'], np.int8))
    spar = np.ascontiguousarray(c['sec_parent'], np.int32)
    rc = L.wfs_schedule(int(cfg['right_raw_extension']), float(cfg['drift_velocity_liquid']),
                        int(cfg.get('save_full_truth', True)), len(prim), p(t), p(z), p(ty), len(sec), p(st), p(sz),
                        p(sty), p(spar), p(end), p(run_of), p(run_type), p(run_group), cap,
                        C.byref(n_runs), C.byref(n_groups))
    assert rc == 0
    want = c['runs']
    # secondaries whose parent call had no photons are never spawned in the reference (afterpulse.py:24-27)
    # and the presets hold none of those; every instruction is therefore scheduled exactly once
    assert n_runs.value == len(want)
    assert n_groups.value == len(c['groups'])
    for r, (typ, ids, grp) in enumerate(want):
        assert run_type[r] == typ
        assert np.array_equal(np.sort(old_index[np.flatnonzero(run_of == r)]), np.sort(ids)), (r, ids)
        if (end[np.searchsorted(old_index, ids)] != INT64_MIN).any():         # a call without pulses belongs to no group
            assert run_group[r] == grp


def test_add_truth_with_double_photoelectrons_equals_reference(gold):
    """Pulse.add_truth incl. `above_threshold[:n_double_pe]` (pulse.py:229-271), called directly."""
    from oracle.wfsim_oracle_sim import OracleSimulator
    cfg = load_c0_config(**json.loads(str(gold['addtruth_cfg'])))
    gains = cfg['gains'].copy()
    gains[[3, 100, 300]] = 0
    cfg['gains'] = gains
    ph = gold['addtruth_photons'].view(PH_DT)
    orc = OracleSimulator(cfg, spe_table=np.zeros((494, 2001)))
    got = orc.truth_counters(ph['t'], ph['channel'], ph['gain'], ph['dpe'].astype(bool), keep_order=True)
    for k, v in got.items():
        want = float(gold['addtruth_' + k])
        assert v == pytest.approx(want, rel=1e-12), k
        if not k.startswith('raw_area'):
            assert v == want, k


@pytest.mark.parametrize('name', list(CASES))
def test_chunk_bounds_follow_from_the_groups(gold, name):
    """The chunk clock of the plugin mirror on the reference's groups == the reference chunker's bounds."""
    from wfsim_b200.strax_interface import chunk_boundaries
    c = load_case(gold, name)
    got = chunk_boundaries(c['cfg'], int(c['prim']['time'].min()), [tuple(g) for g in c['groups']])
    assert [tuple(x) for x in got] == [tuple(x[:2]) for x in c['chunks'].tolist()]


def test_noise_offset_key_is_the_first_sample_of_the_group():
    """The host Philox (wfsim_b200/philox.py), the oracle's (oracle/philox.py) agree word for word."""
    from oracle.philox import philox4x32 as scalar
    from wfsim_b200.philox import philox4x32 as vec
    idx = np.array([0, 1, 2 ** 32 + 5, 2 ** 63 + 11, (-12345) & 0xffffffffffffffff], np.uint64)
    for seed in (0, 7, 2 ** 40 + 3):
        for stream in (1, 9):
            w = vec(seed, stream, idx, draw=3)
            for k, i in enumerate(idx.tolist()):
                assert [int(x) for x in w[:, k]] == scalar(seed, stream, int(i), 3)


def test_aft_smearing_draw_is_keyed_by_the_instruction():
    """the skew-normal factor of s2.py:660-665 is a function of (seed, rng_id), so pieces and
    shards of a run draw what the whole run draws; its law is scipy.stats.skewnorm's."""
    from scipy import stats
    from wfsim_b200.philox import skewnorm
    ids = np.arange(40000, dtype=np.uint64)
    x = skewnorm(5, ids, 1.0, 0.12, -1.5)
    assert stats.kstest(x, stats.skewnorm(a=-1.5, loc=1.0, scale=0.12).cdf).pvalue > 0.01
    sub = ids[[3, 17, 39999]]
    assert np.array_equal(skewnorm(5, sub, 1.0, 0.12, -1.5), x[[3, 17, 39999]])
    assert not np.array_equal(skewnorm(6, sub, 1.0, 0.12, -1.5), x[[3, 17, 39999]])
