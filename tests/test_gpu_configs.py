"""GPU: BASELINE.json configs [2]-[4] as parity cases at sizes the oracle finishes in seconds, plus
size-independent properties at larger sizes.

  C2  high-energy S2-heavy events with PMT afterpulses + photo-ionisation enabled
  C3  mixed S1/S2 event stream with noise + ZLE
  C4  pulse-superposition microbench: synthetic photons over 494 channels -> wfs_simulate_photons

The stochastic front end cannot be compared record by record with another RNG; what is exact is
the BACK END on the photons the GPU itself generated: the oracle's deterministic path
(bit-exact vs the reference, tests/test_oracle_golden.py) is fed those photons and must return
identical records."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_c0_config
from tests.golden.synth_instructions import c0_like
from tests.golden.synth_tables import EleApHist, pmt_ap_tables
from wfsim_b200.dtypes import instruction_dtype

pytestmark = pytest.mark.gpu
IDT = np.dtype(instruction_dtype)
FIELDS = ('time', 'length', 'dt', 'channel', 'pulse_length', 'record_i', 'baseline')


def make_sim(res_extra=None, **cfg_extra):
    from wfsim_b200.resource import Resource
    from wfsim_b200.simulator import Simulator
    cfg = load_c0_config(**cfg_extra)
    z = np.load(os.path.join(GOLDEN, 'c0_tables.npz'))
    res = Resource(cfg, spe_ppf=z['spe_unique'], spe_row=z['spe_row'][:494], **(res_extra or {}))
    return Simulator(cfg, resource=res), cfg


def check_records_sorted_and_consistent(out, cfg):
    for name in ('raw_records', 'raw_records_he'):
        rr = out[name]
        if not len(rr):
            continue
        key = rr['time'].astype(np.int64) * 1024 + rr['channel']
        assert (np.diff(key) >= 0).all(), name
        assert (rr['dt'] == cfg['sample_duration']).all() and (rr['length'] > 0).all() and (rr['length'] <= 110).all()
        assert (rr['record_i'] == 0).sum() > 0
        # fragments of one pulse: lengths add up to pulse_length
        first = rr[rr['record_i'] == 0]
        assert (first['pulse_length'] >= first['length']).all()
    tr = out['truth']
    assert (tr['n_pe'] >= tr['n_photon']).all()


def heavy_s2_events(n, seed, n_e=(10_000, 40_000)):
    rng = np.random.default_rng(seed)
    inst = c0_like(n, seed=seed)
    s2 = inst['type'] == 2
    inst['amp'][s2] = rng.integers(n_e[0], n_e[1], s2.sum())
    inst['amp'][~s2] = rng.integers(2_000, 20_000, (~s2).sum())
    return inst


def test_c2_high_energy_with_afterpulses():
    res = dict(uniform_to_pmt_ap=pmt_ap_tables(), uniform_to_ele_ap=EleApHist())
    sim, cfg = make_sim(res, enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    inst = heavy_s2_events(3, seed=41)
    out = sim.simulate(inst, seed=4)
    c = sim.last_counts
    assert c['n_photons'] > 1_000_000            # ~1e4-4e4 electrons x ~20-30 photons
    check_records_sorted_and_consistent(out, cfg)
    tr = out['truth']
    assert (tr['type'] == 4).sum() > 0           # photo-ionisation trains were simulated
    s2 = tr[tr['type'] == 2]
    # electron survival: Binomial(amp, extraction * exp(-t/tau)) (s2.py:212-256)
    z = s2['z'].astype(np.float64)
    p = cfg['electron_extraction_yield'] * np.exp(-(-z / cfg['drift_velocity_liquid'] + cfg['drift_time_gate'])
                                                  / cfg['electron_lifetime_liquid'])
    assert np.all(np.abs(s2['n_electron'] - s2['amp'] * p) < 6 * np.sqrt(s2['amp'] * p * (1 - p)) + 5)
    # reproducible: same seed, same bytes
    out2 = sim.simulate(inst, seed=4)
    assert out2['raw_records'].tobytes() == out['raw_records'].tobytes()
    assert out2['truth'].tobytes() == out['truth'].tobytes()
    sim.close()


def test_c2_back_end_exact_on_generated_photons():
    """Dense regime (thousands of photons per channel and pulse): the GPU back end on the photons
    the GPU generated == the oracle's deterministic path on the same photons."""
    from oracle import wfsim_oracle as orc
    sim, cfg = make_sim()
    inst = heavy_s2_events(1, seed=42, n_e=(6_000, 8_000))
    ph = sim.sample_stage(inst, stage=0, seed=5)
    out = sim.simulate(inst, seed=5)
    # Pulse calls: one per instruction (save_full_truth); one digitisation group per cluster
    pcall = ph['instruction'].astype(np.int32)
    n_pc = len(inst)
    stime = inst['time'] + (inst['z'] / np.float32(cfg['drift_velocity_liquid'])).astype(np.int64) * ((inst['type'] % 2) - 1)
    order = np.argsort(stime, kind='stable')
    gaps = np.diff(stime[order]) > cfg['right_raw_extension']
    # groups as the scheduler built them: taken from the result (left/right), not re-derived
    groups = out['groups']
    group_of = np.zeros(n_pc, np.int32)
    if len(groups) > 1:
        # assign every pulse call to the group whose time range holds its first photon
        for i in range(n_pc):
            m = ph['instruction'] == i
            if m.any():
                s = ph['t'][m].min() // cfg['sample_duration']
                group_of[i] = int(np.argmax((groups['left'] <= s) & (s <= groups['right'])))
    want = orc.simulate_photons(cfg, pcall, ph['channel'].astype(np.int32), ph['t'].astype(np.int64),
                                ph['gain'].astype(np.float64), group_of)
    got, exp = out['raw_records'], want['raw_records']
    assert len(got) == len(exp)
    for f in FIELDS:
        np.testing.assert_array_equal(got[f], exp[f], err_msg=f)
    np.testing.assert_array_equal(got['data'], exp['data'])
    sim.close()


def test_c3_mixed_stream_with_noise():
    rng = np.random.default_rng(3)
    noise = np.round(rng.normal(0, 1.6, (4096, 494)) * 2) / 2
    sim, cfg = make_sim(dict(noise_data=noise), enable_noise=True)
    inst = c0_like(40, seed=43)
    out = sim.simulate(inst, seed=6)
    check_records_sorted_and_consistent(out, cfg)
    rr = out['raw_records']
    # noise is added everywhere: baseline samples are no longer constant
    assert rr['data'][rr['length'] == 110].std() > 0.5
    # the HE rows carry baseline + noise only (int(0.05) == 0) and stay below threshold: no records
    assert len(out['raw_records_he']) == 0
    # chunk-independent: two halves simulated separately give the same records (events are
    # independent; the noise offset is a Philox draw per digitisation group index)
    half = len(inst) // 2
    a = sim.simulate(inst[:half], seed=6, rng_id=np.arange(half))
    assert a['raw_records'].tobytes() == rr[rr['time'] < inst['time'][half] - 1_000_000].tobytes()
    sim.close()


def test_c4_superposition_microbench_properties():
    """C4 at 2e6 photons: (a) permutation invariance, (b) checksum of the ADC deficit equals the
    sum over photons of the rounded template area within rounding bounds, (c) idempotence."""
    from wfsim_b200.simulator import Simulator
    cfg = load_c0_config()
    sim = Simulator(cfg, resource=None)
    rng = np.random.default_rng(44)
    n, n_groups = 2_000_000, 400
    grp = rng.integers(0, n_groups, n)
    t = 1_000_000_000 + grp.astype(np.int64) * 3_000_000 + rng.integers(0, 20_000, n)
    ch = rng.integers(0, 494, n).astype(np.int32)
    g = cfg['gains'][ch] * (0.5 + rng.random(n))
    pcall = grp.astype(np.int32)
    group_of = np.arange(n_groups, dtype=np.int32)
    a = sim.simulate_photons(t, ch, g, pcall, group_of)
    perm = rng.permutation(n)
    b = sim.simulate_photons(t[perm], ch[perm], g[perm], pcall[perm], group_of)
    assert a['raw_records'].tobytes() == b['raw_records'].tobytes()
    rr = a['raw_records']
    deficit = (16000 - rr['data'].astype(np.int64))
    valid = np.arange(110)[None, :] < rr['length'][:, None]
    total = int(deficit[valid].sum())
    # every photon deposits gain * sum(template) * current_2_adc ADC counts; one rounding per
    # (pulse, sample): |error| <= 0.5 per touched sample
    z = np.load(os.path.join(GOLDEN, 'c0_tables.npz'))
    area = float((g * z['templates'].sum(axis=1)[t % 10]).sum() * float(z['current_2_adc']))
    touched = int((deficit[valid] != 0).sum())
    assert abs(total - area) <= 0.5 * touched + 1
    assert abs(total - area) / area < 2e-3
    sim.close()


def test_group_local_order_with_afterpulses_and_secondaries(monkeypatch):
    """With PMT afterpulses and photo-ionisation electrons a group's photons sit in up to four runs of
    the photon array (primaries, secondaries, afterpulse children of either); the per-group
    shared-memory sort gathers them and gives the bytes of the device-wide radix sort."""
    from tests.golden.synth_instructions import c1_like
    sim, cfg = make_sim(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                        enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    inst = c1_like(400, seed=19)
    outs = {}
    monkeypatch.setenv('WFS_FUSED', '0')       # the photon order of the multi-pass back end is what is compared
    for mode in ('0', '1'):
        monkeypatch.setenv('WFS_SEGMENT_SORT', mode)
        o = sim.simulate(inst, seed=44)
        outs[mode] = {k: np.array(v) for k, v in o.items() if k != '_pinned'}
        seg_batches = sim.last_counts['ms_phase'][10]
        assert (seg_batches > 0) == (mode == '1')
    for k in ('raw_records', 'raw_records_he', 'truth', 'groups'):
        assert outs['0'][k].tobytes() == outs['1'][k].tobytes(), k
    check_records_sorted_and_consistent(outs['1'], cfg)
    ph = sim.sample_stage(inst, stage=0, seed=44)
    assert ((ph['flags'] >> 1) & 1).sum() > 100          # afterpulse photons are there
    sim.close()


def test_c1_large_run_properties(monkeypatch):
    """BASELINE config [1] at 2e4 events (a fifth of the bench size; 2.7e7 photons, 2.5e7 records, 6 GB):
    size-independent properties instead of an oracle run -- every data type sorted by (time, channel),
    fragments consistent, truth photon sum equal to the photons superposed, the same bytes from a
    second run, and the same bytes when the device batches are cut differently."""
    import zlib
    from tests.golden.synth_instructions import c1_like
    sim, cfg = make_sim()
    inst = c1_like(20_000, seed=100)

    def run():
        out = sim.simulate(inst, seed=1)
        c = dict(sim.last_counts)
        rr = out['raw_records']
        key = rr['time'] * 1024 + rr['channel']
        assert (np.diff(key) >= 0).all()
        assert (rr['length'] > 0).all() and (rr['length'] <= 110).all() and (rr['dt'] == cfg['sample_duration']).all()
        first = rr[rr['record_i'] == 0]
        assert (first['pulse_length'] >= first['length']).all()
        # a record is full unless it is the last fragment of its pulse
        assert ((rr['length'] == 110) | (rr['pulse_length'] - 110 * rr['record_i'].astype(np.int64) == rr['length'])).all()
        assert out['truth']['n_photon'].sum() == c['n_photons']
        assert len(out['truth']) == len(inst)
        crc = zlib.crc32(rr.view(np.uint8)[:len(rr) * 244:1].tobytes()) if len(rr) < 4_000_000 else \
            zlib.crc32(np.ascontiguousarray(rr[::17]).tobytes())
        return crc, len(rr), int(rr['data'].astype(np.int64)[::101].sum()), c['n_batches']

    a = run()
    b = run()
    assert a == b
    monkeypatch.setenv('WFS_BATCH_SAMPLES', '300000000')
    c = run()
    assert c[3] > a[3] and c[:3] == a[:3]
    assert a[1] > 2e7
    sim.close()


def test_forced_batch_cuts_keep_the_records_sorted(monkeypatch):
    """A dense stream with photo-ionisation secondaries and a batch budget of three instructions: cuts are
    forced where no quiet gap exists, delayed secondaries of one batch reach behind the start of the next,
    and simulate() hands the records out in (time, channel) order all the same."""
    sim, cfg = make_sim(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                        enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    inst = c0_like(60, seed=23, event_rate=3000.0, e_range=(1, 8))      # one event every 333 us: clusters, but no quiet gap
    monkeypatch.setenv('WFS_BATCH_INSTRUCTIONS', '3')
    out = sim.simulate(inst, seed=2)
    assert sim.last_counts['n_batches'] > 5
    check_records_sorted_and_consistent(out, cfg)
    monkeypatch.delenv('WFS_BATCH_INSTRUCTIONS')
    one = sim.simulate(inst, seed=2)
    assert sim.last_counts['n_batches'] == 1
    check_records_sorted_and_consistent(one, cfg)
    sim.close()
