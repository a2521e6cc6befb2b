"""GPU: the group-resident fused back end (csrc/fused.cu: one CTA per digitisation group, photons ->
records) against the multi-pass back end (csrc/backend.cu) on the same photons: records, truth rows,
group bookkeeping and counters must be identical byte for byte.  The multi-pass path is itself bit-exact
against the reference's records (tests/test_gpu_deterministic.py, golden det_*.npz), and both are
compared with the oracle on the GPU's photons in tests/test_gpu_replay.py."""
import os

import numpy as np
import pytest

from tests.golden.synth_instructions import c0_like, c1_like
from tests.test_gpu_afterpulse_plugin import make_sim

pytestmark = pytest.mark.gpu


class env:
    def __init__(self, **kv):
        self.kv = kv

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update({k: str(v) for k, v in self.kv.items()})

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


def both_paths(sim, inst, seed, **kw):
    with env(WFS_FUSED=1):
        a = sim.simulate(inst, seed=seed, **kw)
        ca = dict(sim.last_counts)
    with env(WFS_FUSED=0):
        b = sim.simulate(inst, seed=seed, **kw)
        cb = dict(sim.last_counts)
    assert ca['n_fused_batches'] == ca['n_batches'] > 0, 'the fused kernel did not run'
    assert cb['n_fused_batches'] == 0
    return a, b, ca, cb


def assert_same(a, b, ca, cb):
    for k in ('raw_records', 'raw_records_he', 'raw_records_aqmon', 'truth', 'groups'):
        assert len(a[k]) == len(b[k]), k
        assert a[k].tobytes() == b[k].tobytes(), k
    for k in ('n_records', 'n_records_total', 'n_truth', 'n_photons', 'n_pe', 'n_pulses', 'n_windows', 'n_intervals',
              'n_samples', 'n_groups', 'n_pulse_calls'):
        assert ca[k] == cb[k], k
    assert len(a['raw_records']) > 0


# energies that keep every digitisation group within the photons one CTA holds (kFusedMaxPhotons)
CONFIGS = {
    'c0': (dict(), lambda: c0_like(14, seed=3, e_range=(1, 12))),
    'c1_low_energy': (dict(), lambda: c1_like(300, seed=5)),
    'afterpulses': (dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True), lambda: c0_like(12, seed=4, e_range=(2, 10))),
    'pile_up_merged': (dict(enable_pmt_afterpulses=True, save_full_truth=False),
                       lambda: c0_like(30, seed=6, event_rate=6000.0, e_range=(0.5, 3))),
    'pile_up_full_truth': (dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True),
                           lambda: c0_like(24, seed=16, event_rate=9000.0, e_range=(0.5, 3))),
    'thresholds': (dict(zle_threshold=40, special_thresholds={'7': 60, '255': 5, '300': 25}), lambda: c0_like(10, seed=7, e_range=(1, 12))),
    'gate': (dict(enable_gate_afterpulses=True, photoelectric_p=0.004), lambda: c0_like(10, seed=8, e_range=(1, 12))),
    # many groups per launch (CTAs take several groups in turn) with channels that hold several pulse calls
    'afterpulses_many_groups': (dict(enable_pmt_afterpulses=True), lambda: c1_like(400, seed=9)),
    'bright_s1': (dict(), lambda: c0_like(10, seed=21, e_range=(8, 12), s1_per_kev=400.0, s2_per_kev=2.0)),
}


@pytest.mark.parametrize('name', list(CONFIGS))
def test_fused_back_end_equals_multi_pass_back_end(name):
    extra, make = CONFIGS[name]
    sim, cfg = make_sim(**extra)
    a, b, ca, cb = both_paths(sim, make(), seed=31)
    assert_same(a, b, ca, cb)
    sim.close()


def test_fused_with_dead_pmts_and_per_pmt_truth():
    from wfsim_b200.resource import Resource
    from wfsim_b200.simulator import Simulator
    from tests.conftest import load_c0_config
    from tests.test_gpu_afterpulse_plugin import spe
    cfg = load_c0_config(per_pmt_truth=True)
    gains = cfg['gains'].copy()
    gains[[3, 100, 300, 493]] = 0
    cfg['gains'] = gains
    uniq, row = spe()
    sim = Simulator(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row))
    a, b, ca, cb = both_paths(sim, c0_like(8, seed=9, e_range=(1, 12)), seed=5, per_pmt_truth=True)
    assert_same(a, b, ca, cb)
    assert a['truth']['n_photon_per_pmt'].sum() == a['truth']['n_photon'].sum()
    assert not np.isin(a['raw_records']['channel'], [3, 100, 300, 493]).any()
    sim.close()


def test_fused_with_few_live_pmts_long_channel_lists():
    """All but 40 PMTs switched off: the photons of a group crowd on a few channels -- channel lists longer than the
    rank sort takes (a warp's bitonic network orders them), piled-up channels evaluated by a warp, PMT afterpulses on
    top of the signal (photons whose samples another pulse call reaches), and hundreds of records in one time bin of
    the record order."""
    from wfsim_b200.resource import Resource
    from wfsim_b200.simulator import Simulator
    from tests.conftest import load_c0_config
    from tests.golden.synth_tables import pmt_ap_tables
    from tests.test_gpu_afterpulse_plugin import spe
    cfg = load_c0_config(enable_pmt_afterpulses=True)
    gains = cfg['gains'].copy()
    live = np.r_[np.arange(0, 250, 10), np.arange(260, 490, 16)]
    dead = np.setdiff1d(np.arange(len(gains)), live)
    gains[dead] = 0
    cfg['gains'] = gains
    uniq, row = spe()
    sim = Simulator(cfg, resource=Resource(cfg, spe_ppf=uniq, spe_row=row, uniform_to_pmt_ap=pmt_ap_tables()))
    inst = c0_like(12, seed=13, e_range=(4, 12), s1_per_kev=120.0)
    a, b, ca, cb = both_paths(sim, inst, seed=17)
    assert_same(a, b, ca, cb)
    per_channel = np.bincount(a['raw_records']['channel'], minlength=len(gains))
    assert per_channel[dead].sum() == 0 and per_channel[live].min() > 0
    assert ca['n_photons'] / max(ca['n_groups'], 1) / len(live) > 20, 'the channel lists are not long'
    sim.close()


def test_fused_over_several_batches_lanes_and_transports():
    """Several device batches on several lanes; compact transport (pageable destination) against the plain
    DMA (pinned destination); a record buffer that is too small is grown and the call repeated."""
    sim, cfg = make_sim(enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    inst = c0_like(40, seed=11, e_range=(1, 6))
    with env(WFS_FUSED=1):
        ref = sim.simulate(inst, seed=2)
        with env(WFS_BATCH_INSTRUCTIONS=10):
            cut = sim.simulate(inst, seed=2)
            assert sim.last_counts['n_batches'] > 2 and sim.last_counts['n_fused_batches'] == sim.last_counts['n_batches']
            pinned = sim.simulate(inst, seed=2, pinned=True)
            small = sim.simulate(inst, seed=2, cap_records=100)
            with env(WFS_FUSED_REC_CAP=64):       # the groups outgrow the record lists: second attempt with large ones
                retried = sim.simulate(inst, seed=2)
                assert sim.last_counts['n_fused_batches'] == sim.last_counts['n_batches']
    for other in (cut, pinned, small, retried):
        for k in ('raw_records', 'raw_records_he', 'truth', 'groups'):
            assert other[k].tobytes() == ref[k].tobytes(), k
    sim.close()


@pytest.mark.parametrize('fraction', ['adaptive', '0', '0.3', '1'])
def test_split_transport_into_page_locked_destinations(fraction):
    """A destination the caller page-locked (Simulator.pin -> wfs_host_register) receives part of every batch as
    plain rows by DMA and the rest compact + expanded; the records do not depend on the share."""
    from wfsim_b200.dtypes import raw_record_dtype
    sim, cfg = make_sim(enable_pmt_afterpulses=True)
    inst = c1_like(400, seed=9)
    ref = sim.simulate(inst, seed=5)
    n = len(ref['raw_records'])
    dest = np.empty(n + 5000, raw_record_dtype())
    dest.view(np.uint8)[:] = 0xAB
    assert sim.pin(dest)
    kv = dict(WFS_BATCH_INSTRUCTIONS=100)
    if fraction != 'adaptive':
        kv['WFS_PLAIN_FRACTION'] = fraction
    else:
        kv.update(WFS_PLAIN_ADAPTIVE=1, WFS_EXPAND_THREADS=2)
    with env(**kv):
        for rep in range(3):          # the adaptive share moves from call to call
            out = sim.simulate(inst, seed=5, records_out=dest)
            c = sim.last_counts
            assert c['n_batches'] > 2 and c['n_fused_batches'] == c['n_batches']
            if fraction == '0':
                assert c['n_plain_records'] == 0
            elif fraction == '1':
                assert c['n_plain_records'] >= n - 4 * c['n_batches']
            elif fraction == '0.3':
                assert 0 < c['n_plain_records'] < n
            assert out['raw_records'].tobytes() == ref['raw_records'].tobytes()
            assert out['truth'].tobytes() == ref['truth'].tobytes()
    del out, dest
    sim.close()


def test_repeated_group_with_many_pulse_calls_counts_its_triggers_once():
    """A pile-up group with more Pulse calls than shared-memory trigger counters (they then go straight to HBM) that
    also outgrows the record list of its size class and is repeated with the largest lists: n_pe_trigger must not
    be counted twice.  (Found by profiles/tools/fuzz_fused.py.)"""
    sim, cfg = make_sim()
    inst = c0_like(20, seed=16, event_rate=50000.0, e_range=(0.2, 0.5))      # 40 Pulse calls within 0.4 ms
    with env(WFS_FUSED_REC_CAP=64):
        a, b, ca, cb = both_paths(sim, inst, 5)
    g = a['groups']
    per_group = [int(((inst['time'] >= l * 10 - 1_000_000) & (inst['time'] <= r * 10)).sum()) for l, r in zip(g['left'], g['right'])]
    assert max(per_group) > 32, 'the events were meant to pile up: more than 32 Pulse calls in one group'
    assert_same(a, b, ca, cb)
    assert (a['truth']['n_pe_trigger'] > a['truth']['n_photon_trigger']).any()
    sim.close()


def test_overlapping_groups_take_the_multi_pass_path():
    """Delayed secondaries can form a digitisation group that starts before the previous one has ended; the records
    of the two interleave in time, which the group-ordered fused back end cannot deliver: such a batch is detected
    (k_groups_disjoint) and handed to the multi-pass back end.  (Found by profiles/tools/fuzz_simulate.py, seed 12,
    iteration 77.)"""
    from tests.golden.synth_tables import EleApHist, pmt_ap_tables
    from tests.test_gpu_configs import make_sim as make_sim_res
    from wfsim_b200.dtypes import instruction_dtype
    sim, cfg = make_sim_res(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                            enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
    found = False
    for seed in range(40):
        rng = np.random.default_rng(1000 + seed)
        n = 31
        inst = np.zeros(n, instruction_dtype)
        inst['type'] = rng.choice([1, 2], n)
        inst['time'] = np.sort(rng.integers(0, 70000, n))
        r = np.sqrt(rng.uniform(0, 55 ** 2, n)); th = rng.uniform(-np.pi, np.pi, n)
        inst['x'], inst['y'] = r * np.cos(th), r * np.sin(th)
        inst['z'] = rng.uniform(-110, 0, n)
        inst['amp'] = (10 ** rng.uniform(0, 3.7, n)).astype(int)
        inst['recoil'], inst['local_field'], inst['event_number'] = 7, 82.0, np.arange(n)
        inst = inst[inst['amp'] > 0]
        with env(WFS_FUSED=1):
            a = sim.simulate(inst, seed=seed)
            ca = dict(sim.last_counts)
        g = a['groups']
        rr = a['raw_records']
        key = rr['time'].astype(np.int64) * 1024 + rr['channel']
        assert (np.diff(key) >= 0).all(), seed
        with_records = g[g['n_intervals'] > 0]
        overlap = len(with_records) > 1 and (with_records['left'][1:] <= with_records['right'][:-1]).any()
        if overlap:
            found = True
            assert ca['n_fused_batches'] == 0
            with env(WFS_FUSED=0):
                b = sim.simulate(inst, seed=seed)
            assert a['raw_records'].tobytes() == b['raw_records'].tobytes()
            break
    assert found, 'no case with overlapping groups among the seeds'
    sim.close()


def test_groups_that_do_not_fit_take_the_multi_pass_path():
    """A heavy S2 (more photons than a CTA holds) sends its batch through the multi-pass back end; the
    light batches of the same call still use the fused kernel; the result does not depend on that."""
    sim, cfg = make_sim()
    inst = c0_like(12, seed=13, e_range=(2, 20))
    inst['amp'][inst['type'] == 2][:1] = 1
    heavy = np.flatnonzero(inst['type'] == 2)[5]
    inst['amp'][heavy] = 4000                      # ~ 70 000 photons in one group
    with env(WFS_BATCH_INSTRUCTIONS=4):
        with env(WFS_FUSED=1):
            a = sim.simulate(inst, seed=3)
            ca = dict(sim.last_counts)
        with env(WFS_FUSED=0):
            b = sim.simulate(inst, seed=3)
    assert 0 < ca['n_fused_batches'] < ca['n_batches']
    for k in ('raw_records', 'truth', 'groups'):
        assert a[k].tobytes() == b[k].tobytes(), k
    sim.close()
