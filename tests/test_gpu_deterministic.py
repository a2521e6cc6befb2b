"""GPU parity of the deterministic back end (through the C-ABI) against
(a) the golden vectors produced by the unmodified reference and (b) the CPU oracle on fresh
seeded inputs.  Bit-exact: integer/byte work (ADC counts, fragment boundaries, headers)."""
import numpy as np
import pytest

from tests.conftest import DET_CASES, load_c0_config, load_det_case

pytestmark = pytest.mark.gpu

FIELDS = ('time', 'length', 'dt', 'channel', 'pulse_length', 'record_i', 'baseline')


def assert_records_equal(got, want, what=''):
    assert len(got) == len(want), f'{what}: {len(got)} records vs {len(want)}'
    for f in FIELDS:
        np.testing.assert_array_equal(got[f], want[f], err_msg=f'{what}:{f}')
    np.testing.assert_array_equal(got['data'], want['data'], err_msg=f'{what}:data')


def make_sim(cfg, noise=None):
    from wfsim_b200.simulator import Simulator
    res = {'noise_data': noise} if noise is not None else None
    return Simulator(cfg, resource=res)


@pytest.mark.parametrize('name', DET_CASES)
def test_matches_reference_golden(name):
    c = load_det_case(name)
    sim = make_sim(c['cfg'], c['noise'])
    out = sim.simulate_photons(c['t'], c['channel'], c['gain'], c['pcall'], c['group_of'],
                               ix_rand=c['ix_rand'] if c['noise'] is not None else None)
    assert_records_equal(out['raw_records'], c['rr'], name + ' tpc')
    assert_records_equal(out['raw_records_he'], c['rr_he'], name + ' he')
    assert len(out['raw_records_aqmon']) == 0          # row 800 is never emitted (rawdata.py:250-254)
    assert sim.last_counts['gpu_launches'] > 0
    sim.close()


@pytest.mark.parametrize('seed,n_groups,big', [(101, 12, False), (102, 4, True)])
def test_matches_oracle_fresh_inputs(seed, n_groups, big):
    from oracle import wfsim_oracle as orc
    from tests.golden.synth import synth_photons
    cfg = load_c0_config()
    gains = cfg['gains'].copy()
    gains[[5, 77, 400]] = 0
    cfg['gains'] = gains
    rng = np.random.default_rng(seed)
    pcall, ch, t, g, group_of = synth_photons(cfg, rng, n_groups, big=big)
    # shuffle: the entry accepts photons in any order
    perm = rng.permutation(len(t))
    want = orc.simulate_photons(cfg, pcall, ch, t, g, group_of)
    sim = make_sim(cfg)
    out = sim.simulate_photons(t[perm], ch[perm], g[perm], pcall[perm], group_of)
    assert_records_equal(out['raw_records'], want['raw_records'], 'tpc')
    assert_records_equal(out['raw_records_he'], want['raw_records_he'], 'he')
    for gi, lr in zip(out['groups'], want['groups_lr']):
        assert (gi['left'], gi['right']) == lr
    sim.close()


def test_edge_cases():
    cfg = load_c0_config()
    gains = cfg['gains'].copy()
    gains[9] = 0
    cfg['gains'] = gains
    sim = make_sim(cfg)
    z = np.zeros(0)
    out = sim.simulate_photons(z, z, z, z, np.zeros(0, np.int32))
    assert len(out['raw_records']) == 0
    # photons only on a dead PMT / invalid channel: nothing is produced
    out = sim.simulate_photons([1000, 1010], [9, -1], [1e6, 1e6], [0, 0], [0])
    assert len(out['raw_records']) == 0 and out['groups']['n_intervals'][0] == -1
    # a single photon: one pulse of 52 + 70 + 1 samples, window +-50, one interval
    g = float(gains[0])
    out = sim.simulate_photons([1_000_005], [300], [g], [0], [0])
    rr = out['raw_records']
    assert len(rr) >= 1 and rr['channel'][0] == 300 and rr['record_i'][0] == 0
    assert rr['data'][0].min() < 16000 - 15
    assert np.all(np.diff(rr['time']) >= 0)
    sim.close()


def test_linearity_and_idempotence():
    """Size-independent properties: the same input gives the same bytes twice; two disjoint
    groups simulated together equal the two simulated separately (concatenated)."""
    from tests.golden.synth import synth_photons
    cfg = load_c0_config()
    rng = np.random.default_rng(7)
    pcall, ch, t, g, group_of = synth_photons(cfg, rng, 6)
    sim = make_sim(cfg)
    a = sim.simulate_photons(t, ch, g, pcall, group_of)['raw_records']
    b = sim.simulate_photons(t, ch, g, pcall, group_of)['raw_records']
    assert a.tobytes() == b.tobytes()
    parts = []
    for grp in range(int(group_of.max()) + 1):
        pcs = np.flatnonzero(group_of == grp)
        m = np.isin(pcall, pcs)
        remap = np.full(len(group_of), -1, np.int32)
        remap[pcs] = np.arange(len(pcs))
        parts.append(sim.simulate_photons(t[m], ch[m], g[m], remap[pcall[m]],
                                          np.zeros(len(pcs), np.int32))['raw_records'])
    cat = np.concatenate(parts)
    assert cat.tobytes() == a.tobytes()
    sim.close()


@pytest.mark.parametrize('n_calls,n_ch_used', [(6, 3), (40, 5)])
def test_many_overlapping_pulse_calls(n_calls, n_ch_used):
    """Several Pulse calls of one digitisation group overlapping on the same samples of the same
    channels (pile-up, G4-style inputs): every call is rounded on its own and the integers add
    (rawdata.py:236-239).  Exercises the layer-collision and >= 31-calls paths of k_digitize."""
    from oracle import wfsim_oracle as orc
    cfg = load_c0_config()
    gains = cfg['gains']
    rng = np.random.default_rng(300 + n_calls)
    pcall, ch, t, g = [], [], [], []
    for pc in range(n_calls):
        n = int(rng.integers(3, 30))
        pcall.append(np.full(n, pc))
        c = rng.integers(0, n_ch_used, n) * 50 + 3
        ch.append(c)
        t.append(2_000_000 + rng.integers(0, 1500, n))
        g.append(gains[c] * (0.3 + rng.exponential(0.7, n)))
    pcall, ch, t, g = (np.concatenate(x) for x in (pcall, ch, t, g))
    group_of = np.zeros(n_calls, np.int32)
    want = orc.simulate_photons(cfg, pcall.astype(np.int32), ch.astype(np.int32), t.astype(np.int64), g, group_of)
    sim = make_sim(cfg)
    out = sim.simulate_photons(t, ch, g, pcall, group_of)
    assert_records_equal(out['raw_records'], want['raw_records'], 'tpc')
    sim.close()


@pytest.mark.parametrize('noise', [False, True])
def test_compact_transport_equals_plain_copy(noise, monkeypatch):
    """Records that cross PCIe in the compact form (k_pack<true> + host expansion, transport.cuh)
    are byte-identical to the plain 244-byte copy -- with a constant baseline (few blocks in the
    stream) and with noise (every block in the stream)."""
    from tests.golden.synth import synth_photons
    cfg = load_c0_config()
    noise_data = None
    if noise:
        cfg['enable_noise'] = True
        noise_data = np.round(np.random.default_rng(3).normal(0, 2, (4096, 494)))
    rng = np.random.default_rng(11)
    pcall, ch, t, g, group_of = synth_photons(cfg, rng, 10, big=True)
    ix = np.arange(int(group_of.max()) + 1, dtype=np.int64) * 7
    outs = {}
    for mode in ('0', '2'):
        monkeypatch.setenv('WFS_COMPACT', mode)
        sim = make_sim(cfg, noise_data)
        outs[mode] = sim.simulate_photons(t, ch, g, pcall, group_of, ix_rand=ix if noise else None)
        d2h = sim.last_counts['d2h_bytes']
        n = sim.last_counts['n_records_total']
        assert n > 0
        if mode == '0':
            assert d2h == 244 * n
        elif not noise:
            assert d2h < 0.8 * 244 * n
        sim.close()
    for k in ('raw_records', 'raw_records_he'):
        assert outs['0'][k].tobytes() == outs['2'][k].tobytes()
