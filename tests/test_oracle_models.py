"""CPU: the oracle's restatements of the optional S2 model rows (oracle/wfsim_oracle_models.py) against
samples drawn by the unmodified reference (tests/golden/stoch_lumw.npz, stoch_diffuse.npz; generating
scripts tests/golden/make_golden_lumw.py, make_golden_diffuse.py).  p > 0.01 throughout."""
import os

import numpy as np

from oracle import wfsim_oracle_models as OM
from tests.conftest import GOLDEN, load_c0_config
from tests.golden import synth_maps as SM
from tests.stat_helpers import P_MIN, ks_p


def jitter(a, seed=0):
    return np.asarray(a, float) + np.random.default_rng(seed).random(len(a))


def test_warped_gas_gap_luminescence_matches_reference():
    from tests.golden.make_golden_lumw import POSITIONS
    gold = np.load(os.path.join(GOLDEN, 'stoch_lumw.npz'))
    cfg = load_c0_config(enable_gas_gap_warping=True)
    gap_of = SM.GasGapLength()
    rng = np.random.default_rng(7)
    n = 60_000
    for name, xy in POSITIONS.items():
        t = OM.luminescence_timings_warped(gap_of(np.array([xy])), [n], cfg, rng)
        assert ks_p(jitter(t), jitter(gold['lumw_' + name], 1)) > P_MIN, name
    # two instructions in one call: the grid starts at the larger gap, which shifts the smaller one's times
    gaps = gap_of(np.array([POSITIONS['a'], POSITIONS['c']]))
    t = OM.luminescence_timings_warped(gaps, [n, n], cfg, rng)
    assert ks_p(jitter(t[:n]), jitter(gold['lumw_a_with_c'], 1)) > P_MIN
    assert ks_p(jitter(t[n:]), jitter(gold['lumw_c_with_a'], 1)) > P_MIN
    assert ks_p(jitter(t[:n]), jitter(gold['lumw_a'], 1)) < 1e-6
    # the product's host-side field scale is the oracle's
    from wfsim_b200 import tables
    assert np.allclose(tables.luminescence_field_scale(cfg, gaps), OM.field_scale(cfg, gaps), rtol=1e-14)


def test_transverse_diffusion_pattern_matches_reference():
    from tests.golden.make_golden_diffuse import CASES, N_ELECTRON, N_ELECTRON_FEW, normalised
    gold = np.load(os.path.join(GOLDEN, 'stoch_diffuse.npz'))
    efd = dict(survival_probability_map=False, drift_speed_map=True, diffusion_longitudinal_map=False,
               diffusion_transverse_map=True)
    cfg = load_c0_config(enable_field_dependencies=efd, diffusion_constant_transverse=1.0)
    fd = SM.FieldDependencies()
    grid = SM.s2_pattern_grid(int(cfg['n_top_pmts']))
    rng = np.random.default_rng(11)
    for name, (x, y, z) in CASES.items():
        sr, sa = OM.hdiff_sigmas(cfg, np.array([z]), np.array([[x, y]]), fd.field_dependencies_map)
        if name != 'edge':           # displacement samples of the reference (untruncated cases)
            assert ks_p(rng.normal(0, sr[0], 10000), gold[f'diff_{name}_radial']) > P_MIN
            assert ks_p(rng.normal(0, sa[0], 10000), gold[f'diff_{name}_azimuthal']) > P_MIN
        for tag, n_i, n_e in (('', 1500, N_ELECTRON), ('_few', 3000, N_ELECTRON_FEW)):
            pat = OM.s2_pattern_diffuse(np.full(n_i, n_e), np.tile([[x, y]], (n_i, 1)), np.full(n_i, sr[0]),
                                        np.full(n_i, sa[0]), grid, cfg['tpc_radius'], rng)
            ok = ~np.isnan(pat).any(axis=1)
            nan_rows, n_gold = gold[f'diff_{name}{tag}_nan_rows']
            f, f_gold = (~ok).mean(), nan_rows / n_gold
            assert abs(f - f_gold) < 5 * np.sqrt(max(f_gold, 1e-3) * (1 / n_i + 1 / n_gold)), (name, tag)
            p = normalised(pat[ok], cfg)
            k = int(gold[f'diff_{name}{tag}_peak'][0])
            ref = gold[f'diff_{name}{tag}_peak_p']
            # (the golden's channel was picked for its large fluctuation in that very sample, so its mean
            # there is biased upwards by about one standard error: compare the spread, and all means below)
            assert abs(p[:, k].std() - ref.std()) < 0.12 * ref.std(), (name, tag)
            assert abs(p[:, k].mean() - ref.mean()) < 5 * np.sqrt(ref.var() / len(ref) + p[:, k].var() / len(p))
            # the whole mean pattern: every channel within 5 standard errors of the reference's mean
            se = np.sqrt(p.var(axis=0) / len(p) + p.var(axis=0) / n_gold)
            assert (np.abs(p.mean(axis=0) - gold[f'diff_{name}{tag}_p']) <= 5 * se + 1e-12).all(), (name, tag)


def test_timing_model_rows_match_reference():
    """S1 optical propagation / custom models, garfield luminescence (+ confined, + optical propagation),
    garfield gas-gap luminescence: oracle restatements against tests/golden/stoch_models.npz, stoch_gg.npz."""
    from tests.golden.make_golden_gg import POSITIONS as GG_POS
    from tests.stat_helpers import discrete_p
    gm = np.load(os.path.join(GOLDEN, 'stoch_models.npz'))
    gg = np.load(os.path.join(GOLDEN, 'stoch_gg.npz'))
    rng = np.random.default_rng(3)
    n = 60_000
    ch = np.concatenate([np.full(n, 10), np.full(n, 300)])
    # S1: optical propagation + simple
    cfg = load_c0_config(s1_model_type='optical_propagation+simple')
    t = OM.s1_photon_delays(ch, -60.0, cfg, rng, spline=SM.s1_optical_spline())
    assert ks_p(jitter(t[:n]), jitter(gm['s1_op_top'], 1)) > P_MIN
    assert ks_p(jitter(t[n:]), jitter(gm['s1_op_bottom'], 1)) > P_MIN
    assert ks_p(jitter(t[:n]), jitter(gm['s1_op_bottom'], 1)) < 1e-6
    # S1: custom
    cfg = load_c0_config(s1_model_type='custom', led_pulse_length=300.0)
    for name, rc in (('nr', 0), ('alpha', 6), ('led', 20)):
        t = OM.s1_photon_delays(ch, -60.0, cfg, rng, recoil=rc)
        assert ks_p(jitter(t), jitter(gm['s1_custom_' + name], 1)) > P_MIN, name
    # S2: garfield + optical propagation
    cfg = load_c0_config(s2_luminescence_model='garfield', s2_time_model='optical_propagation')
    table, spline = SM.garfield_table(), SM.s2_optical_spline()
    for name, xy in (('a', [3.0, -4.0]), ('b', [10.1, 20.3])):
        lum = OM.garfield_luminescence(np.array([xy]), [2 * n], table, cfg, rng)
        t = OM.s2_photon_delays(lum, ch, cfg, rng, spline=spline)
        assert discrete_p(t[:n], gm[f's2_gf_op_{name}_top']) > P_MIN, name
        assert discrete_p(t[n:], gm[f's2_gf_op_{name}_bottom']) > P_MIN, name
    n_c = 40_000
    lum = OM.garfield_luminescence(np.tile([[3.0, -4.0]], (n_c, 1)), np.ones(n_c, np.int64), table, cfg, rng,
                                   confine_position=0.1)
    t = OM.s2_photon_delays(lum, np.full(n_c, 10), cfg, rng, spline=spline)
    assert discrete_p(t, gm['s2_gf_confined_top']) > P_MIN
    # S2: garfield gas gap
    tab, gap_of = SM.garfield_gas_gap_table(), SM.GasGapMap()
    for name, xy in GG_POS.items():
        t = OM.garfield_gas_gap_luminescence(gap_of(np.array([xy])), [120_000], tab, rng).astype(np.int64)
        assert ks_p(jitter(t), jitter(gg['gg_' + name], 1)) > P_MIN, name
    t = OM.garfield_gas_gap_luminescence(np.full(3000, gap_of(np.array([GG_POS['b']]))[0]), np.full(3000, 40), tab, rng)
    sums = t.astype(np.int64).reshape(3000, 40).sum(axis=1)
    assert ks_p(jitter(sums), jitter(gg['gg_small_sum'], 1)) > P_MIN


def test_photoelectric_electrons_and_aft_smearing_match_reference():
    from tests.stat_helpers import discrete_p
    gm = np.load(os.path.join(GOLDEN, 'stoch_models.npz'))
    rng = np.random.default_rng(5)
    cfg = load_c0_config(enable_gate_afterpulses=True, photoelectric_p=0.004)
    photons = np.arange(5000, dtype=np.int64) + 1_000_000
    calls = [OM.photoelectric_electrons(photons, cfg, rng) for _ in range(600)]
    n = np.array([len(c['time']) for c in calls])
    assert discrete_p(n, gm['pe_n']) > P_MIN
    z = np.concatenate([c['z'] for c in calls]).astype(np.float64)
    assert ks_p(-z / cfg['drift_velocity_liquid'], gm['pe_delay']) > P_MIN
    r2 = np.concatenate([c['x'].astype(np.float64) ** 2 + c['y'].astype(np.float64) ** 2 for c in calls])
    assert ks_p(r2, gm['pe_r2']) > P_MIN
    t0 = np.concatenate([c['time'] for c in calls]) - 1_000_000
    assert ks_p(jitter(t0), jitter(gm['pe_t0'], 1)) > P_MIN
    # area fraction top: constant pattern over the live PMTs, 400 photons per S2
    cfg = load_c0_config(s2_aft_sigma=0.12, s2_aft_skewness=-1.5)
    n_top = int(cfg['n_top_pmts'])
    pat = (np.asarray(cfg['gains']) != 0).astype(np.float64)
    pat /= pat.sum()
    counts = []
    for _ in range(3000):
        p = OM.smear_area_fraction_top(pat, n_top, 0.12, -1.5, rng)
        counts.append(rng.binomial(400, p[:n_top].sum()))
    assert discrete_p(np.array(counts), gm['aft_top_count']) > P_MIN
