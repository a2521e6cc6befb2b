"""Pin the stochastic part of the CPU oracle (oracle/wfsim_oracle_sim.py) against samples the
unmodified reference drew (tests/golden/stoch_c0.npz): two-sample KS / chi-square, p > 0.01."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_c0_config
from tests.golden.make_golden_stoch import S1_AMP, S1_N, S1_Z, S2_AMP, S2_N, S2_Z, fixed_rows
from tests.stat_helpers import P_MIN, chi2_counts_p, discrete_p, ks_p, mean_p
from wfsim_b200.dtypes import instruction_dtype, truth_dtype
from oracle import wfsim_oracle_sim as osim


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLDEN, 'stoch_c0.npz'))


def spe_table():
    z = np.load(os.path.join(GOLDEN, 'c0_tables.npz'))
    return z['spe_unique'][z['spe_row'][:494]]


@pytest.fixture(scope='module')
def sim():
    return osim.OracleSimulator(load_c0_config(), spe_table(), seed=99)


def test_s1_stage(sim, gold):
    rows = fixed_rows(np.dtype(instruction_dtype), 1, S1_AMP, 3 * S1_N, S1_Z)
    hits, trel, chs, npe, area = [], [], [], [], []
    for r in rows:
        r = np.array([r])
        t, ch = sim.s1_photons(r)
        t, dpe, gain = sim.pmt_stage(t, ch)
        tc = sim.truth_counters(t, ch, gain, dpe)
        hits.append(len(t)); trel.append(t - r['time'][0]); chs.append(ch)
        npe.append(tc['n_pe']); area.append(tc['raw_area'])
    assert discrete_p(hits, gold['s1_n_photon']) > P_MIN
    assert ks_p(np.concatenate(trel) + np.random.default_rng(0).random(sum(hits)),
                gold['s1_t_rel'] + np.random.default_rng(1).random(len(gold['s1_t_rel']))) > P_MIN
    assert chi2_counts_p(np.bincount(np.concatenate(chs), minlength=494), gold['s1_ch_hist']) > P_MIN
    assert mean_p(npe, gold['s1_n_pe']) > P_MIN
    assert mean_p(area, gold['s1_raw_area']) > P_MIN


def test_s2_stage(sim, gold):
    rows = fixed_rows(np.dtype(instruction_dtype), 2, S2_AMP, 2 * S2_N, S2_Z)
    ne, erel, nph, trel, chs, area = [], [], [], [], [], []
    for r in rows:
        r = np.array([r])
        t, ch, te = sim.s2_photons(r)
        t, dpe, gain = sim.pmt_stage(t, ch)
        tc = sim.truth_counters(t, ch, gain, dpe)
        ne.append(len(te)); erel.append(te - r['time'][0]); nph.append(len(t))
        trel.append(t - sim.last_te_per_photon); chs.append(ch); area.append(tc['raw_area'])
    assert discrete_p(ne, gold['s2_n_electron']) > P_MIN
    assert ks_p(np.concatenate(erel), gold['s2_e_rel']) > P_MIN
    assert mean_p(nph, gold['s2_n_photon']) > P_MIN
    tt = np.concatenate(trel)
    tt = tt[np.random.default_rng(3).choice(len(tt), 150000, replace=False)]
    assert discrete_p(tt, gold['s2_dt_photon']) > P_MIN
    assert chi2_counts_p(np.bincount(np.concatenate(chs), minlength=494), gold['s2_ch_hist']) > P_MIN
    assert mean_p(area, gold['s2_raw_area']) > P_MIN


def test_spe_and_photons_per_electron(sim, gold):
    u = np.random.default_rng(5).random(100000)
    mine = sim.spe[5, (u * 2000).astype(np.int64) + 1]
    assert discrete_p(np.round(mine), np.round(gold['spe_factor'])) > P_MIN
    rows = fixed_rows(np.dtype(instruction_dtype), 2, 100, 200, S2_Z)
    _, _, nph, _ = sim.s2_electrons(rows)
    assert discrete_p(nph, gold['s2_ph_per_e']) > P_MIN


def test_chain_truth_matches_reference_distributions(gold):
    """Whole chain through the oracle's scheduler on the same 30 events the reference simulated:
    photons per electron-amp, areas and interval counts agree."""
    cfg = load_c0_config()
    inst = gold['chain_instructions'].view(np.dtype(instruction_dtype))
    s = osim.OracleSimulator(cfg, spe_table(), seed=7)
    out = s.simulate(inst, truth_dtype=truth_dtype())
    tr = out['truth']
    assert len(tr) == len(gold['chain_truth_type'])
    for typ in (1, 2):
        m, g = tr['type'] == typ, gold['chain_truth_type'] == typ
        # per-call yields scale with amp: compare the ratios
        assert mean_p(tr['n_photon'][m] / tr['amp'][m], gold['chain_truth_n_photon'][g] / gold['chain_truth_amp'][g]) > P_MIN
        assert mean_p(tr['raw_area'][m] / np.maximum(tr['n_photon'][m], 1),
                      gold['chain_truth_raw_area'][g] / np.maximum(gold['chain_truth_n_photon'][g], 1)) > P_MIN
    n_itv = sum(g[2] for g in out['groups'])
    ref_itv, ref_samples, ref_area = gold['chain_totals']
    assert abs(n_itv - ref_itv) < 0.05 * ref_itv
    area = int((16000 - out['records']['data'].astype(np.int64))[
        np.arange(110)[None, :] < out['records']['length'][:, None]].sum())
    assert abs(area - ref_area) < 0.05 * ref_area
    # truth invariant of the reference's own test-suite (tests/test_wfsim.py:140-142 analogue)
    assert np.all(tr['n_pe'] >= tr['n_photon'])
    assert np.all(tr['n_photon_bottom'] <= tr['n_photon'])
