"""Stochastic-stage golden samples drawn by the UNMODIFIED reference (see make_golden.py).

Inputs are fixed instruction rows; outputs are samples of the reference's own stage outputs,
stored relative to the instruction time so they are small integers.  The GPU tests compare the
Philox kernels against these with two-sample KS / chi-square tests (p > 0.01), as BASELINE.json
specifies for the stochastic stages."""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden.synth_instructions import c0_like

HERE = os.path.dirname(os.path.abspath(__file__))

S1_AMP, S1_N, S1_Z = 4000, 150, -40.0
S2_AMP, S2_N, S2_Z = 250, 60, -60.0


def fixed_rows(dtype, typ, amp, n, z, t0=10_000_000, spacing=5_000_000):
    rows = np.zeros(n, dtype=dtype)
    rows['type'] = typ
    rows['time'] = t0 + spacing * np.arange(n)
    rows['x'], rows['y'], rows['z'] = 3.0, -4.0, z
    rows['amp'] = amp
    rows['recoil'] = 7
    rows['local_field'] = 82.0
    rows['event_number'] = np.arange(n)
    return rows


def main(ref, c0_config):
    cfg, _, _ = c0_config()
    RL.seed_reference_rngs(1234)
    idt = ref.strax_interface.instruction_dtype
    out = {}
    # ---- S1 stage + PMT stage ----
    s1 = ref.S1(dict(cfg))
    rows = fixed_rows(idt, 1, S1_AMP, S1_N, S1_Z)
    n_ph, t_rel, chs, gains_rel, n_pe, area = [], [], [], [], [], []
    for r in rows:
        s1(np.array([r]))
        n_ph.append(len(s1._photon_timings))
        t_rel.append(s1._photon_timings - r['time'])
        chs.append(s1._photon_channels)
        n_pe.append(s1._truth_buffer['n_pe'])
        area.append(s1._truth_buffer['raw_area'])
    out['s1_n_photon'] = np.array(n_ph, np.int32)
    out['s1_t_rel'] = np.concatenate(t_rel).astype(np.int32)
    out['s1_ch_hist'] = np.bincount(np.concatenate(chs), minlength=494).astype(np.int64)
    out['s1_n_pe'] = np.array(n_pe, np.int32)
    out['s1_raw_area'] = np.array(area, np.float64)
    # ---- S2 stage ----
    s2 = ref.S2(dict(cfg))
    rows = fixed_rows(idt, 2, S2_AMP, S2_N, S2_Z)
    n_e, e_rel, n_ph, t_rel, chs, n_pe, area, ph_per_e = [], [], [], [], [], [], [], []
    for r in rows:
        s2(np.array([r]))
        n_e.append(len(s2._electron_timings))
        e_rel.append(s2._electron_timings - r['time'])
        n_ph.append(len(s2._photon_timings))
        t_rel.append(s2._photon_timings - r['time'])
        chs.append(s2._photon_channels)
        n_pe.append(s2._truth_buffer['n_pe'])
        area.append(s2._truth_buffer['raw_area'])
    out['s2_n_electron'] = np.array(n_e, np.int32)
    out['s2_e_rel'] = np.concatenate(e_rel).astype(np.int32)
    out['s2_n_photon'] = np.array(n_ph, np.int32)
    # photon delay w.r.t. its electron: S2.photon_timings with ONE electron at t = 0 (photons of
    # a real S2 share electron times, so they are not independent samples), plus the transit
    # time term exactly as Pulse.__call__ adds it (pulse.py:54-56)
    n_one = 150_000
    res0 = ref.load_resource.load_config(dict(cfg))
    dt_ph = ref.S2.photon_timings(np.array([[3.0, -4.0]]), np.array([n_one]), np.zeros(1, np.int64),
                                  np.array([n_one]), np.zeros(n_one, np.int64), 'gas', cfg, res0)
    dt_ph = dt_ph + np.random.normal(cfg['pmt_transit_time_mean'], cfg['pmt_transit_time_spread'] / 2.35482,
                                     n_one).astype(np.int64)
    out['s2_dt_photon'] = dt_ph.astype(np.int32)
    out['s2_ch_hist'] = np.bincount(np.concatenate(chs), minlength=494).astype(np.int64)
    out['s2_n_pe'] = np.array(n_pe, np.int32)
    out['s2_raw_area'] = np.array(area, np.float64)
    # photons per electron and the SPE gain distribution (one more direct draw)
    xy = np.tile([[3.0, -4.0]], (200, 1))
    res = ref.load_resource.load_config(dict(cfg))
    sc = ref.S2.get_s2_light_yield(xy, cfg, res)
    _, per_e, _ = ref.S2.get_n_photons(np.zeros(200, np.int64), np.full(200, 100), np.full(200, S2_Z),
                                       xy, sc, cfg, res)
    out['s2_ph_per_e'] = np.asarray(per_e, np.int32)
    u = np.random.random(100_000)
    out['spe_factor'] = s2.uniform_to_pe_arr(u, 5).astype(np.float32)
    # ---- whole chain through the reference scheduler on C0-like events ----
    inst = c0_like(30, seed=77)
    rd = ref.RawData(dict(cfg))
    tdt = np.dtype(ref.strax_interface.instruction_dtype + ref.strax_interface.truth_extra_dtype + [('fill', bool)])
    tb = np.zeros(1000, tdt)
    n_itv, n_samples = 0, 0
    per_group = []
    area_adc = 0
    for ch, left, right, data in rd(inst.astype(idt), tb, progress_bar=False):
        n_itv += 1
        n_samples += right - left + 1
        area_adc += int((16000 - data).sum())
    tb = tb[tb['fill']]
    out['chain_truth_type'] = tb['type'].astype(np.int8)
    out['chain_truth_amp'] = tb['amp'].astype(np.int32)
    out['chain_truth_n_photon'] = tb['n_photon'].astype(np.int32)
    out['chain_truth_n_electron'] = tb['n_electron'].astype(np.int32)
    out['chain_truth_raw_area'] = tb['raw_area']
    out['chain_truth_z'] = tb['z']
    out['chain_truth_t_sigma_photon'] = tb['t_sigma_photon']
    out['chain_totals'] = np.array([n_itv, n_samples, area_adc], np.int64)
    out['chain_instructions'] = inst.view(np.uint8)
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, 'stoch_c0.npz'), **out)
