"""Afterpulse golden samples drawn by the UNMODIFIED reference (PMT_Afterpulse.photon_afterpulse,
afterpulse.py:172-249, and PhotoIonization_Electron.electron_afterpulse, :29-88) with the
synthetic tables of tests/golden/synth_tables.py injected as resources."""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden.make_golden_stoch import fixed_rows
from tests.golden.synth_tables import EleApHist, pmt_ap_tables

HERE = os.path.dirname(os.path.abspath(__file__))
AP_S2_AMP, AP_S2_N = 400, 40


def main(ref, c0_config):
    cfg, _, _ = c0_config(enable_pmt_afterpulses=True, enable_electron_afterpulses=True,
                          photon_ap_cdfs='synthetic_ap.json.gz', ele_ap_pdfs='synthetic_ele_ap.dill')
    res = ref.load_resource.load_config(dict(cfg))
    res.uniform_to_pmt_ap = pmt_ap_tables()
    res.uniform_to_ele_ap = EleApHist()
    RL.seed_reference_rngs(4321)
    idt = ref.strax_interface.instruction_dtype
    s2 = ref.S2(dict(cfg))
    ap = ref.PMT_Afterpulse(dict(cfg))
    pi = ref.PhotoIonization_Electron(dict(cfg))
    rows = fixed_rows(idt, 2, AP_S2_AMP, AP_S2_N, -30.0, spacing=20_000_000)
    n_parent, n_ap, d_he, d_uni, a_he, pi_n, pi_delay, pi_amp, pi_r2 = [], [], [], [], [], [], [], [], []
    for r in rows:
        s2(np.array([r]))
        t, ch, g = ref.PMT_Afterpulse.photon_afterpulse(s2, res, cfg)
        n_parent.append(len(s2._photon_timings))
        n_ap.append(len(t))
        # identify element by gain: Uniform elements have amplitude exactly 1
        amp = g / np.asarray(cfg['gains'])[ch]
        # delays: match each afterpulse to its parent is not possible from the return value; use the
        # delay w.r.t. the S2 median instead (wide window), and amplitudes directly
        a_he.append(amp[np.abs(amp - 1.0) > 1e-9])
        sec = pi.generate_instruction(s2, np.array([r]))
        pi_n.append(len(sec))
        if len(sec):
            z = sec['z'].astype(np.float64)
            pi_delay.append(-z / cfg['drift_velocity_liquid'])
            pi_amp.append(sec['amp'])
            pi_r2.append(sec['x'].astype(np.float64) ** 2 + sec['y'].astype(np.float64) ** 2)
    out = dict(n_parent=np.array(n_parent), n_ap=np.array(n_ap), amp_he=np.concatenate(a_he).astype(np.float32),
               pi_n=np.array(pi_n), pi_delay=np.concatenate(pi_delay).astype(np.float32),
               pi_amp=np.concatenate(pi_amp).astype(np.int32), pi_r2=np.concatenate(pi_r2).astype(np.float32))
    # a direct, parent-resolved sample of the He-element delay: one channel, many photons
    class FakePulse:
        pass
    fp = FakePulse()
    n = 400_000
    fp._photon_timings = np.zeros(n, np.int64)
    fp._photon_channels = np.full(n, 11, np.int64)
    fp._photon_is_dpe = np.zeros(n, bool)
    fp._photon_is_dpe[: n // 4] = True
    t, ch, g = ref.PMT_Afterpulse.photon_afterpulse(fp, res, cfg)
    amp = g / np.asarray(cfg['gains'])[ch]
    is_uni = np.abs(amp - 1.0) < 1e-9
    out['delay_he'] = t[~is_uni].astype(np.float32)
    out['delay_uniform'] = t[is_uni].astype(np.float32)
    out['n_direct'] = np.array([n, (~is_uni).sum(), is_uni.sum()])
    # a larger photo-ionisation sample (drawn last, so the samples above stay what they were): many draws of
    # electron_afterpulse (afterpulse.py:29-88) on S2 calls of known size -- per call the number of type-4
    # instructions, per instruction the coarse delay (= -z / v), the electron count, r^2 and whether its
    # time zero is one of the parent's photons
    n2, d2, a2, r2, nph2, ok2 = [], [], [], [], [], []
    for r in rows[:6]:
        s2(np.array([r]))
        for _ in range(60):
            sec = pi.generate_instruction(s2, np.array([r]))
            n2.append(len(sec))
            nph2.append(len(s2._photon_timings))
            if len(sec):
                d2.append(-sec['z'].astype(np.float64) / cfg['drift_velocity_liquid'])
                a2.append(sec['amp'])
                r2.append(sec['x'].astype(np.float64) ** 2 + sec['y'].astype(np.float64) ** 2)
                ok2.append(np.isin(sec['time'] + cfg['drift_time_gate'], s2._photon_timings))
    out.update(pi2_n=np.array(n2, np.int32), pi2_n_parent_photons=np.array(nph2, np.int32),
               pi2_delay=np.concatenate(d2).astype(np.float32), pi2_amp=np.concatenate(a2).astype(np.int32),
               pi2_r2=np.concatenate(r2).astype(np.float32), pi2_t0_is_parent_photon=np.concatenate(ok2))
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.mean(v)))
    np.savez_compressed(os.path.join(HERE, 'stoch_ap.npz'), **out)
