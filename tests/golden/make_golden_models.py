"""Golden samples for the model rows beyond config[0] -- optical propagation (s1.py:241-260,
s2.py:486-501), custom S1 timing (s1.py:201-217, 263-337), garfield luminescence (s2.py:381-409),
photo-electric electrons (afterpulse.py:92-139), field distortion (s2.py:30-71), field
dependencies (s2.py:139-179) and area-fraction-top smearing (s2.py:660-665) -- drawn by the
UNMODIFIED reference with the synthetic maps of tests/golden/synth_maps.py injected."""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden import synth_maps as SM
from tests.golden.make_golden_stoch import fixed_rows

HERE = os.path.dirname(os.path.abspath(__file__))
N = 120_000


def main(ref, c0_config):
    out = {}
    RL.seed_reference_rngs(9876)
    # ---- S1 optical propagation + simple --------------------------------------------------
    cfg, _, _ = c0_config(s1_model_type='optical_propagation+simple')
    res = ref.load_resource.load_config(dict(cfg))
    res.s1_optical_propagation_spline = SM.s1_optical_spline()
    ch = np.concatenate([np.full(N // 2, 10), np.full(N // 2, 300)]).astype(np.int64)
    pos = np.array([[3.0, -4.0, -60.0]])
    t = ref.S1.photon_timings(np.zeros(1, np.int64), np.array([N]), np.array([7]), cfg, 'liquid',
                              channels=ch, positions=pos, resource=res)
    out['s1_op_top'], out['s1_op_bottom'] = t[:N // 2].astype(np.int32), t[N // 2:].astype(np.int32)
    # ---- S1 custom: NR, alpha, LED ----------------------------------------------------------
    cfg, _, _ = c0_config(s1_model_type='custom', led_pulse_length=300.0)
    for name, rc in (('nr', 0), ('alpha', 6), ('led', 20)):
        t = ref.S1.photon_timings(np.zeros(1, np.int64), np.array([N]), np.array([rc]), cfg, 'liquid',
                                  channels=ch, positions=pos, resource=res)
        out['s1_custom_' + name] = t.astype(np.int32)
    # ---- S2: garfield luminescence + optical propagation -----------------------------------
    cfg, _, _ = c0_config(s2_luminescence_model='garfield', s2_time_model='optical_propagation')
    res = ref.load_resource.load_config(dict(cfg))
    res.s2_luminescence = SM.garfield_table()
    res.s2_optical_propagation_spline = SM.s2_optical_spline()
    for name, xy in (('a', [3.0, -4.0]), ('b', [10.1, 20.3])):
        t = ref.S2.photon_timings(np.array([xy]), np.array([N]), np.zeros(1, np.int64), np.array([N]), ch,
                                  'gas', cfg, res)
        out[f's2_gf_op_{name}_top'], out[f's2_gf_op_{name}_bottom'] = t[:N // 2].astype(np.int32), t[N // 2:].astype(np.int32)
    cfg2 = dict(cfg)
    cfg2['s2_garfield_confine_position'] = 0.1
    # one photon per position: every sample draws its own distance to the wire (independent samples)
    n_c = 40_000
    t = ref.S2.photon_timings(np.tile([[3.0, -4.0]], (n_c, 1)), np.ones(n_c, np.int64), np.zeros(n_c, np.int64),
                              np.ones(n_c, np.int64), np.full(n_c, 10, np.int64), 'gas', cfg2, res)
    out['s2_gf_confined_top'] = t.astype(np.int32)
    # ---- photo-electric electrons ---------------------------------------------------------
    cfg, _, _ = c0_config(enable_gate_afterpulses=True, photoelectric_p=0.004)
    pe = ref.afterpulse.PhotoElectric_Electron(dict(cfg))

    class FakePulse:
        pass
    idt = ref.strax_interface.instruction_dtype
    row = fixed_rows(idt, 2, 100, 1, -30.0)
    n_e, delay, r2, dt0 = [], [], [], []
    for k in range(300):
        fp = FakePulse()
        fp._photon_timings = np.arange(5000, dtype=np.int64) + 1_000_000
        sec = pe.generate_instruction(fp, row)
        n_e.append(len(sec))
        if len(sec):
            assert (sec['type'] == 6).all() and (sec['amp'] == 1).all()
            delay.append(-sec['z'].astype(np.float64) / cfg['drift_velocity_liquid'])
            r2.append(sec['x'].astype(np.float64) ** 2 + sec['y'].astype(np.float64) ** 2)
            dt0.append(sec['time'] - 1_000_000)
    out['pe_n'] = np.array(n_e, np.int32)
    out['pe_delay'] = np.concatenate(delay).astype(np.float32)
    out['pe_r2'] = np.concatenate(r2).astype(np.float32)
    out['pe_t0'] = np.concatenate(dt0).astype(np.int32)
    # ---- field distortion (deterministic) ------------------------------------------------
    rng = np.random.default_rng(5)
    x, y = rng.uniform(-35, 35, 200), rng.uniform(-35, 35, 200)
    z = rng.uniform(-140, -1, 200)

    class R:
        pass
    r = R()
    r.fdc_3d = SM.fdc_3d_map()
    r.fd_comsol = SM.fd_comsol_map()
    zo, po = ref.S2.inverse_field_distortion_correction(x, y, z, r)
    out['fd_xyz'] = np.stack([x, y, z])
    out['fd_inverse_fdc'] = np.stack([po[:, 0], po[:, 1], zo])
    zo, po = ref.S2.field_distortion_comsol(x, y, z, r)
    out['fd_comsol'] = np.stack([po[:, 0], po[:, 1], zo])
    # ---- field dependencies: drift time parameters + electron times -------------------------
    cfg, _, _ = c0_config(enable_field_dependencies=dict(survival_probability_map=True, drift_speed_map=True,
                                                         diffusion_longitudinal_map=True,
                                                         diffusion_transverse_map=False))
    res = ref.load_resource.load_config(dict(cfg))
    fd = SM.FieldDependencies()
    res.field_dependencies_map = fd.field_dependencies_map
    res.diffusion_longitudinal_map = fd.diffusion_longitudinal_map
    res.drift_velocity_scaling = 1.0
    xy = np.array([[20.0, 15.0]])
    zz = np.array([-80.0])
    mean, spread = ref.S2.get_s2_drift_time_params(zz, xy, cfg, res)
    out['fdep_mean_spread'] = np.array([mean[0], spread[0]])
    ne = ref.S2.get_electron_yield(np.full(400, 500), np.tile(xy, (400, 1)), np.full(400, -80.0),
                                   np.tile(xy, (400, 1)), cfg, res)
    out['fdep_n_electron'] = ne.astype(np.int32)
    sc = ref.S2.get_s2_light_yield(xy, cfg, res)
    _, _, et = ref.S2.get_n_photons(np.zeros(1, np.int64), np.array([50_000]), zz, xy, sc, cfg, res)
    out['fdep_e_t'] = et.astype(np.int32)
    # ---- area-fraction-top smearing ------------------------------------------------------
    cfg, _, _ = c0_config(s2_aft_sigma=0.12, s2_aft_skewness=-1.5)
    cfg['turned_off_pmts'] = np.arange(len(cfg['gains']))[np.array(cfg['gains']) == 0]   # pulse.py:31
    res = ref.load_resource.load_config(dict(cfg))
    n_i, n_ph = 600, 400
    chans = ref.S2.photon_channels(np.full(n_i, 10), np.full(n_i, -30.0), np.tile([[3.0, -4.0]], (n_i, 1)),
                                   np.repeat(np.arange(n_i), n_ph), cfg, res)
    out['aft_top_count'] = (chans.reshape(n_i, n_ph) < cfg['n_top_pmts']).sum(axis=1).astype(np.int32)
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.mean(v)))
    np.savez_compressed(os.path.join(HERE, 'stoch_models.npz'), **out)
