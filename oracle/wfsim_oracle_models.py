"""TEST INFRASTRUCTURE ONLY -- CPU restatements (numpy) of two optional S2 model rows of the WFSim hot
path, pinned in tests/test_oracle_models.py to samples drawn by the unmodified reference
(tests/golden/stoch_lumw.npz, stoch_diffuse.npz):

  simple luminescence with a per-position gas gap   wfsim/core/s2.py:317-378 (enable_gas_gap_warping)
  transverse diffusion of the S2 hit pattern        wfsim/core/s2.py:560-613 (diffusion_transverse_map)

Random numbers come from a numpy Generator: comparable to the reference and to the CUDA path only
statistically.  Only tests/ may import this module.
"""
import numpy as np

_E_SI = 1.602176565e-19
_KB = 1.3806488e-23 / _E_SI                   # eV / K          (wfsim/units.py)
_BAR = 1e5 / _E_SI / 100.0 / 100.0 ** 2
_KV_PER_CM = 1000.0


def field_scale(cfg, gas_gap):
    """E0 of s2.py:365-370 for gas gaps [cm]."""
    gap = np.asarray(gas_gap, dtype=np.float64)
    r_a, r_w = cfg['anode_field_domination_distance'], cfg['anode_wire_radius']
    liquid = cfg['gate_to_anode_distance'] - gap
    v_gas = cfg['anode_voltage'] / (1 + liquid / gap / cfg['lxe_dielectric_constant'])
    return v_gas / ((gap - r_a) / r_a + np.log(r_a / r_w))


def luminescence_timings_warped(gas_gaps, n_photons, cfg, rng):
    """Emission times of the photons of ONE S2 call whose instructions sit at gas gaps `gas_gaps`
    (s2.py:317-378).  The radial grid (step 1e-4 cm) runs from the largest gap of the call to the wire;
    every instruction uses its own field scale on that grid, subtracts the yield-weighted mean time over
    the WHOLE grid, and inverts the cumulative yield from its own gap downwards."""
    gaps = np.asarray(gas_gaps, dtype=np.float64)
    n_gas = cfg['pressure'] / (_KB * cfg['temperature'])
    alpha = cfg['gas_drift_velocity_slope'] / n_gas
    pressure = cfg['pressure'] / _BAR
    r_a, r_w = cfg['anode_field_domination_distance'], cfg['anode_wire_radius']
    e0 = field_scale(cfg, gaps)
    step = 0.0001
    r = np.arange(gaps.max(), r_w, -step)
    inv_r = np.clip(1 / r, 1 / r_a, 1 / r_w)
    out = []
    for gap, e, n in zip(gaps, e0, n_photons):
        dt = step / (alpha * e * inv_r)
        dy = e * inv_r / _KV_PER_CM - 0.8 * pressure
        mean_t = np.sum(np.cumsum(dt) * dy) / np.sum(dy)
        first = int(np.argmax(r <= gap))
        t = np.cumsum(dt[first:]) - mean_t
        y = np.cumsum(dy[first:])
        out.append(np.interp(rng.random(int(n)), y / y[-1], t).astype(np.int64))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


def s2_pattern_diffuse(n_electron, xy, sigma_radial, sigma_azimuthal, pattern_map, tpc_radius, rng):
    """Hit pattern [n_instr, n_pmt] of S2 instructions with transverse diffusion (s2.py:560-613): every
    electron is displaced by N(0, sigma_radial) along the radius through the interaction and
    N(0, sigma_azimuthal) across it; the pattern is the mean of the map over the instruction's electrons
    that stay inside `tpc_radius` (NaN when none does)."""
    xy = np.asarray(xy, dtype=np.float64)
    theta = np.arctan2(xy[:, 1], xy[:, 0])
    rows = []
    for i, n in enumerate(np.asarray(n_electron, dtype=np.int64)):
        along = rng.normal(0.0, 1.0, n) * sigma_radial[i]
        across = rng.normal(0.0, 1.0, n) * sigma_azimuthal[i]
        c, s = np.cos(theta[i]), np.sin(theta[i])
        pos = xy[i] + np.stack([c * along - s * across, s * along + c * across], axis=1)
        pos = pos[(pos ** 2).sum(axis=1) <= tpc_radius ** 2]
        val = np.asarray(pattern_map(pos), dtype=np.float64)
        rows.append(val.mean(axis=0) if len(pos) else np.full(val.shape[-1], np.nan))
    return np.array(rows)


def hdiff_sigmas(cfg, z_obs, xy_obs, field_dependencies_map, drift_velocity_scaling=1.0):
    """sigma of one electron's radial / azimuthal displacement (s2.py:573-586): sqrt(2 D t_drift) with D
    from the field maps [cm^2/s] and the drift time from the average drift velocity (s2.py:139-155)."""
    efd = cfg.get('enable_field_dependencies', {})
    if efd.get('drift_speed_map'):
        v = np.asarray(field_dependencies_map(z_obs, xy_obs, map_name='drift_speed_map'), np.float64).reshape(-1)
        v = v * 1e-4 * drift_velocity_scaling
    else:
        v = cfg['drift_velocity_liquid']
    t_drift = -np.asarray(z_obs, np.float64) / v
    out = []
    for name in ('diffusion_radial_map', 'diffusion_azimuthal_map'):
        d = np.asarray(field_dependencies_map(z_obs, xy_obs, map_name=name), np.float64).reshape(-1) * 1e-9
        out.append(np.sqrt(2 * d * t_drift))
    return out
