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


# ---- further optional model rows, pinned to tests/golden/stoch_models.npz and stoch_gg.npz ------------

def singlet_triplet_delays(n, singlet_fraction, t_singlet, t_triplet, rng):
    """Excimer decay delay (pulse.py:321-341): lifetime picked per photon, Exp(1) x lifetime, truncated."""
    life = np.where(rng.random(n) < singlet_fraction, t_singlet, t_triplet)
    return (rng.exponential(1.0, n) * life).astype(np.int64)


def s1_photon_delays(channels, z, cfg, rng, spline=None, recoil=None):
    """Arrival-time terms of S1 photons relative to the interaction time (s1.py:162-260), each truncated to
    integer ns on its own: optical propagation from the (z, U) spline of the photon's array, the `simple`
    exponential decay + normal spread, the `custom` recoil-dependent excimer delays (NR 0, alpha 6: singlet
    / triplet in liquid; LED 20: uniform over the pulse length)."""
    channels = np.asarray(channels)
    n = len(channels)
    model = cfg['s1_model_type']
    t = np.zeros(n, np.int64)
    if 'optical_propagation' in model:
        pts = np.stack([np.full(n, float(z)), rng.random(n)], axis=1)
        top = channels < cfg['n_top_pmts']
        prop = np.zeros(n, np.int64)
        prop[top] = np.asarray(spline(pts[top], map_name='top')).astype(np.int64)
        prop[~top] = np.asarray(spline(pts[~top], map_name='bottom')).astype(np.int64)
        t += prop
    if 'simple' in model:
        t += rng.exponential(cfg['s1_decay_time'], n).astype(np.int64)
        t += rng.normal(0, cfg['s1_decay_spread'], n).astype(np.int64)
    if 'custom' in model:
        if recoil == 20:
            t += rng.uniform(0, cfg['led_pulse_length'], n).astype(np.int64)
        elif recoil in (0, 6):
            frac = cfg['s1_NR_singlet_fraction'] if recoil == 0 else cfg['s1_ER_alpha_singlet_fraction']
            t += singlet_triplet_delays(n, frac, cfg['singlet_lifetime_liquid'], cfg['triplet_lifetime_liquid'], rng)
        else:
            raise AttributeError('Recoil type must be ER, NR, alpha or LED')
    return t


def garfield_luminescence(xy, n_photons, table, cfg, rng, confine_position=None):
    """'garfield' luminescence (s2.py:381-409): the table row nearest to the instruction's distance from
    the closest anode wire (or to a uniform draw within +-confine_position), a uniformly drawn column per
    photon, minus the integer mean of the whole table."""
    xy = np.asarray(xy, dtype=np.float64)
    if isinstance(confine_position, float):
        dist = rng.uniform(-confine_position, confine_position, len(xy))
    else:
        tilt, pitch = cfg.get('anode_xaxis_angle', np.pi / 4), cfg.get('anode_pitch', 0.5)
        across = -xy[:, 0] * np.sin(tilt) + xy[:, 1] * np.cos(tilt)
        dist = (across + pitch / 2) % pitch - pitch / 2
    rows = np.array([int(np.argmin(np.abs(d - table['x']))) for d in dist], dtype=np.int64)
    rows = np.repeat(rows, n_photons)
    cols = rng.integers(0, table['t'].shape[1], len(rows))
    return table['t'][rows, cols].astype(np.int64) - int(np.average(table['t']))


def garfield_gas_gap_luminescence(gas_gaps, n_photons, gg_table, rng):
    """'garfield_gas_gap' luminescence (s2.py:411-483): inverse CDF blended linearly between the tabulated
    gas gap below the local one (np.digitize - 1, python indexing) and the next, sampled at U(0, len - 2)
    with linear interpolation between entries, minus the mean over the instruction's photons (float;
    the caller truncates, s2.py:532-533)."""
    grid = np.asarray(gg_table['gas_gap'], dtype=np.float64)
    cdfs = np.asarray(gg_table['timing_inv_cdf'], dtype=np.float64)
    spacing = grid[1] - grid[0]
    out = []
    for gap, n in zip(np.asarray(gas_gaps, dtype=np.float64), n_photons):
        lo = int(np.digitize(gap, grid)) - 1
        hi = min(max(lo + 1, 0), len(grid) - 1)
        curve = (cdfs[hi] - cdfs[lo]) * ((gap - grid[lo]) / spacing) + cdfs[lo]
        s = rng.uniform(0, cdfs.shape[1] - 2, int(n))
        a, b = curve[np.floor(s).astype(int)], curve[np.ceil(s).astype(int)]
        t = (b - a) * (s - np.floor(s)) + a
        out.append(t - t.mean() if n else t)
    return np.concatenate(out) if out else np.zeros(0)


def s2_photon_delays(luminescence, channels, cfg, rng, spline=None):
    """S2 photon time relative to its electron (s2.py:504-557): luminescence + excimer delay in gas +
    optical propagation (U spline per array) / zero / normal spread, each truncated on its own."""
    channels = np.asarray(channels)
    n = len(channels)
    t = np.asarray(luminescence).astype(np.int64)
    t = t + singlet_triplet_delays(n, cfg['singlet_fraction_gas'], cfg['singlet_lifetime_gas'],
                                   cfg['triplet_lifetime_gas'], rng)
    model = cfg['s2_time_model']
    if 'optical_propagation' in model:
        u = rng.random(n)[:, None]
        top = channels < cfg['n_top_pmts']
        prop = np.zeros(n, np.int64)
        prop[top] = np.asarray(spline(u[top], map_name='top')).astype(np.int64)
        prop[~top] = np.asarray(spline(u[~top], map_name='bottom')).astype(np.int64)
        t = t + prop
    elif 'zero_delay' in model:
        pass
    elif 's2_time_spread around zero' in model:
        t = t + rng.normal(0, cfg['s2_time_spread'], n).astype(np.int64)
    else:
        raise KeyError(model)
    return t


def photoelectric_electrons(photon_times, cfg, rng):
    """Photo-electric (gate) electrons of one S2 pulse call (afterpulse.py:105-139): Poisson(p x photons x
    modifier) single-electron instructions of type 6; creation time = a random parent photon + gate drift
    time, depth from a normal delay clipped at 0, position uniform over the TPC cross-section."""
    photon_times = np.asarray(photon_times)
    n = rng.poisson(cfg['photoelectric_p'] * len(photon_times) * cfg.get('photoelectric_modifier', 1))
    delay = np.clip(rng.normal(cfg['photoelectric_t_center'] + cfg['drift_time_gate'],
                               cfg['photoelectric_t_spread'], n), 0, None)
    time = photon_times[rng.integers(0, len(photon_times), n)] + cfg['drift_time_gate']
    r = np.sqrt(rng.uniform(0, cfg['tpc_radius'] ** 2, n))
    phi = rng.uniform(-np.pi, np.pi, n)
    return dict(time=time.astype(np.int64), x=(r * np.cos(phi)).astype(np.float32),
                y=(r * np.sin(phi)).astype(np.float32),
                z=(-delay * cfg['drift_velocity_liquid']).astype(np.float32))


def smear_area_fraction_top(pattern, n_top, sigma, skewness, rng):
    """s2.py:660-665 on one normalised pattern row: the top-array share is multiplied by a skew-normal
    factor around 1 (clipped to [0, 1]) and the bottom share rescaled to keep the sum."""
    from scipy.stats import skewnorm
    p = np.array(pattern, dtype=np.float64)
    cur = p[:n_top].sum() / p.sum()
    new = float(np.clip(cur * skewnorm.rvs(loc=1.0, scale=sigma, a=skewness, random_state=rng), 0, 1))
    p[:n_top] *= new / cur
    p[n_top:] *= (1 - new) / (1 - cur)
    return p
