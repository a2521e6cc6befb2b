"""Host-side resources (maps / tables) for the B200 path -- the counterpart of
wfsim/load_resource.py:Resource, restricted to what can be loaded without the XENON private
file servers.  Maps given as ["constant dummy", value, shape] become `DummyMap`s exactly as in
the reference (load_resource.py:383-391, 438-457); file-backed maps need straxen's
InterpolatingMap (third party) and are accepted as ready-made callables via `overrides`.
"""
import numpy as np


class DummyMap:
    """Constant map: result shape [len(x)] + shape (load_resource.py:438-457)."""

    def __init__(self, const, shape=()):
        self.const = const
        self.shape = tuple(shape)

    def __call__(self, x, **kwargs):
        return np.ones([len(x)] + list(self.shape)) * self.const


class GridMap:
    """Named value arrays on one regular grid, evaluated with multilinear interpolation and
    linear extrapolation outside the grid -- the arithmetic of
    scipy.interpolate.RegularGridInterpolator(method='linear', bounds_error=False, fill_value=None),
    which is what straxen.InterpolatingMap(method='RegularGridInterpolator') (third party, used
    by load_resource.py:399,433) wraps.  The same grids are uploaded to the device for the
    per-photon maps (optical propagation), so host and device evaluate identical tables.

    axes: [(lo, hi, n), ...] one per coordinate; maps: {name: array of shape (n0, n1, ...[, k])}."""

    def __init__(self, axes, maps):
        self.axes = [(float(lo), float(hi), int(n)) for lo, hi, n in axes]
        self.maps = {k: np.asarray(v, dtype=np.float64) for k, v in maps.items()}
        for k, v in self.maps.items():
            if tuple(v.shape[:len(self.axes)]) != tuple(n for _, _, n in self.axes):
                raise ValueError(f'map {k}: shape {v.shape} does not match the grid')

    def grid(self, map_name='map'):
        return self.axes, self.maps[map_name]

    def __call__(self, points, map_name='map'):
        v = self.maps[map_name]
        pts = np.asarray(points, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts[:, None]
        nd = len(self.axes)
        idx, w = [], []
        for d, (lo, hi, n) in enumerate(self.axes):
            f = (pts[:, d] - lo) / (hi - lo) * (n - 1)
            i = np.clip(np.floor(f).astype(np.int64), 0, n - 2)
            idx.append(i)
            w.append(f - i)
        out = 0.0
        for corner in range(1 << nd):
            sel, wt = [], 1.0
            for d in range(nd):
                bit = (corner >> (nd - 1 - d)) & 1
                sel.append(idx[d] + bit)
                wt = wt * (w[d] if bit else 1.0 - w[d])
            val = v[tuple(sel)]
            out = out + (wt[(...,) + (None,) * (val.ndim - 1)] * val)
        return out


def _make_map(spec, name):
    if callable(spec):
        return spec
    if isinstance(spec, (list, tuple)):
        if spec[0] != 'constant dummy':
            raise AssertionError('Alternative file input can only be ("constant dummy", constant: int, shape: list')
        return DummyMap(spec[1], spec[2])
    if isinstance(spec, str):
        try:
            import straxen  # noqa: F401  (third party, optional)
        except ImportError:
            raise RuntimeError(
                f'{name}={spec!r} is a file-backed map; loading it needs straxen.InterpolatingMap. '
                'Pass a callable through `overrides` or use a ["constant dummy", ...] spec.')
        fmt = spec.split('.', 1)[-1]
        return straxen.InterpolatingMap(straxen.get_resource(spec, fmt=fmt))
    raise TypeError("Can't handle map_file except a string or a list")


class Resource:
    """Attribute names follow the reference's Resource so user code can be shared."""

    MAPS = ('s1_pattern_map', 's2_pattern_map', 's1_lce_correction_map', 's2_correction_map',
            'se_gain_map', 'field_dependencies_map')

    def __init__(self, config, **overrides):
        for name in self.MAPS:
            spec = overrides.get(name, config.get(name))
            if spec is None or spec == '':
                continue
            m = _make_map(spec, name)
            if name == 'field_dependencies_map' and not isinstance(m, DummyMap) and not callable(spec):
                m = _rz_wrapper(m)
            elif name == 'field_dependencies_map' and isinstance(m, DummyMap):
                m = _rz_wrapper(m)
            setattr(self, name, m)
        if not hasattr(self, 's1_lce_correction_map') and isinstance(getattr(self, 's1_pattern_map', None), DummyMap):
            # load_resource.py:245-250: LCE map = pattern map summed over PMTs
            pm = self.s1_pattern_map
            self.s1_lce_correction_map = DummyMap(pm.const * (pm.shape[-1] if pm.shape else 1), ())
        for name in ('photon_area_distribution', 'spe_ppf', 'spe_row', 'noise_data',
                     'uniform_to_pmt_ap', 'uniform_to_ele_ap',
                     # load_resource.py:262-330: timing splines, luminescence tables, field maps
                     's1_optical_propagation_spline', 's2_optical_propagation_spline', 's2_luminescence',
                     'fdc_3d', 'fd_comsol', 'diffusion_longitudinal_map', 'drift_velocity_scaling',
                     's2_luminescence_gg', 'garfield_gas_gap_map', 'gas_gap_length'):
            if name in overrides:
                setattr(self, name, overrides[name])
        if not hasattr(self, 'drift_velocity_scaling'):
            self.drift_velocity_scaling = 1.0


def resource_from_reference(ref_resource, config):
    """Adapter for a WFSim deployment: a Resource holding the maps and tables of the reference's own
    `wfsim.load_config(config)` object (load_resource.py:176-380), attribute by attribute.  Constant
    dummy maps are re-wrapped so that they share one pattern row; every other map object is used as
    the callable it is (evaluated on the host per instruction) -- pass wfsim_b200.resource.GridMap
    objects instead for the maps that should be interpolated on the device."""
    names = Resource.MAPS + ('photon_area_distribution', 'noise_data', 'uniform_to_pmt_ap',
                             'uniform_to_ele_ap', 's1_optical_propagation_spline',
                             's2_optical_propagation_spline', 's2_luminescence', 's2_luminescence_gg',
                             'garfield_gas_gap_map', 'gas_gap_length', 'fdc_3d', 'fd_comsol',
                             'diffusion_longitudinal_map', 'drift_velocity_scaling')
    overrides = {}
    for name in names:
        if not hasattr(ref_resource, name):
            continue
        v = getattr(ref_resource, name)
        if type(v).__name__ == 'DummyMap' and hasattr(v, 'const') and hasattr(v, 'shape'):
            v = DummyMap(v.const, v.shape)
        overrides[name] = v
    return Resource(config, **overrides)


def _rz_wrapper(m):
    def rz_map(z, xy, **kwargs):            # load_resource.py:335-338
        r = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
        return m(np.array([r, z]).T, **kwargs)
    return rz_map


def inverse_field_distortion_correction(x, y, z, resource):
    """s2.py:30-53: six fixed-point iterations of the fdc_3d map."""
    positions = np.array([x, y, z]).T
    dr_pre = None
    for i_iter in range(6):
        dr = np.asarray(resource.fdc_3d(positions), dtype=np.float64).reshape(-1)
        if i_iter > 0:
            dr = 0.5 * dr + 0.5 * dr_pre
        dr_pre = dr
        r_obs = np.sqrt(x ** 2 + y ** 2) - dr
        x_obs = x * r_obs / (r_obs + dr)
        y_obs = y * r_obs / (r_obs + dr)
        z_obs = -np.sqrt(z ** 2 + dr ** 2)
        positions = np.array([x_obs, y_obs, z_obs]).T
    return z_obs, np.array([x_obs, y_obs]).T


def field_distortion_comsol(x, y, z, resource):
    """s2.py:56-71."""
    positions = np.array([np.sqrt(x ** 2 + y ** 2), z]).T
    theta = np.arctan2(y, x)
    r_obs = np.asarray(resource.fd_comsol(positions, map_name='r_distortion_map'), dtype=np.float64).reshape(-1)
    return z, np.array([r_obs * np.cos(theta), r_obs * np.sin(theta)]).T


def _on_device(m, nd, n_ch, need_full):
    """Is this pattern map a regular grid the library evaluates on the device (params.build_tables)?"""
    if not hasattr(m, 'grid') or 'map' not in getattr(m, 'maps', {}):
        return False
    axes, vals = m.grid('map')
    ok = len(axes) == nd and vals.ndim == nd + 1 and vals.shape[-1] <= n_ch and min(a[2] for a in axes) >= 2
    return ok and (vals.shape[-1] == n_ch or not need_full)


def evaluate_instruction_maps(config, resource, instructions, seed=0, device_patterns=True, rng_id=None):
    """Per-instruction map values handed to the device (struct wfs_instr_maps):
    S1 light yield (s1.py:125), observed S2 positions after the field-distortion model
    (s2.py:80-87), S2 secondary-scintillation gain (s2.py:182-209), the survival / extraction
    factor of the electron yield (s2.py:227-252), drift velocity / longitudinal diffusion from the
    field-dependency maps (s2.py:139-179) and the un-normalised PMT patterns (s1.py:148,
    s2.py:637-665, with the optional area-fraction-top smearing -- its draw is a Philox function of
    (seed, rng_id of the instruction; default: its index), like every draw on the device)."""
    n = len(instructions)
    rng_id = np.arange(n, dtype=np.uint64) if rng_id is None else np.asarray(rng_id, np.uint64)
    n_ch = len(config['gains'])
    typ = instructions['type']
    is_s1 = typ == 1
    is_s2 = ~is_s1
    xyz = np.stack([instructions['x'], instructions['y'], instructions['z']], axis=1).astype(np.float64)
    xy = xyz[:, :2]
    s1_lce = np.ones(n)
    sc_gain = np.zeros(n)
    cy_extra = np.ones(n)
    out = {}
    p_dpe = config['p_double_pe_emision']
    efd = config.get('enable_field_dependencies', {})
    if is_s1.any():
        ly = np.asarray(resource.s1_lce_correction_map(xyz[is_s1]), dtype=np.float64)
        if ly.ndim != 1:
            ly = np.squeeze(ly, axis=-1)
        s1_lce[is_s1] = ly
    pos_obs = xy.copy()            # observed xy of the S2-like instructions
    if is_s2.any():
        x, y, z = xyz[is_s2, 0], xyz[is_s2, 1], xyz[is_s2, 2]
        fdm = config.get('field_distortion_model', 'none')
        if fdm == 'inverse_fdc':
            z_obs, pos = inverse_field_distortion_correction(x, y, z, resource)
        elif fdm == 'comsol':
            z_obs, pos = field_distortion_comsol(x, y, z, resource)
        else:
            z_obs, pos = z, xy[is_s2]
        pos_obs[is_s2] = pos
        if fdm in ('inverse_fdc', 'comsol'):
            out['x_obs'], out['y_obs'] = pos_obs[:, 0].copy(), pos_obs[:, 1].copy()
        if config.get('se_gain_from_map', False):
            g = np.asarray(resource.se_gain_map(pos), dtype=np.float64)
        else:
            g = np.asarray(resource.s2_correction_map(pos), dtype=np.float64) * config['s2_secondary_sc_gain']
        if g.ndim != 1:
            g = np.squeeze(g, axis=-1)
        g = g / (1 + p_dpe)
        g[np.isnan(g)] = 0
        sc_gain[is_s2] = g
        if config.get('ext_eff_from_map', False):
            rel = np.asarray(resource.s2_correction_map(pos), dtype=np.float64).flatten()
            se = (np.asarray(resource.se_gain_map(pos)).flatten() if config.get('se_gain_from_map', False)
                  else rel * config['s2_secondary_sc_gain'])
            cy_extra[is_s2] = config['g2_mean'] * rel / se / config['electron_extraction_yield']
        # survival probability / drift maps are in TRUE coordinates (s2.py:91-93)
        if efd.get('survival_probability_map'):
            ps = np.asarray(resource.field_dependencies_map(
                z, xy[is_s2], map_name='survival_probability_map'), dtype=np.float64).reshape(-1)
            cy_extra[is_s2] *= np.clip(ps, 0, 1)
        if efd.get('drift_speed_map'):
            v = np.asarray(resource.field_dependencies_map(z, xy[is_s2], map_name='drift_speed_map'),
                           dtype=np.float64).reshape(-1) * 1e-4 * resource.drift_velocity_scaling
            vd = np.full(n, float(config['drift_velocity_liquid']))
            vd[is_s2] = v
            out['drift_velocity'] = vd
        if efd.get('diffusion_longitudinal_map'):
            d = np.asarray(resource.diffusion_longitudinal_map(z, xy[is_s2]), dtype=np.float64).reshape(-1)
            dl = np.full(n, float(config['diffusion_constant_longitudinal']))
            dl[is_s2] = d
            out['diffusion_long'] = dl
    diffuse = bool(efd.get('diffusion_transverse_map')) and config.get('diffusion_constant_transverse', 0) > 0
    if diffuse and is_s2.any():
        # s2_pattern_map_diffuse (s2.py:560-613): sigma of one electron's radial / azimuthal displacement,
        # from the maps at the OBSERVED position (photon_channels gets z_obs, positions: s2.py:112-117);
        # the per-electron averaging of the pattern runs on the device (k_pattern_diffuse)
        assert np.all(z_obs < 0), 'All S2 in liquid should have z < 0'
        if efd.get('drift_speed_map'):                  # get_avg_drift_velocity, s2.py:139-155
            v_avg = np.asarray(resource.field_dependencies_map(z_obs, pos, map_name='drift_speed_map'),
                               dtype=np.float64).reshape(-1) * 1e-4 * resource.drift_velocity_scaling
        else:
            v_avg = float(config['drift_velocity_liquid'])
        t_mean = -z_obs / v_avg
        sr, sa = np.zeros(n), np.zeros(n)
        for dst, name in ((sr, 'diffusion_radial_map'), (sa, 'diffusion_azimuthal_map')):
            d = np.asarray(resource.field_dependencies_map(z_obs, pos, map_name=name),
                           dtype=np.float64).reshape(-1) * 1e-9          # cm^2/s -> cm^2/ns
            dst[is_s2] = np.sqrt(2 * d * t_mean)
        out['hdiff_sigma_r'], out['hdiff_sigma_a'] = sr, sa
    if config.get('s2_luminescence_model', 'simple') == 'simple' and config.get('enable_gas_gap_warping', False) \
            and is_s2.any():
        # s2.py:361-370: local gas gap at the observed position and the field scale that follows from it
        from .tables import luminescence_field_scale
        if not hasattr(resource, 'gas_gap_length'):
            # load_resource.py:319-321 reads the warping map (a pickled histogram) through straxen
            raise RuntimeError('enable_gas_gap_warping needs resource.gas_gap_length: pass a callable '
                               'xy -> gas gap [cm] through the Resource overrides')
        gap = np.zeros(n)
        gap[is_s2] = np.asarray(resource.gas_gap_length(pos_obs[is_s2]), dtype=np.float64).reshape(-1)
        e0 = np.zeros(n)
        e0[is_s2] = luminescence_field_scale(config, gap[is_s2])
        out['lum_gap'], out['lum_e0'] = gap, e0
    if config.get('s2_luminescence_model', 'simple') == 'garfield_gas_gap' and is_s2.any():
        # s2.py:460-483: the excitation-time inverse CDF is interpolated between the two tabulated gas gaps
        # around the local one (np.digitize - 1: below the first gap python's index -1 picks the LAST row)
        gg = resource.s2_luminescence_gg
        gaps = np.asarray(gg['gas_gap'], dtype=np.float64)
        n_rows = len(gaps)
        cont = np.asarray(resource.garfield_gas_gap_map(pos_obs[is_s2]), dtype=np.float64).reshape(-1)
        draw = np.digitize(cont, gaps) - 1
        lo = np.zeros(n, np.int32); hi = np.zeros(n, np.int32); frac = np.zeros(n)
        lo[is_s2] = np.where(draw < 0, draw + n_rows, draw)
        hi[is_s2] = np.clip(draw + 1, 0, n_rows - 1)
        frac[is_s2] = (cont - gaps[draw]) / (gaps[1] - gaps[0])
        out['gg_lo_row'], out['gg_hi_row'], out['gg_frac'] = lo, hi, frac
    # patterns: constant maps share one row per signal type
    s1m, s2m = getattr(resource, 's1_pattern_map', None), getattr(resource, 's2_pattern_map', None)
    rows, row_of = [], np.zeros(n, np.int32)
    aft_sigma = config.get('s2_aft_sigma', 0.0)

    def add_rows(mask, m, is_s2_map):
        if not mask.any():
            return
        smear = is_s2_map and aft_sigma != 0
        if is_s2_map and diffuse:
            # the average over the electrons' displaced positions needs the electron counts, which are
            # drawn on the device: only a device-resident grid can serve it
            if smear or not _on_device(m, 2, n_ch, False):
                raise NotImplementedError('diffusion_transverse_map needs s2_pattern_map as a regular 2-D grid '
                                          '(resource.GridMap) and no s2_aft_sigma smearing')
            row_of[mask] = -1
            return
        if device_patterns and not smear and _on_device(m, 2 if is_s2_map else 3, n_ch, not is_s2_map):
            row_of[mask] = -1        # evaluated on the device from the uploaded grid
            return
        if isinstance(m, DummyMap) and not smear:
            pat = np.asarray(m(np.zeros((1, 2))), dtype=np.float64).reshape(1, -1)
            idx = np.zeros(mask.sum(), np.int64)
        else:
            # s2_pattern_map_diffuse (s2.py:560-613) evaluates the map at the undiffused position
            # unless the transverse-diffusion field map is enabled (handled above)
            pat = np.asarray(m(pos_obs[mask] if is_s2_map else xyz[mask]), dtype=np.float64)
            pat = pat.reshape(mask.sum(), -1)
            idx = np.arange(mask.sum())
        if is_s2_map and pat.shape[1] < n_ch:        # top-only S2 map: s2.py:642-644
            pat = np.pad(pat, [[0, 0], [0, n_ch - pat.shape[1]]], 'constant', constant_values=1)
        if smear:                                    # s2.py:660-665
            from .philox import skewnorm
            n_top = int(config['n_top_pmts'])
            pat = pat.copy()
            pat[:, np.asarray(config['gains']) == 0] = 0
            tot = pat.sum(axis=1, keepdims=True)
            pat = np.divide(pat, tot, out=np.zeros_like(pat), where=tot != 0)
            cur = pat[:, :n_top].sum(axis=1) / pat.sum(axis=1)
            new = np.clip(cur * skewnorm(seed, rng_id[mask], 1.0, aft_sigma, config.get('s2_aft_skewness', 0.0)),
                          0, 1)
            pat[:, :n_top] *= (new / cur)[:, None]
            pat[:, n_top:] *= ((1 - new) / (1 - cur))[:, None]
        base = sum(len(r) for r in rows)
        rows.append(pat.astype(np.float32))
        row_of[mask] = base + idx
    add_rows(is_s1, s1m, False)
    add_rows(is_s2, s2m, True)
    pattern = np.concatenate(rows) if rows else np.ones((1, n_ch), np.float32)
    out.update(s1_lce=s1_lce, s2_sc_gain=sc_gain, s2_cy_extra=cy_extra,
               pattern=np.ascontiguousarray(pattern), pattern_row=row_of)
    return out
