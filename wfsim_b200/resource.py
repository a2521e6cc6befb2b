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
                     'uniform_to_pmt_ap', 'uniform_to_ele_ap'):
            if name in overrides:
                setattr(self, name, overrides[name])


def _rz_wrapper(m):
    def rz_map(z, xy, **kwargs):            # load_resource.py:335-338
        r = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
        return m(np.array([r, z]).T, **kwargs)
    return rz_map


def evaluate_instruction_maps(config, resource, instructions):
    """Per-instruction map values handed to the device (struct wfs_instr_maps):
    S1 light yield (s1.py:125), S2 secondary-scintillation gain (s2.py:182-209), the survival /
    extraction factor of the electron yield (s2.py:227-252) and the un-normalised PMT patterns
    (s1.py:148, s2.py:637-644)."""
    n = len(instructions)
    n_ch = len(config['gains'])
    typ = instructions['type']
    is_s1 = typ == 1
    is_s2 = ~is_s1
    xyz = np.stack([instructions['x'], instructions['y'], instructions['z']], axis=1).astype(np.float64)
    xy = xyz[:, :2]
    s1_lce = np.ones(n)
    sc_gain = np.zeros(n)
    cy_extra = np.ones(n)
    p_dpe = config['p_double_pe_emision']
    if is_s1.any():
        ly = np.asarray(resource.s1_lce_correction_map(xyz[is_s1]), dtype=np.float64)
        if ly.ndim != 1:
            ly = np.squeeze(ly, axis=-1)
        s1_lce[is_s1] = ly
    if is_s2.any():
        if config.get('field_distortion_model', 'none') not in ('none', None):
            raise NotImplementedError('field distortion models are evaluated by the reference map '
                                      'objects; not wired for the device path yet')
        pos = xy[is_s2]
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
        if config['enable_field_dependencies']['survival_probability_map']:
            ps = np.asarray(resource.field_dependencies_map(
                xyz[is_s2, 2], pos, map_name='survival_probability_map'), dtype=np.float64).reshape(-1)
            cy_extra[is_s2] *= np.clip(ps, 0, 1)
        for k in ('drift_speed_map', 'diffusion_longitudinal_map'):
            if config['enable_field_dependencies'].get(k):
                raise NotImplementedError(f'field dependency {k} is not wired for the device path yet')
    # patterns: constant maps share one row per signal type
    s1m, s2m = getattr(resource, 's1_pattern_map', None), getattr(resource, 's2_pattern_map', None)
    rows, row_of = [], np.zeros(n, np.int32)

    def add_rows(mask, m, pad_bottom):
        if not mask.any():
            return
        if isinstance(m, DummyMap):
            pat = np.asarray(m(np.zeros((1, 2))), dtype=np.float64).reshape(1, -1)
            idx = np.zeros(mask.sum(), np.int64)
        else:
            pat = np.asarray(m(xyz[mask] if not pad_bottom else xy[mask]), dtype=np.float64)
            idx = np.arange(mask.sum())
        if pad_bottom and pat.shape[1] < n_ch:        # top-only S2 map: s2.py:642-644
            pat = np.pad(pat, [[0, 0], [0, n_ch - pat.shape[1]]], 'constant', constant_values=1)
        base = sum(len(r) for r in rows)
        rows.append(pat.astype(np.float32))
        row_of[mask] = base + idx
    add_rows(is_s1, s1m, False)
    add_rows(is_s2, s2m, True)
    pattern = np.concatenate(rows) if rows else np.ones((1, n_ch), np.float32)
    return dict(s1_lce=s1_lce, s2_sc_gain=sc_gain, s2_cy_extra=cy_extra,
                pattern=np.ascontiguousarray(pattern), pattern_row=row_of)
