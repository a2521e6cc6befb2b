"""CPU: host-side map arithmetic against the unmodified reference (golden) and scipy."""
import os

import numpy as np
import pytest

from tests.conftest import GOLDEN, load_c0_config
from tests.golden import synth_maps as SM
from wfsim_b200 import resource as R


def test_gridmap_matches_scipy_regular_grid_interpolator():
    from scipy.interpolate import RegularGridInterpolator
    m = SM.fdc_3d_map()
    axes, v = m.grid('map')
    rgi = RegularGridInterpolator(tuple(np.linspace(lo, hi, n) for lo, hi, n in axes), v,
                                  bounds_error=False, fill_value=None)
    pts = np.random.default_rng(0).uniform([-70, -70, -170], [70, 70, 20], (500, 3))   # incl. extrapolation
    assert np.allclose(m(pts), rgi(pts), rtol=1e-12, atol=1e-12)
    m1 = SM.s2_optical_spline()
    axes, v = m1.grid('top')
    rgi = RegularGridInterpolator((np.linspace(*axes[0]),), v, bounds_error=False, fill_value=None)
    u = np.random.default_rng(1).random((100, 1))
    assert np.allclose(m1(u, map_name='top'), rgi(u), rtol=1e-12)


def test_field_distortion_matches_reference():
    g = np.load(os.path.join(GOLDEN, 'stoch_models.npz'))
    x, y, z = g['fd_xyz']

    class Res:
        fdc_3d = SM.fdc_3d_map()
        fd_comsol = SM.fd_comsol_map()
    zo, po = R.inverse_field_distortion_correction(x, y, z, Res)
    assert np.array_equal(np.stack([po[:, 0], po[:, 1], zo]), g['fd_inverse_fdc'])
    zo, po = R.field_distortion_comsol(x, y, z, Res)
    assert np.array_equal(np.stack([po[:, 0], po[:, 1], zo]), g['fd_comsol'])


def test_instruction_maps_field_models():
    from tests.golden.make_golden_stoch import fixed_rows
    from wfsim_b200.dtypes import instruction_dtype
    g = np.load(os.path.join(GOLDEN, 'stoch_models.npz'))
    fd = SM.FieldDependencies()
    efd = dict(survival_probability_map=True, drift_speed_map=True, diffusion_longitudinal_map=True,
               diffusion_transverse_map=False)
    cfg = load_c0_config(enable_field_dependencies=efd, field_distortion_model='comsol')
    res = R.Resource(cfg, field_dependencies_map=fd.field_dependencies_map,
                     diffusion_longitudinal_map=fd.diffusion_longitudinal_map, fd_comsol=SM.fd_comsol_map())
    rows = fixed_rows(np.dtype(instruction_dtype), 2, 500, 3, -80.0)
    rows['x'], rows['y'] = 20.0, 15.0
    m = R.evaluate_instruction_maps(cfg, res, rows)
    v, d = m['drift_velocity'][0], m['diffusion_long'][0]
    mean = 80.0 / v + cfg['drift_time_gate']
    spread = np.sqrt(2 * d * mean) / v
    assert np.allclose([mean, spread], g['fdep_mean_spread'], rtol=1e-12)      # s2.py:157-179
    r_obs = np.hypot(m['x_obs'][0], m['y_obs'][0])
    assert np.isclose(r_obs, 25.0 * (1 - 0.05 * 80.0 / 150.0), rtol=1e-9)


def test_transverse_diffusion_sigmas_match_reference_displacements():
    """diffusion_transverse_map (s2.py:560-613): the per-instruction sigmas handed to the device are the
    spreads of the displacements the unmodified reference draws (tests/golden/stoch_diffuse.npz,
    make_golden_diffuse.py: positions read back through an identity "pattern map")."""
    from scipy import stats
    from tests.golden.make_golden_diffuse import CASES
    from tests.golden.make_golden_stoch import fixed_rows
    from wfsim_b200.dtypes import instruction_dtype
    g = np.load(os.path.join(GOLDEN, 'stoch_diffuse.npz'))
    fd = SM.FieldDependencies()
    efd = dict(survival_probability_map=False, drift_speed_map=True, diffusion_longitudinal_map=False,
               diffusion_transverse_map=True)
    cfg = load_c0_config(enable_field_dependencies=efd, diffusion_constant_transverse=1.0)
    n_top = int(cfg['n_top_pmts'])
    res = R.Resource(cfg, field_dependencies_map=fd.field_dependencies_map, s2_pattern_map=SM.s2_pattern_grid(n_top))
    for name, (x, y, z) in CASES.items():
        rows = fixed_rows(np.dtype(instruction_dtype), 2, 40, 2, z)
        rows['x'], rows['y'] = x, y
        rows['type'][1] = 1                     # an S1 row gets no sigma
        rows['z'][1] = -10.0
        m = R.evaluate_instruction_maps(cfg, res, rows)
        assert m['pattern_row'][0] == -1        # averaged on the device from the grid
        assert m['hdiff_sigma_r'][1] == 0 and m['hdiff_sigma_a'][1] == 0
        if name == 'edge':
            continue                            # truncated by the tpc_radius cut in the golden sample
        for key, sig in (('radial', m['hdiff_sigma_r'][0]), ('azimuthal', m['hdiff_sigma_a'][0])):
            assert stats.kstest(g[f'diff_{name}_{key}'].astype(np.float64) / sig, 'norm').pvalue > 0.01, (name, key)
    # off without the config constant (s2.py:636-640), and not available for maps that are not grids
    m = R.evaluate_instruction_maps(dict(cfg, diffusion_constant_transverse=0), res, rows)
    assert 'hdiff_sigma_r' not in m
    res2 = R.Resource(cfg, field_dependencies_map=fd.field_dependencies_map,
                      s2_pattern_map=lambda p, **kw: np.ones((len(p), n_top)))
    with pytest.raises(NotImplementedError):
        R.evaluate_instruction_maps(cfg, res2, rows)


def test_gas_gap_warping_map_values():
    """enable_gas_gap_warping with the 'simple' luminescence model: per-instruction gas gap from
    resource.gas_gap_length at the observed position and the field scale E0 of s2.py:365-370; the field
    scalars replace the constant-gap table in wfs_tables."""
    from tests.golden.make_golden_stoch import fixed_rows
    from wfsim_b200 import params, tables
    from wfsim_b200.dtypes import instruction_dtype
    cfg = load_c0_config(enable_gas_gap_warping=True)
    res = R.Resource(cfg, gas_gap_length=SM.GasGapLength())
    rows = fixed_rows(np.dtype(instruction_dtype), 2, 40, 3, -30.0)
    rows['x'], rows['y'] = [3.0, 20.0, 0.0], [-4.0, 22.5, 0.0]
    rows['type'][2] = 1
    m = R.evaluate_instruction_maps(cfg, res, rows)
    assert np.allclose(m['lum_gap'], [0.215 + 0.075 * 25 / 2500, 0.215 + 0.075 * (20 ** 2 + 22.5 ** 2) / 2500, 0])
    dG = m['lum_gap'][:2]
    VG = cfg['anode_voltage'] / (1 + (cfg['gate_to_anode_distance'] - dG) / dG / cfg['lxe_dielectric_constant'])
    rA, rW = cfg['anode_field_domination_distance'], cfg['anode_wire_radius']
    assert np.allclose(m['lum_e0'][:2], VG / ((dG - rA) / rA + np.log(rA / rW)), rtol=1e-14) and m['lum_e0'][2] == 0
    with pytest.raises(RuntimeError, match='gas_gap_length'):
        R.evaluate_instruction_maps(cfg, R.Resource(cfg), rows)
    t = params.build_tables(cfg, res)
    assert t.struct.lum_len == 0 and t.struct.lumw_dr == 0.0001 and t.struct.lumw_ra == rA and t.struct.lumw_rw == rW
    # constant gap: E0 of the table builder is the same formula
    t0 = params.build_tables(load_c0_config(), R.Resource(load_c0_config()))
    assert t0.struct.lum_len > 0 and t0.struct.lumw_dr == 0
    assert 'lum_gap' not in R.evaluate_instruction_maps(load_c0_config(), R.Resource(load_c0_config()), rows)


def test_resource_from_reference_object():
    """The deployment adapter copies the reference Resource's attributes (load_resource.py:176-380):
    dummy maps keep sharing one pattern row, other map objects are used as callables, tables travel."""
    from tests.golden.make_golden_stoch import fixed_rows
    from wfsim_b200.dtypes import instruction_dtype

    class DummyMap:                       # the reference's own class has this name and these attributes
        def __init__(self, const, shape=()):
            self.const, self.shape = const, shape

        def __call__(self, x, **kw):
            return np.ones([len(x)] + list(self.shape)) * self.const

    class RefResource:
        pass
    cfg = load_c0_config()
    ref = RefResource()
    ref.s1_pattern_map = DummyMap(14e-5, [494])
    ref.s2_pattern_map = lambda pos, **kw: np.tile(np.linspace(1, 2, 494), (len(pos), 1))
    ref.s1_lce_correction_map = DummyMap(14e-5 * 494, [1])
    ref.s2_correction_map = DummyMap(1.0, [1])
    ref.noise_data = np.zeros((100, 494))
    ref.gas_gap_length = SM.GasGapLength()
    ref.drift_velocity_scaling = 0.97
    res = R.resource_from_reference(ref, cfg)
    assert isinstance(res.s1_pattern_map, R.DummyMap) and not isinstance(res.s2_pattern_map, R.DummyMap)
    assert res.noise_data is ref.noise_data and res.gas_gap_length is ref.gas_gap_length
    assert res.drift_velocity_scaling == 0.97
    rows = fixed_rows(np.dtype(instruction_dtype), 2, 40, 4, -30.0)
    rows['type'][::2] = 1
    m = R.evaluate_instruction_maps(cfg, res, rows)
    # S1 rows share the dummy row, every S2 row got its own host-evaluated row
    assert len(set(m['pattern_row'][::2])) == 1 and len(set(m['pattern_row'][1::2])) == 2
    assert np.allclose(m['pattern'][m['pattern_row'][1]], np.linspace(1, 2, 494))
    assert np.allclose(m['s1_lce'][::2], 14e-5 * 494)
    from wfsim_b200.strax_interface import resource_from_reference
    assert resource_from_reference is R.resource_from_reference
