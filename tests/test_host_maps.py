"""CPU: host-side map arithmetic against the unmodified reference (golden) and scipy."""
import os

import numpy as np

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
