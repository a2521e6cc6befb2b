"""Synthetic regular-grid maps / tables standing in for the XENON resource files that are not
public (optical-propagation splines, garfield luminescence table, field maps).  The same objects
are injected into the unmodified reference (golden generation) and into wfsim_b200 (tests)."""
import numpy as np

from wfsim_b200.resource import GridMap


def s1_optical_spline():
    z = np.linspace(-150.0, 0.0, 16)
    u = np.linspace(0.0, 1.0, 65)
    zz, uu = np.meshgrid(z, u, indexing='ij')
    top = 4.0 + 60.0 * uu ** 2 * (1.0 + (-zz) / 150.0)
    bottom = 2.0 + 35.0 * uu * (1.0 + (150.0 + zz) / 300.0)
    return GridMap([(-150.0, 0.0, 16), (0.0, 1.0, 65)], dict(top=top, bottom=bottom))


def s2_optical_spline():
    u = np.linspace(0.0, 1.0, 129)
    return GridMap([(0.0, 1.0, 129)], dict(top=3.0 + 40.0 * u ** 3, bottom=5.0 + 70.0 * u ** 2))


def garfield_table():
    rng = np.random.default_rng(2024)
    x = np.linspace(-0.25, 0.25, 11)
    t = np.stack([np.sort(rng.exponential(120.0 + 600.0 * abs(xx), 400)) + 200.0 * abs(xx) for xx in x])
    return dict(x=x, t=np.round(t).astype(np.int64))


def fdc_3d_map():
    """r-correction dr(x, y, z) [cm] on a coarse 3-D grid."""
    ax = [(-60.0, 60.0, 13), (-60.0, 60.0, 13), (-155.0, 5.0, 9)]
    g = [np.linspace(lo, hi, n) for lo, hi, n in ax]
    X, Y, Z = np.meshgrid(*g, indexing='ij')
    R = np.sqrt(X ** 2 + Y ** 2)
    return GridMap(ax, dict(map=0.02 * R * (-Z) / 150.0))


def fd_comsol_map():
    ax = [(0.0, 70.0, 15), (-155.0, 5.0, 9)]
    g = [np.linspace(lo, hi, n) for lo, hi, n in ax]
    R, Z = np.meshgrid(*g, indexing='ij')
    return GridMap(ax, dict(r_distortion_map=R * (1.0 - 0.05 * (-Z) / 150.0)))


class FieldDependencies:
    """resource.field_dependencies_map / diffusion_longitudinal_map take (z, xy) (load_resource.py:335-338)."""

    def __init__(self):
        ax = [(0.0, 70.0, 8), (-155.0, 5.0, 9)]
        g = [np.linspace(lo, hi, n) for lo, hi, n in ax]
        R, Z = np.meshgrid(*g, indexing='ij')
        self.m = GridMap(ax, dict(drift_speed_map=0.9 + 0.5 * (-Z) / 150.0 + 0.002 * R,      # mm/us
                                  survival_probability_map=np.clip(1.05 - 0.004 * R, 0, 1.2),
                                  diffusion=(20.0 + 0.2 * R + 0.05 * (-Z)) * 1e-9,           # cm^2/ns
                                  # transverse diffusion [cm^2/s], far larger than in xenon so that the
                                  # smearing is visible on the coarse synthetic pattern grid
                                  diffusion_radial_map=3000.0 + 20.0 * R + 4.0 * (-Z),
                                  diffusion_azimuthal_map=800.0 + 5.0 * R + 1.0 * (-Z)))

    def field_dependencies_map(self, z, xy, map_name):
        r = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
        return self.m(np.array([r, z]).T, map_name=map_name)

    def diffusion_longitudinal_map(self, z, xy):
        r = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
        return self.m(np.array([r, z]).T, map_name='diffusion')


def garfield_gas_gap_table():
    """resource.s2_luminescence_gg stand-in (load_resource.py:286-290): inverse CDFs of the excitation
    time [ns] on 10 gas gaps 0.1 mm apart (s2.py:411-483); the last entries are the 'strange tail' the
    reference does not sample from."""
    gas_gap = np.round(np.linspace(0.24, 0.33, 10), 3)
    u = np.linspace(0.0, 1.0, 203)
    rows = [(300.0 + 4000.0 * (g - 0.24)) * (-np.log(1.0 - 0.995 * u)) ** 0.8 + 50.0 * g for g in gas_gap]
    return dict(gas_gap=gas_gap, timing_inv_cdf=np.stack(rows))


class GasGapMap:
    """resource.garfield_gas_gap_map stand-in: gas gap [cm] growing with the radius."""

    def __call__(self, xy, **kw):
        xy = np.asarray(xy, dtype=np.float64)
        return 0.243 + 0.085 * np.hypot(xy[:, 0], xy[:, 1]) / 70.0


def s2_pattern_grid(n_top, seed=4):
    """S2 hit-pattern map over (x, y), top array only (the layout of the XENONnT map; the bottom array is
    padded with ones, s2.py:642-644): strongly peaked values so that neighbouring cells differ a lot."""
    rng = np.random.default_rng(seed)
    return GridMap([(-55, 55, 12), (-55, 55, 11)], {'map': 40.0 * rng.random((12, 11, n_top)) ** 4 + 0.04})


class ReferencePatternMap:
    """What the reference reads of straxen.InterpolatingMap: a callable with .data['map'] (s2.py:601-604)."""

    def __init__(self, grid_map):
        self.g = grid_map
        self.data = {'map': grid_map.maps['map']}

    def __call__(self, points, **kw):
        return self.g(points, **kw)


class GasGapLength:
    """resource.gas_gap_length stand-in (load_resource.py:319-321, the sagging anode: `lookup(x, y)` of the
    warping map): gas gap [cm] shrinking towards the centre of the TPC."""

    def __call__(self, xy, **kw):
        xy = np.asarray(xy, dtype=np.float64)
        return 0.215 + 0.075 * (np.hypot(xy[:, 0], xy[:, 1]) / 50.0) ** 2
