"""Golden vectors of the transverse-diffusion S2 hit pattern (S2.s2_pattern_map_diffuse, s2.py:560-613, with
enable_field_dependencies.diffusion_transverse_map), drawn by the UNMODIFIED reference with the synthetic
pattern grid and field maps of tests/golden/synth_maps.py:  python tests/golden/make_golden.py diffuse"""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden import synth_maps as SM

HERE = os.path.dirname(os.path.abspath(__file__))
# on a grid node (largest effect of the smearing) / generic / close to the wall (electrons pushed
# beyond tpc_radius are left out of the average)
CASES = {'node': (25.0, 11.0, -80.0), 'generic': (-12.3, 31.7, -50.0), 'edge': (43.0, 22.0, -90.0)}
N_INST, N_ELECTRON = 4000, 40
N_INST_FEW, N_ELECTRON_FEW = 3000, 3


def normalised(pat, cfg):
    """s2.py:642-651: pad the bottom array with ones, zero the turned-off PMTs, normalise every row."""
    n_ch = len(cfg['gains'])
    pat = np.pad(pat, [[0, 0], [0, n_ch - pat.shape[1]]], 'constant', constant_values=1)
    pat[:, np.asarray(cfg['gains']) == 0] = 0
    return pat / pat.sum(axis=1, keepdims=True)


def main(ref, c0_config):
    out = {}
    RL.seed_reference_rngs(2468)
    efd = dict(survival_probability_map=False, drift_speed_map=True, diffusion_longitudinal_map=False,
               diffusion_transverse_map=True)
    cfg, _, _ = c0_config(enable_field_dependencies=efd, diffusion_constant_transverse=1.0)
    res = ref.load_resource.load_config(dict(cfg))
    fd = SM.FieldDependencies()
    res.field_dependencies_map = fd.field_dependencies_map
    res.drift_velocity_scaling = 1.0
    res.s2_pattern_map = SM.ReferencePatternMap(SM.s2_pattern_grid(int(cfg['n_top_pmts'])))
    for name, (x, y, z) in CASES.items():
        for tag, n_i, n_e in (('', N_INST, N_ELECTRON), ('_few', N_INST_FEW, N_ELECTRON_FEW)):
            pat = ref.S2.s2_pattern_map_diffuse(np.full(n_i, n_e), np.full(n_i, z), np.tile([[x, y]], (n_i, 1)),
                                                cfg, res)
            ok = ~np.isnan(pat).any(axis=1)         # no electron inside the TPC: photons get channel -1
            p = normalised(pat[ok], cfg)
            out[f'diff_{name}{tag}_p'] = p.mean(axis=0)
            out[f'diff_{name}{tag}_nan_rows'] = np.array([(~ok).sum(), n_i])
            # the channel whose per-instruction probability fluctuates most against its multinomial noise
            k = int(np.argmax((p.var(axis=0) / np.maximum(p.mean(axis=0), 1e-12))[:int(cfg['n_top_pmts'])]))
            out[f'diff_{name}{tag}_peak'] = np.array([k])
            out[f'diff_{name}{tag}_peak_p'] = p[:, k]
            print(name + tag, 'nan rows', (~ok).sum(), 'peak ch', k, 'mean', p[:, k].mean(), 'std', p[:, k].std())
    # the displacements themselves: a "pattern map" that returns the position, one electron per row
    class Identity:
        data = {'map': np.zeros((1, 1, 2))}

        def __call__(self, xy, **kw):
            return np.asarray(xy, dtype=np.float64)
    res.s2_pattern_map = Identity()
    n_i = 2500            # the reference re-splits the electron list for every row: quadratic in the rows
    for name, (x, y, z) in CASES.items():
        pos = np.concatenate([ref.S2.s2_pattern_map_diffuse(np.ones(n_i, np.int64), np.full(n_i, z),
                                                            np.tile([[x, y]], (n_i, 1)), cfg, res) for _ in range(4)])
        d = pos[~np.isnan(pos).any(axis=1)] - [x, y]
        th = np.arctan2(y, x)
        out[f'diff_{name}_radial'] = (d[:, 0] * np.cos(th) + d[:, 1] * np.sin(th)).astype(np.float32)
        out[f'diff_{name}_azimuthal'] = (-d[:, 0] * np.sin(th) + d[:, 1] * np.cos(th)).astype(np.float32)
        print(name, 'kept', len(d), 'sigma radial', out[f'diff_{name}_radial'].std(), 'azimuthal', out[f'diff_{name}_azimuthal'].std())
    np.savez_compressed(os.path.join(HERE, 'stoch_diffuse.npz'), **out)
