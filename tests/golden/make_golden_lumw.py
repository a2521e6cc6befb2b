"""Golden samples of the 'simple' S2 luminescence model with enable_gas_gap_warping (s2.py:317-378 with
resource.gas_gap_length), drawn by the UNMODIFIED reference with the synthetic gas-gap map of
tests/golden/synth_maps.py:  python tests/golden/make_golden.py lumw"""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden import synth_maps as SM

HERE = os.path.dirname(os.path.abspath(__file__))
N = 60_000
# gas gaps 0.2158, 0.2426, 0.2812 cm
POSITIONS = {'a': [3.0, -4.0], 'b': [20.0, 22.5], 'c': [-42.0, 21.0]}


def main(ref, c0_config):
    out = {}
    RL.seed_reference_rngs(1357)
    cfg, _, _ = c0_config(enable_gas_gap_warping=True)
    res = ref.load_resource.load_config(dict(cfg, enable_gas_gap_warping=False))
    res.gas_gap_length = SM.GasGapLength()
    # each position in an S2 call of its own ...
    for name, xy in POSITIONS.items():
        t = ref.S2.luminescence_timings_simple(np.array([xy]), np.array([N]), cfg, res)
        out['lumw_' + name] = t.astype(np.int16)
    # ... and 'a' together with 'c' in one call: the radial grid starts at the larger gap of the two, which
    # moves the mean that is subtracted from the times of 'a' (s2.py:372-373, 329-331)
    t = ref.S2.luminescence_timings_simple(np.array([POSITIONS['a'], POSITIONS['c']]), np.array([N, N]), cfg, res)
    out['lumw_a_with_c'], out['lumw_c_with_a'] = t[:N].astype(np.int16), t[N:].astype(np.int16)
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.mean(v)), float(np.std(v)), v.min(), v.max())
    np.savez_compressed(os.path.join(HERE, 'stoch_lumw.npz'), **out)
