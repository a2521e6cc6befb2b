"""Golden samples of the 'garfield_gas_gap' S2 luminescence model (s2.py:411-483), drawn by the UNMODIFIED
reference with the synthetic table / map of tests/golden/synth_maps.py:  python tests/golden/make_golden.py gg"""
import os

import numpy as np

from oracle import ref_loader as RL
from tests.golden import synth_maps as SM

HERE = os.path.dirname(os.path.abspath(__file__))
N = 120_000
POSITIONS = {'a': [3.0, -4.0], 'b': [30.5, 41.0], 'c': [-0.3, 0.2]}     # gas gaps 0.249, 0.305, 0.2434


def main(ref, c0_config):
    out = {}
    RL.seed_reference_rngs(4321)
    cfg, _, _ = c0_config(s2_luminescence_model='garfield_gas_gap', s2_time_model='zero_delay',
                          singlet_fraction_gas=1.0, singlet_lifetime_gas=1e-9)
    res = ref.load_resource.load_config(dict(cfg))
    res.s2_luminescence_gg = SM.garfield_gas_gap_table()
    res.garfield_gas_gap_map = SM.GasGapMap()
    for name, xy in POSITIONS.items():
        t = ref.S2.luminescence_timings_garfield_gasgap(np.array([xy]), np.array([N]), resource=res)
        out['gg_' + name] = t.astype(np.int64).astype(np.int32)      # photon_timings applies astype(int64), s2.py:532-533
    # several small instructions: the mean is subtracted per instruction
    n_i, n_ph = 3000, 40
    t = ref.S2.luminescence_timings_garfield_gasgap(np.tile([POSITIONS['b']], (n_i, 1)), np.full(n_i, n_ph), resource=res)
    out['gg_small_sum'] = t.astype(np.int64).reshape(n_i, n_ph).sum(axis=1).astype(np.int32)
    out['gg_small'] = t.astype(np.int64).astype(np.int32)[:60000]
    for k, v in out.items():
        print(k, v.shape, v.dtype, float(np.mean(v)), float(np.std(v)))
    np.savez_compressed(os.path.join(HERE, 'stoch_gg.npz'), **out)
