"""wfsim_b200.instructions against the reference's instruction generator (tests/golden/rand_instructions.npz: the
unmodified strax_interface._rand_instructions / rand_instructions run with a fixed-yield nestpy stand-in):
deterministic columns exactly, random ones by two-sample KS / chi-square at p > 0.01, and the relations between
columns that the generator guarantees (S1 and S2 row of an event share everything but type and amp)."""
import os

import numpy as np

from tests.conftest import GOLDEN
from tests.stat_helpers import P_MIN, discrete_p, ks_p
from wfsim_b200.dtypes import instruction_dtype
from wfsim_b200.instructions import _rand_instructions, fixed_yields, rand_instructions

IDT = np.dtype(instruction_dtype)


def gold():
    z = np.load(os.path.join(GOLDEN, 'rand_instructions.npz'))
    return z, z['inst'].view(IDT), z['inst_config'].view(IDT)


def test_rand_instructions_matches_reference_distributions():
    z, want, _ = gold()
    kw = eval(str(z['kwargs']))
    got = _rand_instructions(yields=fixed_yields, seed=11, **kw)
    assert got.dtype == want.dtype and len(got) == len(want)
    for name in ('time', 'event_number', 'type', 'local_field', 'g4id', 'vol_id'):
        assert np.array_equal(got[name], want[name]), name
    s1g, s1w = got[got['type'] == 1], want[want['type'] == 1]
    assert ks_p(np.hypot(s1g['x'], s1g['y']) ** 2, np.hypot(s1w['x'], s1w['y']) ** 2) > P_MIN       # r^2 uniform
    assert ks_p(np.arctan2(s1g['y'], s1g['x']), np.arctan2(s1w['y'], s1w['x'])) > P_MIN
    assert ks_p(s1g['z'], s1w['z']) > P_MIN
    assert ks_p(s1g['e_dep'], s1w['e_dep']) > P_MIN
    assert discrete_p(s1g['amp'], s1w['amp']) > P_MIN
    assert discrete_p(got['amp'][got['type'] == 2], want['amp'][want['type'] == 2]) > P_MIN
    assert discrete_p(s1g['recoil'], s1w['recoil']) > P_MIN
    for inst in (got, want):                   # the two rows of an event
        a, b = inst[0::2], inst[1::2]
        assert (a['type'] == 1).all() and (b['type'] == 2).all()
        for name in ('time', 'x', 'y', 'z', 'x_pri', 'y_pri', 'z_pri', 'e_dep', 'recoil', 'event_number'):
            assert np.array_equal(a[name], b[name]), name
        assert np.array_equal(a['amp'], np.floor(45 * a['e_dep']).astype(a['amp'].dtype))
        assert np.array_equal(b['amp'], np.floor(28 * b['e_dep']).astype(b['amp'].dtype))
        assert (np.hypot(inst['x'], inst['y']) <= kw['tpc_radius'] * (1 + 1e-6)).all()
        assert ((inst['z'] <= 0) & (inst['z'] >= -kw['tpc_length'])).all()


def test_rand_instructions_from_config_and_seed():
    z, _, want = gold()
    c = eval(str(z['config']))
    got = rand_instructions(c, yields=fixed_yields, seed=3)
    assert len(got) == len(want)
    for name in ('time', 'event_number', 'type', 'local_field', 'recoil'):
        assert np.array_equal(got[name], want[name]), name
    assert (np.hypot(got['x'], got['y']) <= c['tpc_radius']).all() and (got['z'] >= -c['tpc_length']).all()
    assert ((got['e_dep'] >= 1) & (got['e_dep'] <= 100)).all()
    # reproducible per seed, different across seeds
    again = rand_instructions(c, yields=fixed_yields, seed=3)
    other = rand_instructions(c, yields=fixed_yields, seed=4)
    assert again.tobytes() == got.tobytes() and other.tobytes() != got.tobytes()
