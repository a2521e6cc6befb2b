"""wfsim_b200.optical against the unmodified reference (tests/golden/optical_adjustment.npz, produced by
tests/golden/make_golden_optical.py from wfsim/utils.py:121-165): instructions, photon times and channels byte
for byte, including the split-off entries and the order the reference's swaps leave the photons in."""
import os

import numpy as np

from tests.conftest import GOLDEN
from wfsim_b200.dtypes import instruction_dtype, optical_extra_dtype
from wfsim_b200.optical import PULSE_MAX_DURATION, optical_adjustment


def test_optical_adjustment_equals_reference():
    z = np.load(os.path.join(GOLDEN, 'optical_adjustment.npz'))
    idt = np.dtype(instruction_dtype + optical_extra_dtype)
    n_split = 0
    for k in range(int(z['n_cases'])):
        inst = z[f'in_inst_{k}'].view(idt).copy()
        t, ch = z[f'in_t_{k}'].copy(), z[f'in_ch_{k}'].copy()
        got = optical_adjustment(inst, t, ch)
        want = z[f'out_inst_{k}'].view(idt)
        assert len(got) == len(want), k
        assert got.tobytes() == want.tobytes(), k
        assert np.array_equal(t, z[f'out_t_{k}']) and np.array_equal(ch, z[f'out_ch_{k}']), k
        n_split += len(got) - len(z[f'in_inst_{k}'].view(idt))
        # what the function promises: the entries that were not split off hold no photon later than the limit,
        # the photons of every original entry are still the same multiset
        n0 = len(z[f'in_inst_{k}'].view(idt))
        for row in got[:n0]:
            assert (t[row['_first']:row['_last']] <= PULSE_MAX_DURATION).all()
        assert np.array_equal(np.sort(ch), np.sort(z[f'in_ch_{k}']))
    assert n_split > 0


def test_optical_adjustment_empty_and_short_entries():
    idt = np.dtype(instruction_dtype + optical_extra_dtype)
    inst = np.zeros(3, idt)
    inst['time'] = [100, 200, 300]
    inst['_first'], inst['_last'] = [0, 2, 2], [2, 2, 5]
    t = np.array([50, 40, 7, 9, 8], np.int64)
    ch = np.arange(5, dtype=np.int64)
    out = optical_adjustment(inst, t, ch)
    assert len(out) == 3 and list(out['time']) == [140, 199, 307]        # an empty entry moves by -1, as in the reference
    assert list(t) == [10, 0, 0, 2, 1]
