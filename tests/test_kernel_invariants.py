"""CPU: numpy mirrors of two index tricks the kernels of round 2 rely on, checked against their plain definitions --
the log-spaced time bins of the fused back end's record order (csrc/fused.cu: time_bin) must be monotone in the record
time and fit the bin array of every size class, and a search inside a guide-table cell (csrc/frontend_kernels.cuh:
k_pattern_cdf / interp_table) must return the same upper bound as the search over the whole row (the channel of a
photon, s1.py:148-158 / s2.py:646-677, and the luminescence inverse CDF, s2.py:317-378)."""
import numpy as np
import pytest


def time_bin(t, m):
    t = np.asarray(t, dtype=np.int64)
    k = np.floor(np.log2(np.maximum(t, 1))).astype(np.int64)
    log_bin = ((k - m + 1) << m) + ((t >> np.maximum(k - m, 0)) & ((1 << m) - 1))
    return np.where(t < (1 << m), t, log_bin)


@pytest.mark.parametrize('m', [6, 7, 8])
def test_time_bins_are_monotone_and_fit(m):
    t = np.unique(np.r_[np.arange(0, 1 << 14), np.random.default_rng(m).integers(0, 1 << 21, 200000), (1 << 21) - 1])
    b = time_bin(t, m)
    assert (np.diff(b) >= 0).all() and b[0] == 0
    assert b.max() == ((22 - m) << m) - 1                    # fused_bins(m): the bin array of the class
    assert (b[t < (1 << m)] == t[t < (1 << m)]).all()        # one sample wide behind the group's first photon
    # bins never get wider than 2^-m of the time: at most 2^(k - m) samples in a bin of octave k
    counts = np.bincount(time_bin(np.arange(1 << 16), m))
    assert counts.max() == max(1, (1 << 15) >> m)


@pytest.mark.parametrize('cells,n', [(256, 494), (1024, 2000)])
def test_search_inside_a_guide_cell_equals_the_full_search(cells, n):
    rng = np.random.default_rng(cells)
    p = rng.exponential(1.0, n) * (rng.random(n) > 0.1)      # some dead channels: flat stretches of the CDF
    cdf = np.cumsum(p) / p.sum()
    guide = np.searchsorted(cdf, np.arange(cells + 1) / cells, side='right')       # first entry above j / cells
    u = np.r_[rng.random(200000), np.arange(cells) / cells, np.nextafter(np.arange(1, cells + 1) / cells, 0)]
    cell = (u * cells).astype(np.int64)
    lo, hi = guide[cell], guide[cell + 1]
    full = np.searchsorted(cdf, u, side='right')
    assert (lo <= full).all() and (full <= hi).all()
    inside = np.array([a + np.searchsorted(cdf[a:b], x, side='right') for a, b, x in zip(lo[:5000], hi[:5000], u[:5000])])
    assert (inside == full[:5000]).all()
