"""Compact record transport (csrc/transport.cuh): the host expander against a numpy restatement of
the record layout of strax_interface.py:425-436.  No GPU needed; the GPU side (k_pack<true>) is
checked in test_gpu_deterministic.py by running both transports on the same photons."""
import ctypes as C

import numpy as np
import pytest

from wfsim_b200 import lib as wlib
from wfsim_b200.dtypes import raw_record_dtype

HDR = np.dtype([('time', np.int64), ('pulse_length', np.int32), ('channel', np.int16), ('record_i', np.int16),
                ('boff', np.uint32), ('mask', np.uint32)])


def compact(records, fill):
    """numpy restatement of what k_pack<true> emits for `records`."""
    n = len(records)
    hdr = np.zeros(n, HDR)
    for k in ('time', 'pulse_length', 'channel', 'record_i'):
        hdr[k] = records[k]
    data = np.zeros((n, 112), np.int16)
    data[:, :110] = records['data']
    expect = np.where(np.arange(112)[None, :] < records['length'][:, None], fill, 0).astype(np.int16)
    expect[:, 110:] = 0
    differs = (data != expect).reshape(n, 28, 4).any(axis=2)
    hdr['mask'] = (differs * (1 << np.arange(28))[None, :]).sum(axis=1)
    counts = differs.sum(axis=1)
    # the block stream may hold the records' runs in any order: shuffle them
    order = np.random.default_rng(5).permutation(n)
    off = np.zeros(n, np.int64)
    off[order] = np.concatenate([[0], np.cumsum(counts[order])[:-1]])
    hdr['boff'] = off
    blocks = np.zeros((int(counts.sum()) + 1, 4), np.int16)
    for j in range(n):
        blocks[off[j]:off[j] + counts[j]] = data[j].reshape(28, 4)[differs[j]]
    return hdr, blocks


def random_records(n, seed, fill):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, raw_record_dtype())
    r['time'] = rng.integers(-2**40, 2**60, n)
    r['dt'] = 10
    r['channel'] = rng.integers(0, 800, n)
    r['pulse_length'] = rng.integers(1, 10**6, n)
    r['record_i'] = rng.integers(0, r['pulse_length'] // 110 + 1)    # strax_interface.py:427-433
    r['length'] = np.clip(r['pulse_length'] - 110 * r['record_i'].astype(np.int64), 0, 110)
    d = np.full((n, 110), fill, np.int16)
    hit = rng.random((n, 110)) < rng.choice([0.0, 0.02, 0.3, 1.0], n)[:, None]
    d[hit] = rng.integers(0, 16384, int(hit.sum()))
    d[np.arange(110)[None, :] >= r['length'][:, None]] = 0
    r['data'] = d
    return r


@pytest.mark.parametrize('n,threads,misalign', [(0, 1, 0), (1, 1, 0), (63, 1, 4), (64, 1, 8), (1000, 1, 12),
                                                (5000, 4, 0), (70001, 3, 4)])
def test_expander_matches_record_layout(n, threads, misalign):
    lib = wlib.load()
    fill = 16000
    rec = random_records(n, seed=n + threads, fill=fill)
    hdr, blocks = compact(rec, fill)
    buf = np.full(n * 244 + 64, 0x5a, np.uint8)
    dst = buf[misalign:misalign + n * 244]
    rc = lib.wfs_expand_compact(hdr.ctypes.data, blocks.ctypes.data, n, dst.ctypes.data, fill, 10, threads)
    assert rc == 0
    assert dst.tobytes() == rec.tobytes()
    # nothing outside the destination range was touched
    assert (buf[:misalign] == 0x5a).all() and (buf[misalign + n * 244:] == 0x5a).all()
