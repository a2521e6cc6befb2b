"""Throughput of the host expander (wfs_expand_compact) on C1-like compact records; no GPU needed."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from wfsim_b200 import lib as wlib
HDR = np.dtype([('time', np.int64), ('pulse_length', np.int32), ('channel', np.int16), ('record_i', np.int16),
                ('boff', np.uint32), ('mask', np.uint32)])
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
rng = np.random.default_rng(0)
hdr = np.zeros(n, HDR)
hdr['pulse_length'] = 106
start = rng.integers(8, 16, n)
nb = rng.integers(2, 5, n)
hdr['mask'] = ((1 << nb) - 1) << start
hdr['boff'] = np.concatenate([[0], np.cumsum(nb)[:-1]])
blocks = rng.integers(0, 16000, (int(nb.sum()) + 1, 4)).astype(np.int16)
dst = np.zeros(n * 244, np.uint8)
lib = wlib.load()
for th in (1, 2, 4, 8, 12, 16):
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        lib.wfs_expand_compact(hdr.ctypes.data, blocks.ctypes.data, n, dst.ctypes.data, 16000, 10, th)
        best = min(best, time.perf_counter() - t0)
    print(f'threads {th:2d}: {n * 244 / best / 1e9:6.1f} GB/s out, {best / n * 1e9 * th:6.1f} ns/record/thread')
