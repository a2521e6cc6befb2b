"""Counter-based Philox4x32-10 on the host (numpy, vectorised) -- the same function as
csrc/philox.cuh: every draw is philox(seed, stream, entity index, draw index), so host-side draws
(the area-fraction-top smearing of the S2 pattern rows) are keyed by the instruction identity like the
device-side ones and do not depend on how a run is cut into pieces or GPU shards."""
import numpy as np

RS_AFT = 9          # per S2 instruction: skew-normal factor of s2.py:660-665 (streams 1..8: philox.cuh)

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xffffffff)


def philox4x32(seed, stream, idx, draw=0):
    """Four uint32 words per entry of `idx` (uint64 array): shape [4, len(idx)]."""
    idx = np.asarray(idx, dtype=np.uint64).reshape(-1)
    n = len(idx)
    c0 = idx & _MASK
    c1 = idx >> np.uint64(32)
    c2 = np.full(n, int(draw) & 0xffffffff, np.uint64)
    c3 = np.full(n, int(stream) & 0xffffffff, np.uint64)
    seed = int(seed) & 0xffffffffffffffff
    k0, k1 = seed & 0xffffffff, seed >> 32
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + _W0) & 0xffffffff, (k1 + _W1) & 0xffffffff
    return np.stack([c0, c1, c2, c3]).astype(np.uint32)


def normal_pair(w):
    """Two independent standard normals per column of the [4, n] word array (Box-Muller on 53 + 32 bits)."""
    u1 = 1.0 - ((w[0].astype(np.uint64) << np.uint64(32) | w[1].astype(np.uint64)) >> np.uint64(11)) \
        * (1.0 / 9007199254740992.0)
    u2 = w[2].astype(np.float64) * (1.0 / 4294967296.0)
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


def skewnorm(seed, ids, loc, scale, a):
    """Skew-normal variates (scipy.stats.skewnorm's law: loc + scale * (d |Z0| + sqrt(1 - d^2) Z1),
    d = a / sqrt(1 + a^2)), one per identity in `ids`."""
    z0, z1 = normal_pair(philox4x32(seed, RS_AFT, ids))
    d = a / np.sqrt(1.0 + a * a)
    return loc + scale * (d * np.abs(z0) + np.sqrt(1.0 - d * d) * z1)
