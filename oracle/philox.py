"""TEST INFRASTRUCTURE ONLY -- Philox4x32-10 (Salmon et al., SC'11) in plain Python integers, written
independently of wfsim_b200/csrc/philox.cuh and wfsim_b200/philox.py so that the tests can predict the
draws of the CUDA path that are deterministic functions of physics quantities (the noise start offset of
a digitisation group, rawdata.py:407-417, keyed by the first sample of the group's window).
Counter = (index low word, index high word, draw, stream); key = the 64-bit seed."""

RS_NOISE = 1
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_M32 = 0xffffffff


def philox4x32(seed, stream, idx, draw=0):
    idx &= 0xffffffffffffffff
    c = [idx & _M32, idx >> 32, draw & _M32, stream & _M32]
    k0, k1 = seed & _M32, (seed >> 32) & _M32
    for _ in range(10):
        p0, p1 = _M0 * c[0], _M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & _M32, (p0 >> 32) ^ c[3] ^ k1, p0 & _M32]
        k0, k1 = (k0 + _W0) & _M32, (k1 + _W1) & _M32
    return c


def noise_offset(seed, first_sample, span, noise_len):
    """Start index into the noise array for a digitisation group whose window starts at absolute sample
    `first_sample` (= min pulse left - trigger_window) and spans `span` = right - left samples:
    randint(0, high) of rawdata.py:407-417 as floor(u64 * high / 2^64)."""
    high = noise_len - span - 1
    if high < 0:
        high = noise_len - 1
    if high <= 0:
        return 0
    w = philox4x32(seed, RS_NOISE, first_sample & 0xffffffffffffffff)
    return (((w[0] << 32) | w[1]) * high) >> 64
