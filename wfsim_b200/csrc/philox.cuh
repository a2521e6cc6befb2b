// Counter-based Philox4x32-10 (Salmon et al., SC'11).  Every random quantity on the path is a pure
// function of (seed, stream, entity index, draw index), so results do not depend on batching,
// sharding over GPUs or launch geometry -- and a chunk can be recomputed in isolation.
#pragma once
#include <stdint.h>

namespace wfs {

struct Philox4 {
    uint32_t v[4];
};

__host__ __device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// counter = (idx_lo, idx_hi, draw, stream); key = seed
__host__ __device__ __forceinline__ Philox4 philox4x32(uint64_t seed, uint32_t stream, uint64_t idx,
                                                        uint32_t draw) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), draw, stream};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 out;
    out.v[0] = c[0]; out.v[1] = c[1]; out.v[2] = c[2]; out.v[3] = c[3];
    return out;
}

// RNG streams (the `stream` word of the counter)
enum : uint32_t {
    RS_NOISE = 1,       // per digitisation group: noise offset (rawdata.py:417)
    RS_INSTR = 2,       // per instruction: binomial yields (s1.py:133, s2.py:254)
    RS_ELECTRON = 3,    // per electron: trap/drift time, photon count (s2.py:280-281,308-309)
    RS_PHOTON = 4,      // per photon: channel, timing terms, TTS, DPE, SPE
    RS_AP = 5,          // per parent photon: PMT afterpulses (afterpulse.py:189-217)
    RS_PI = 6,          // per S2 pulse call: photo-ionisation (afterpulse.py:37-59)
    RS_PE = 7,          // per S2 pulse call: photo-electric (gate) electrons (afterpulse.py:105-135)
    RS_HDIFF = 8,       // per electron: transverse displacement for the hit pattern (s2.py:588-589)
};

// uniform in [0,1) with 53 bits from two words
__host__ __device__ __forceinline__ double u01_53(uint32_t a, uint32_t b) {
    uint64_t x = (((uint64_t)a << 32) | b) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);
}

// uniform in [0,1) with 32 bits
__host__ __device__ __forceinline__ double u01_32(uint32_t a) {
    return (double)a * (1.0 / 4294967296.0);
}

}  // namespace wfs
