// Device-side random variate generation on top of Philox4x32-10 (philox.cuh).
#pragma once
#include "philox.cuh"
#include <math.h>

namespace wfs {

// A tiny per-thread stream: words of philox(seed, stream, idx, draw0 + k), k = 0, 1, ...
struct Rng {
    uint64_t seed, idx;
    uint32_t stream, draw;
    Philox4 buf;
    int pos;
    __device__ __forceinline__ Rng(uint64_t seed_, uint32_t stream_, uint64_t idx_, uint32_t draw0)
        : seed(seed_), idx(idx_), stream(stream_), draw(draw0), pos(4) {}
    __device__ __forceinline__ uint32_t u32() {
        if (pos == 4) {
            buf = philox4x32(seed, stream, idx, draw++);
            pos = 0;
        }
        return buf.v[pos++];
    }
    // (0,1) open interval, 32-bit resolution
    __device__ __forceinline__ float uf() { return ((float)(u32() >> 8) + 0.5f) * (1.0f / 16777216.0f); }
    // [0,1), 32-bit resolution, as double
    __device__ __forceinline__ double ud32() { return u01_32(u32()); }
    // [0,1), 53-bit
    __device__ __forceinline__ double ud53() {
        uint32_t a = u32(), b = u32();
        return u01_53(a, b);
    }
};

// standard normal pair by Box-Muller (fp32 is ample: every use is truncated to integer ns)
__device__ __forceinline__ void normal_pair(uint32_t a, uint32_t b, float &z0, float &z1) {
    float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

// standard exponential
__device__ __forceinline__ float exp1(uint32_t a) {
    float u = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    return -__logf(u);
}

// double-precision normal for the few large-mean uses (electron drift time)
__device__ __forceinline__ double normal_d(Rng &r) {
    double u1 = 1.0 - r.ud53();   // (0,1]
    double u2 = r.ud53();
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// Exact inversion, visiting the support outward from the mode: value k is returned with
// probability pmf(k) because [0,1) is partitioned into consecutive segments of length pmf(k) in
// visiting order.  Expected iterations ~ 1.6 sigma.
__device__ inline int64_t sample_binomial(int64_t n, double p, double u) {
    if (n <= 0 || p <= 0.0) return 0;
    if (p >= 1.0) return n;
    const double q = 1.0 - p;
    int64_t m = (int64_t)floor((double)(n + 1) * p);
    if (m > n) m = n;
    double pm = exp(lgamma((double)n + 1.0) - lgamma((double)m + 1.0) - lgamma((double)(n - m) + 1.0) +
                    (double)m * log(p) + (double)(n - m) * log1p(-p));
    if (u < pm) return m;
    u -= pm;
    int64_t lo = m, hi = m;
    double plo = pm, phi = pm;
    const double r = p / q, ri = q / p;
    for (int it = 0; it < 100000000; it++) {
        bool moved = false;
        if (hi < n) {
            phi *= (double)(n - hi) / (double)(hi + 1) * r;
            hi++;
            if (u < phi) return hi;
            u -= phi;
            moved = true;
        }
        if (lo > 0) {
            plo *= (double)lo / (double)(n - lo + 1) * ri;
            lo--;
            if (u < plo) return lo;
            u -= plo;
            moved = true;
        }
        if (!moved || (phi < 1e-300 && plo < 1e-300)) break;
    }
    return m;
}

__device__ inline int64_t sample_poisson(double lam, double u) {
    if (!(lam > 0.0)) return 0;
    int64_t m = (int64_t)floor(lam);
    double pm = exp((double)m * log(lam) - lam - lgamma((double)m + 1.0));
    if (u < pm) return m;
    u -= pm;
    int64_t lo = m, hi = m;
    double plo = pm, phi = pm;
    for (int it = 0; it < 100000000; it++) {
        bool moved = false;
        {
            phi *= lam / (double)(hi + 1);
            hi++;
            if (u < phi) return hi;
            u -= phi;
            if (phi > 1e-300) moved = true;
        }
        if (lo > 0) {
            plo *= (double)lo / lam;
            lo--;
            if (u < plo) return lo;
            u -= plo;
            moved = true;
        }
        if (!moved) break;
    }
    return m;
}

}  // namespace wfs
