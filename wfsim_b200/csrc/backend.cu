// Deterministic back end of the WFSim hot path on B200 (sm_100a).
//
//   photons (t_ns, channel, gain, pulse call) resident in HBM
//     -> ordered by (digitisation group, channel, pulse call, time)      [per-group sort in shared memory
//                                                                          or device radix sort, primitives.cu]
//     -> pulses  = runs of equal (group, channel, pulse call)            Pulse.__call__ pulse.py:82-144
//     -> windows = runs of equal (group, channel)                        rawdata.py:231-235,258-259
//     -> k_digitize: one warp per 1024-sample tile of the window-contiguous dense buffer: template
//        superposition in fp64 with the reference's summation order, one rounding per pulse, integer
//        sum over pulses, noise, baseline, clamp, int16 store + ZLE flag bits
//                                                        pulse.py:276-318, rawdata.py:236-272,392-458
//     -> k_zle: hysteresis interval finding on the flag bits             utils.py:13-58, rawdata.py:296-308
//     -> record keys (class, time, channel) -> per-(data type, group) sort in shared memory or
//        device radix sort                                               strax.sort_by_time
//     -> k_pack: 244-byte raw_records written at their final position, or the compact transport
//        form for host destinations (transport.cuh)                      strax_interface.py:425-436
//
// HBM-bound integer/byte work: no tensor cores.  Accumulation is fp64 (B200 has full-rate FP64
// FMA pipes; the mul/add are issued unfused to match the reference bit for bit).
#include "backend.cuh"
#include "philox.cuh"

#include <algorithm>
#include <limits.h>
#include <stdlib.h>

namespace wfs {

constexpr int kBlk = 8;                 // samples per digitize thread

enum Scalar {
    S_NVALID = 0, S_NPULSES, S_NWIN, S_NTILES, S_NITVSLOTS, S_NREC, S_MINSAMPLE, S_MAXSAMPLE,
    S_ERR, S_NITV, S_NSAMPLES, S_CLASS0, S_CLASS1, S_CLASS2, S_NBLOCKS, S_RECSEG_MAX, S_GROUPS_OVERLAP, S_DENSE_TILES, S_COUNT
};

struct WinMeta {
    int64_t left;        // absolute sample index of the first sample of the window
    int32_t len;         // samples
    int32_t channel;     // output channel (he_first + ch for HE rows)
    int32_t group;
    int32_t mult;        // 1, or he_mult for HE rows
    int32_t p0, p1;      // pulses [p0, p1) contributing to this window
    uint32_t ph_lo, ph_hi;   // their photons [ph_lo, ph_hi) (contiguous in the sorted arrays)
    uint32_t blk0;           // first 8-sample block of the window in the dense buffer
    uint32_t slot0;          // first interval slot
    int32_t s_first, s_last; // window-local sample range any photon of the window can touch
    int32_t pad0, pad1;
};

struct Interval {
    int64_t left;        // absolute sample index
    int64_t src;         // offset of the first sample in the dense buffer
    int32_t len;         // pulse_length
    int32_t channel;
};

struct KeyLayout {
    int bits_rank, bits_group;
    int shift_rank, shift_ch, shift_group;
    int total_bits;
};

__device__ __forceinline__ int64_t floordiv(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_init(int64_t *group_tmin, int64_t *group_lr, uint32_t *group_nitv,
                       int64_t n_groups, int64_t *scalars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_groups) {
        group_tmin[i] = LLONG_MAX;
        group_lr[2 * i] = LLONG_MAX;
        group_lr[2 * i + 1] = LLONG_MIN;
        group_nitv[i] = 0;
    }
    if (i < S_COUNT) {
        int64_t v = 0;
        if (i == S_MINSAMPLE) v = LLONG_MAX;
        if (i == S_MAXSAMPLE) v = LLONG_MIN;
        scalars[i] = v;
    }
}

__device__ __forceinline__ int32_t pc_of(const PhotonBatch &b, int64_t i) {
    int32_t v = b.pulse_call[i];
    if (b.instr_run) {
        int32_t run = b.instr_run[v];
        return run < 0 ? -1 : 2 * run + ((b.flags[i] >> 1) & 1);
    }
    return v;
}

__device__ __forceinline__ bool photon_valid(int32_t ch, int32_t pc, const DeviceConfig &c,
                                             int64_t n_pc) {
    // dead PMTs are skipped by Pulse.__call__ (pulse.py:89-90); channel -1 = "no pattern" (s2.py:670)
    return ch >= 0 && ch < c.p.n_tpc_pmts && pc >= 0 && pc < n_pc && c.gains[ch] != 0.0;
}

__global__ void k_group_tmin(PhotonBatch b, DeviceConfig c, int64_t *group_tmin, uint32_t *group_nvalid) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool valid = false;
    int g = -1;
    int64_t t = LLONG_MAX;
    if (i < b.n) {
        int32_t ch = b.channel[i], pc = pc_of(b, i);
        valid = photon_valid(ch, pc, c, b.n_pulse_calls);
        if (valid) {
            g = b.pc_group[pc];
            t = b.t[i];
        }
    }
    // Photons come instruction by instruction, so a warp -- mostly a whole CTA -- sees one group: the minimum on the
    // 32-bit reduce unit (high word, then low word), one atomic per CTA through shared memory.  A warp with several
    // groups takes one atomic per distinct group.
    __shared__ int64_t s_mn[32];
    __shared__ int s_g[32], s_cnt[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    const int g0 = __shfl_sync(0xffffffffu, g, vm ? __ffs(vm) - 1 : 0);
    const bool uniform = __all_sync(0xffffffffu, !valid || g == g0);
    int wg = -1, wcnt = 0;
    int64_t wmn = LLONG_MAX;
    if (uniform) {
        if (vm) {
            const int32_t hi = (int32_t)(t >> 32);
            const int32_t mh = __reduce_min_sync(0xffffffffu, hi);
            const uint32_t ml = __reduce_min_sync(0xffffffffu, hi == mh ? (uint32_t)t : 0xffffffffu);
            wmn = (int64_t)(((uint64_t)(uint32_t)mh << 32) | ml);
            wg = g0; wcnt = __popc(vm);
        }
    } else {
        const unsigned m = __match_any_sync(0xffffffffu, g);
        int64_t mn = t;
        for (int o = 0; o < 32; o++) {
            int64_t other = __shfl_sync(0xffffffffu, t, o);
            if ((m >> o) & 1u) mn = other < mn ? other : mn;
        }
        if (valid && (m & ((1u << lane) - 1u)) == 0) {
            atomicMin((long long *)&group_tmin[g], (long long)mn);
            if (group_nvalid) atomicAdd(&group_nvalid[g], (uint32_t)__popc(m));   // invalid lanes carry g = -1
        }
    }
    if (lane == 0) { s_g[warp] = wg; s_mn[warp] = wmn; s_cnt[warp] = wcnt; }
    __syncthreads();
    if (warp == 0) {
        // runs of warps with the same group: their first warp's lane sends the atomic for the run
        const int mg = lane < n_warps ? s_g[lane] : -1;
        const unsigned same = __match_any_sync(0xffffffffu, mg);
        if (mg >= 0 && (same & ((1u << lane) - 1u)) == 0) {
            int64_t mn = LLONG_MAX;
            uint32_t cnt = 0;
            for (unsigned rest = same; rest; rest &= rest - 1) {
                const int w = __ffs(rest) - 1;
                mn = s_mn[w] < mn ? s_mn[w] : mn;
                cnt += (uint32_t)s_cnt[w];
            }
            atomicMin((long long *)&group_tmin[mg], (long long)mn);
            if (group_nvalid) atomicAdd(&group_nvalid[mg], cnt);
        }
    }
}

// keys behind the valid photons of a segment-sorted batch read as "invalid" (they sort last)
__global__ void k_fill_invalid_tail(int64_t n, const uint32_t *n_valid, uint64_t invalid_key, uint64_t *keys) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && i >= (int64_t)*n_valid) keys[i] = invalid_key;
}

__global__ void k_build_keys(PhotonBatch b, DeviceConfig c, KeyLayout kl, const int64_t *group_tmin,
                             uint64_t *keys, uint32_t *vals, int64_t *scalars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    int32_t ch = b.channel[i], pc = pc_of(b, i);
    uint64_t key;
    if (photon_valid(ch, pc, c, b.n_pulse_calls)) {
        int g = b.pc_group[pc];
        int64_t rel = b.t[i] - group_tmin[g];
        if (rel >= (int64_t(1) << kRelTimeBits)) {
            scalars[S_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
            rel = (int64_t(1) << kRelTimeBits) - 1;
        }
        key = (uint64_t)rel | ((uint64_t)b.pc_rank[pc] << kl.shift_rank) |
              ((uint64_t)ch << kl.shift_ch) | ((uint64_t)g << kl.shift_group);
    } else {
        key = (uint64_t)b.n_groups << kl.shift_group;   // sorts behind every valid photon
    }
    keys[i] = key;
    vals[i] = (uint32_t)i;
}

// After the sort: gather time/gain into sorted order and flag pulse / window starts.
__global__ void k_gather_flags(PhotonBatch b, KeyLayout kl, const uint64_t *keys,
                               const uint32_t *vals, int64_t *st, double *sg, uint64_t *flags,
                               uint8_t *pstart, int64_t *scalars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n) return;
    uint64_t k = keys[i];
    bool valid = (k >> kl.shift_group) < (uint64_t)b.n_groups;
    uint64_t f = 0;
    if (valid) {
        uint32_t src = vals[i];
        st[i] = b.t[src];
        sg[i] = b.gain[src];
        uint64_t kp = i > 0 ? keys[i - 1] : ~0ull;
        bool newpulse = (i == 0) || ((k >> kl.shift_rank) != (kp >> kl.shift_rank));
        bool newwin = (i == 0) || ((k >> kl.shift_ch) != (kp >> kl.shift_ch));
        f = (newpulse ? 1ull : 0ull) | (newwin ? (1ull << 32) : 0ull);
        bool next_valid = (i + 1 < b.n) && ((keys[i + 1] >> kl.shift_group) < (uint64_t)b.n_groups);
        if (!next_valid) scalars[S_NVALID] = i + 1;
    }
    flags[i] = f;
    pstart[i] = (uint8_t)(f & 1ull);
}

__global__ void k_emit_pulses(int64_t n, KeyLayout kl, const uint64_t *keys, const uint64_t *flags,
                              const uint64_t *pos, uint32_t *pulse_first, uint32_t *pulse_win,
                              uint32_t *win_first_pulse, uint32_t *win_key, int64_t *scalars) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) {
        uint64_t tot = pos[n];
        uint32_t np = (uint32_t)tot, nw = (uint32_t)(tot >> 32);
        scalars[S_NPULSES] = np;
        scalars[S_NWIN] = nw;
        pulse_first[np] = (uint32_t)scalars[S_NVALID];
        win_first_pulse[nw] = np;
        return;
    }
    uint64_t f = flags[i];
    if (f & 1ull) {
        uint64_t e = pos[i];
        uint32_t p = (uint32_t)e, w = (uint32_t)(e >> 32);
        if (!(f >> 32)) w -= 1;
        pulse_first[p] = (uint32_t)i;
        pulse_win[p] = w;
        if (f >> 32) {
            win_first_pulse[w] = p;
            win_key[w] = (uint32_t)(keys[i] >> kl.shift_ch);   // (group << 10) | channel
        }
    }
}

// One thread per base window: pulse extents (pulse.py:118-127), window extents
// (rawdata.py:231-235,258-259), HE twin (rawdata.py:241-249), group extents (rawdata.py:215-216).
__global__ void k_window_extents(int64_t n_win, DeviceConfig c, const int64_t *st,
                                 const uint32_t *pulse_first, const uint32_t *win_first_pulse,
                                 const uint32_t *win_key, int64_t *pulse_left, WinMeta *meta,
                                 uint64_t *win_scan_in, int64_t *group_lr, int64_t *scalars,
                                 int he_rows) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const int dt = c.p.dt;
    uint32_t p0 = win_first_pulse[w], p1 = win_first_pulse[w + 1];
    int64_t lo = LLONG_MAX, hi = LLONG_MIN;
    for (uint32_t p = p0; p < p1; p++) {
        uint32_t a = pulse_first[p], b = pulse_first[p + 1];
        int64_t l = floordiv(st[a], dt) - c.p.pulse_left_margin;
        int64_t r = floordiv(st[b - 1], dt) + c.p.pulse_right_margin;
        pulse_left[2 * (int64_t)p] = l;
        pulse_left[2 * (int64_t)p + 1] = r;
        lo = l < lo ? l : lo;
        hi = r > hi ? r : hi;
    }
    uint32_t key = win_key[w];
    int ch = key & ((1u << kChannelBits) - 1u);
    int g = key >> kChannelBits;
    {   // windows are ordered by (group, channel): aggregate the group extents inside the warp
        const int lane = threadIdx.x & 31;
        const unsigned act = __activemask();
        const unsigned grp = __match_any_sync(act, g);
        int64_t glo = lo, ghi = hi;
        for (int o = 0; o < 32; o++) {
            if (!((act >> o) & 1u)) continue;
            int64_t olo = __shfl_sync(act, lo, o), ohi = __shfl_sync(act, hi, o);
            if ((grp >> o) & 1u) { glo = olo < glo ? olo : glo; ghi = ohi > ghi ? ohi : ghi; }
        }
        if ((grp & ((1u << lane) - 1u)) == 0) {
            atomicMin((long long *)&group_lr[2 * g], (long long)glo);
            atomicMax((long long *)&group_lr[2 * g + 1], (long long)ghi);
        }
    }
    const int tw = c.p.trigger_window;
    WinMeta m;
    m.left = lo - tw;
    int64_t len = hi - lo + 2 * tw + 1;
    if (len > kMaxGroupSamples + 1) {
        scalars[S_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
        len = kMaxGroupSamples + 1;
    }
    m.len = (int32_t)len;
    m.channel = ch;
    m.group = g;
    m.mult = 1;
    m.p0 = (int32_t)p0;
    m.p1 = (int32_t)p1;
    m.ph_lo = pulse_first[p0];
    m.ph_hi = pulse_first[p1];
    m.blk0 = m.slot0 = 0;
    // pulse extents are [q_first - left_margin, q_last + right_margin]: recover the photon samples
    m.s_first = (int32_t)((lo + c.p.pulse_left_margin) - m.left);
    m.s_last = (int32_t)((hi - c.p.pulse_right_margin) - m.left) + c.p.template_length - 1;
    m.pad0 = m.pad1 = 0;
    meta[w] = m;
    const int holdoff = 2 * tw + 1;
    uint64_t tiles = (uint64_t)((len + kBlk - 1) / kBlk);   // 8-sample blocks
    uint64_t icap = (uint64_t)(len / (holdoff + 1) + 2);
    win_scan_in[w] = tiles | (icap << 32);
    // HE twin (rawdata.py:241-249).  With int(high_energy_deamplification_factor) == 0 the row is
    // baseline (+ noise) only; it is materialised only if it can drop below its ZLE threshold.
    WinMeta h = m;
    bool he = he_rows && c.p.detector_nt && ch < c.p.n_top_pmts;
    if (he) {
        const int hch = c.p.he_first + ch;
        const bool he_noise = c.p.enable_noise && c.noise_t != nullptr && hch < c.noise_nch;
        he = c.p.he_mult != 0 || he_noise || c.zle_thr[hch] > c.p.baseline;
    }
    if (he) {
        h.channel = c.p.he_first + ch;
        h.mult = c.p.he_mult;
    } else {
        h.len = 0;
    }
    meta[n_win + w] = h;
    win_scan_in[n_win + w] = he ? (tiles | (icap << 32)) : 0ull;
}

// Truth quirk of Pulse.add_truth (pulse.py:251-255): `trigger_dpe` counts the above-threshold
// photons among the FIRST n_double_pe photons of the channel slice, whichever they are.
// Photons in their sorted (time) order.  One thread per pulse for short pulses; pulses of more than 32
// photons (heavy S2s: thousands per PMT) are handed to the whole warp, one after the other, so that
// no single thread walks them alone.  Everything summed is an integer: the order does not matter.
struct PulseTruth {
    uint32_t ndpe = 0;
    int trig = 0, n_trig = 0;
    int64_t area = 0, area_trig = 0;
};

template <bool kWarp>
__device__ __forceinline__ PulseTruth pulse_truth(uint32_t a, uint32_t e, int ch, bool per_pmt, const PhotonBatch &b,
                                                  const DeviceConfig &c, const uint32_t *vals, const int64_t *st,
                                                  const double *sg) {
    const int lane = kWarp ? (int)(threadIdx.x & 31) : 0, step = kWarp ? 32 : 1;
    PulseTruth r;
    for (uint32_t i = a + lane; i < e; i += step) r.ndpe += b.flags[vals[i]] & 1u;
    if (kWarp) r.ndpe = __reduce_add_sync(0xffffffffu, r.ndpe);
    const double thr = (double)(c.p.baseline - 1 - c.zle_thr[ch]) - 0.5;
    const double gch = c.gains[ch];
    const uint32_t stop = per_pmt ? e : a + r.ndpe;
    for (uint32_t i = a + lane; i < stop; i += step) {
        const int64_t t = st[i];
        const int rem = (int)(t - floordiv(t, c.p.dt) * c.p.dt);
        const bool above = sg[i] * c.current_max[rem] * c.p.current_2_adc > thr;
        if (above && i < a + r.ndpe) r.trig++;
        if (per_pmt) {
            const int64_t ar = llrint(sg[i] / gch * 4294967296.0);   // same fixed point as the totals (k_instr_truth)
            r.area += ar;
            if (above) { r.n_trig++; r.area_trig += ar; }
        }
    }
    if (kWarp) {
        r.trig = __reduce_add_sync(0xffffffffu, r.trig);
        r.n_trig = __reduce_add_sync(0xffffffffu, r.n_trig);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            r.area += __shfl_xor_sync(0xffffffffu, r.area, o);
            r.area_trig += __shfl_xor_sync(0xffffffffu, r.area_trig, o);
        }
    }
    return r;
}

__device__ __forceinline__ void store_pulse_truth(const PulseTruth &r, uint32_t a, uint32_t e, int32_t pc, int ch,
                                                  const PhotonBatch &b, const DeviceConfig &c) {
    if (r.trig) {
        atomicAdd(&b.trig_dpe_out[2 * pc], r.trig);
        if (ch >= c.p.n_top_pmts) atomicAdd(&b.trig_dpe_out[2 * pc + 1], r.trig);
    }
    if (b.pmt_counts) {   // per_pmt_truth: one writer per (pulse call, channel)
        const int64_t npmt = c.p.n_tpc_pmts;
        int32_t *cnt = b.pmt_counts + ((int64_t)(pc >> 1) * 4) * npmt + ch;
        int64_t *ar_out = b.pmt_areas + ((int64_t)(pc >> 1) * 2) * npmt + ch;
        cnt[0] = (int32_t)(e - a);
        cnt[npmt] = (int32_t)(e - a + r.ndpe);
        cnt[2 * npmt] = r.n_trig;
        cnt[3 * npmt] = r.n_trig + r.trig;
        ar_out[0] = r.area;
        ar_out[npmt] = r.area_trig;
    }
}

__global__ void k_truth_pulses(int64_t n_pulses, PhotonBatch b, DeviceConfig c, const uint32_t *vals,
                               const int64_t *st, const double *sg, const uint32_t *pulse_first,
                               const uint32_t *pulse_win, const uint32_t *win_key) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool per_pmt = b.pmt_counts != nullptr;
    uint32_t a = 0, e = 0;
    int32_t pc = 1;
    int ch = 0;
    if (p < n_pulses) {
        a = pulse_first[p]; e = pulse_first[p + 1];
        pc = pc_of(b, vals[a]);
        ch = win_key[pulse_win[p]] & ((1u << kChannelBits) - 1u);
    }
    // PMT-afterpulse calls (odd ids) carry preset gains: n_double_pe = 0 (pulse.py:106); their truth
    // buffer is not the one get_truth reads (rawdata.py:317-318)
    const bool mine = p < n_pulses && !(pc & 1);
    const bool big = mine && e - a > 32u;
    if (mine && !big) store_pulse_truth(pulse_truth<false>(a, e, ch, per_pmt, b, c, vals, st, sg), a, e, pc, ch, b, c);
    unsigned todo = __ballot_sync(0xffffffffu, big);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const uint32_t a2 = __shfl_sync(0xffffffffu, a, src), e2 = __shfl_sync(0xffffffffu, e, src);
        const int ch2 = __shfl_sync(0xffffffffu, ch, src);
        const int32_t pc2 = __shfl_sync(0xffffffffu, pc, src);
        const PulseTruth r = pulse_truth<true>(a2, e2, ch2, per_pmt, b, c, vals, st, sg);
        if (lane == 0) store_pulse_truth(r, a2, e2, pc2, ch2, b, c);
    }
}

// Noise start offset per group (rawdata.py:407-417) when not supplied by the caller.
__global__ void k_group_noise(int64_t n_groups, DeviceConfig c, const int64_t *group_lr,
                              const int64_t *ix_in, uint64_t seed,
                              int64_t *ix_out, int64_t *scalars) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    int64_t lo = group_lr[2 * g], hi = group_lr[2 * g + 1];
    if (lo != LLONG_MAX) {
        int64_t left = lo - c.p.trigger_window, right = hi + c.p.trigger_window;
        if (right - left >= kMaxGroupSamples) scalars[S_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
        atomicMin((long long *)&scalars[S_MINSAMPLE], (long long)left);
        atomicMax((long long *)&scalars[S_MAXSAMPLE], (long long)(right + 1));
    }
    int64_t ix = 0;
    if (ix_in) {
        ix = ix_in[g];
    } else if (c.p.enable_noise && c.noise_len > 0 && lo != LLONG_MAX) {
        int64_t span = hi - lo + 2 * c.p.trigger_window;
        int64_t high = c.noise_len - span - 1;
        if (high < 0) high = c.noise_len - 1;
        if (high > 0) {
            // keyed by the first sample of the group's window: unique per group (groups are disjoint in
            // time) and independent of how the run is cut into calls, device batches or GPU shards
            Philox4 r = philox4x32(seed, RS_NOISE, (uint64_t)(lo - c.p.trigger_window), 0);
            uint64_t u = ((uint64_t)r.v[0] << 32) | r.v[1];
            ix = (int64_t)__umul64hi(u, (uint64_t)high);
        }
    }
    ix_out[g] = ix;
}

// ---------------------------------------------------------------------------------------------
// k_digitize: persistent WARPS, each looping over TILES of the dense buffer (up to 128 consecutive
// 8-sample blocks = 1024 samples; the dense buffer is window-contiguous, so a tile covers a few
// whole windows or a slice of a long one).  A warp runs a tile from start to finish with
// __syncwarp only -- no CTA barrier after the template load -- so ~28 independent tiles are in
// flight per SM and the dependent metadata / photon loads of one hide behind the others.
//   setup   the windows overlapping the tile -> W.win (ballot-ordered), photon prefix by shuffles;
//   sparse  (<= kStageCap photons in those windows, the common case): one lane per photon
//           stages it in shared memory (tile-local sample position, merged gain) and marks the
//           8-sample blocks its template reaches.  An ISLAND is a run of photons of one Pulse
//           call whose 22-tap templates overlap; every sample of an island is OWNED by the first
//           photon whose template covers it, and the owner's index is scattered into byte l of
//           the sample's word (layer l = pulse ordinal in the window mod 4, so overlapping Pulse
//           calls do not collide; a collision is detected by the atomicOr and sends the tile to
//           the dense path).  Then one thread per sample of the touched blocks: for each layer,
//           walk the island from the owner and sum template x merged gain in time order exactly
//           like Pulse.add_current (pulse.py:301-318, mul and add unfused), round once per Pulse
//           call (rawdata.py:236-239) and add the integers.
//   dense   (many photons): one thread per 8 samples gathers the photons that reach it, per
//           pulse, with a binary search for the first one (all lanes busy: high photon density);
//   finish  untouched blocks (nothing but the baseline) are one constant 16-byte store; touched
//           blocks: noise, baseline, clamp (rawdata.py:398-458), int16 pack, ZLE flag bits
//           (sample < threshold, rawdata.py:290-296).
// k_zle: ONE THREAD per window runs the hysteresis interval search (utils.py:13-58,
// rawdata.py:296-308) on the flag bytes.
// ---------------------------------------------------------------------------------------------
constexpr int kDigiThreads = 128;
constexpr int kDigiWarps = kDigiThreads / 32;
constexpr int kDigiCtasPerSm = 7;
constexpr int kNeg = -(1 << 29);
constexpr int kTileBlkMax = 128;                 // 8-sample blocks per tile (one tile per warp)
constexpr int kTileSmpMax = kTileBlkMax * kBlk;
constexpr int kTileWinMax = 16;                  // windows overlapping one tile (host sizes the tile)
constexpr int kStageCap = 127;                   // photons staged in shared memory per tile (byte index + 1)
constexpr int kLayers = 4;
constexpr int kTmplMax = 352;                    // dt * template_length doubles staged per CTA (checked at wfs_create)

struct TileWin {
    int32_t off;             // tile-local sample index of window sample 0 (negative: starts before)
    int32_t len;
    int32_t mult;
    int32_t thr;
    int32_t channel;         // -1: no noise row
    int32_t ixbase;          // noise start index of window sample 0, already modulo noise_len
    uint32_t ph_lo, ph_cnt;
    int32_t p0, p1;
    int32_t s_first, s_last;
    uint32_t stage0;
    int32_t all_touched;     // every block needs the full finish (noise, or baseline below threshold)
};

// For every tile the window owning its first block, so no CTA needs a search.  One thread per window.
__global__ void k_tile_index(int64_t n_wtot, int tile_blk, const uint64_t *__restrict__ win_off,
                             WinMeta *meta, uint32_t *tile_first) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_wtot) return;
    const int64_t b0 = (int64_t)(uint32_t)win_off[w], b1 = (int64_t)(uint32_t)win_off[w + 1];
    meta[w].blk0 = (uint32_t)b0;
    meta[w].slot0 = (uint32_t)(win_off[w] >> 32);
    if (b1 <= b0) return;
    for (int64_t g = (b0 + tile_blk - 1) / tile_blk; g * tile_blk < b1; g++) tile_first[g] = (uint32_t)w;
}

// Per sorted photon, packed into 32 bits: ns remainder [0,4), window-local sample index [4,24),
// ordinal of its Pulse call inside the window [24,29) (31 = 31 or more), flags (island start,
// run head, alone = an island of one photon); and -- for the
// first photon of a run of equal-ns photons of one pulse -- the merged gain (pulse.py:301-318:
// gains of coincident photons are summed before the template multiply).  kPhIsland marks the first
// photon of an island: a new Pulse call, or a photon whose template cannot overlap the previous one's.
constexpr uint32_t kPhAlone = 1u << 31, kPhRunHead = 1u << 30, kPhIsland = 1u << 29,
                   kPhQMask = (1u << 20) - 1u, kPhOrdShift = 24, kPhOrdMask = 31u;
static_assert(kMaxGroupSamples + 1 <= (int64_t)kPhQMask, "window-local sample index must fit");

__global__ void k_photon_prep(int64_t n_valid, DeviceConfig c, const int64_t *__restrict__ st,
                              const double *__restrict__ sg, const uint8_t *__restrict__ pstart,
                              const uint64_t *__restrict__ pos, const uint32_t *__restrict__ pulse_win,
                              const WinMeta *__restrict__ meta, uint32_t *__restrict__ phq,
                              double *__restrict__ gm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_valid) return;
    const int64_t t = st[i];
    const bool ps = pstart[i] != 0 || i == 0;
    const int64_t tprev = st[i - (i > 0)];
    const bool head = ps || tprev != t;
    const uint32_t p = (uint32_t)pos[i] + (pstart[i] != 0 ? 1u : 0u) - 1u;
    const WinMeta *m = meta + pulse_win[p];
    const int64_t left = m->left;
    const uint32_t ord = min(p - (uint32_t)m->p0, kPhOrdMask);
    const int64_t q = floordiv(t, c.p.dt);
    const uint32_t r = (uint32_t)(t - q * c.p.dt);
    uint32_t v = (((uint32_t)(q - left) & kPhQMask) << 4) | r | (ord << kPhOrdShift);
    if (ps || q - floordiv(tprev, c.p.dt) >= c.p.template_length) {
        v |= kPhIsland;
        // single-photon island: the next photon (if any) starts an island of its own
        if (i + 1 >= n_valid || pstart[i + 1] || floordiv(st[i + 1], c.p.dt) - q >= c.p.template_length) v |= kPhAlone;
    }
    double g = 0.0;
    if (head) {
        v |= kPhRunHead;
        g = sg[i];
        for (int64_t j = i + 1; j < n_valid && !pstart[j] && st[j] == t; j++) g = __dadd_rn(g, sg[j]);
    }
    phq[i] = v;
    gm[i] = g;
}

__device__ __forceinline__ int ph_q(uint32_t v) { return (int)((v >> 4) & kPhQMask); }

// dense path: the 8 samples [s0, s0 + 8) of window `m`, gathered per pulse
// `starts` (optional): for every pulse of the window (stride `stride`) and every block of the tile the first photon
// that can reach the block, as the warp tabulated them (k_digitize, dense path); without it every block searches.
__device__ __forceinline__ void gather_block(const TileWin &m, int s0, int tlen, double c2a,
                                             const double *s_tmpl, const uint32_t *__restrict__ phq,
                                             const double *__restrict__ gm,
                                             const uint32_t *__restrict__ pulse_first, int acc[kBlk],
                                             const uint32_t *starts = nullptr, int stride = 0, int blk = 0) {
#pragma unroll
    for (int i = 0; i < kBlk; i++) acc[i] = 0;
    if (m.mult == 0 || s0 + kBlk <= m.s_first || s0 > m.s_last) return;
    const int q_lo = s0 - (tlen - 1), q_hi = s0 + kBlk - 1;   // photon samples that reach mine
    double cur[kBlk];
#pragma unroll
    for (int i = 0; i < kBlk; i++) cur[i] = 0.0;
    for (int p = m.p0; p < m.p1; p++) {
        const uint32_t f0 = pulse_first[p], f1 = pulse_first[p + 1];
        uint32_t a = f0, b = f1;
        if (starts) {
            a = starts[(p - m.p0) * stride + blk];
        } else {
            while (a < b) {   // first photon of the pulse that can reach me (a run head, see k_photon_prep)
                const uint32_t mid = (a + b) >> 1;
                if (ph_q(phq[mid]) < q_lo) a = mid + 1; else b = mid;
            }
        }
        bool any = false;
        // the photon word and gain of the next round are requested before this round's taps are worked on
        // the photon word and gain of the next round are requested before this round's taps are worked on
        // (C2 digitize 15.5 -> 13.6 ms per 100 heavy events; C4 3.4 -> 3.9 ms per 5e7 photons)
        uint32_t v_next = a < f1 ? phq[a] : 0u;
        double g_next = a < f1 ? gm[a] : 0.0;
        for (uint32_t i = a; i < f1; i++) {
            const uint32_t v = v_next;
            const double g = g_next;
            if (i + 1 < f1) { v_next = phq[i + 1]; g_next = gm[i + 1]; }
            if (!(v & kPhRunHead)) continue;
            const int q = ph_q(v);
            if (q > q_hi) break;
            const double *tm = s_tmpl + (v & 15u) * tlen;
            const int first = q - s0;          // my sample index of template tap 0
            any = true;
#pragma unroll
            for (int j = 0; j < kBlk; j++) {
                const int tap = j - first;
                if (tap >= 0 && tap < tlen) cur[j] = __dadd_rn(cur[j], __dmul_rn(tm[tap], g));
            }
        }
        if (any) {   // one rounding per (pulse call, channel): rawdata.py:236-239
#pragma unroll
            for (int j = 0; j < kBlk; j++) {
                acc[j] -= __double2int_rn(__dmul_rn(cur[j], c2a)) * m.mult;
                cur[j] = 0.0;
            }
        }
    }
}

// staged photon: s_rf = ns remainder | run head << 4 | island start << 5 | layer << 6 | window slot << 8 | alone << 13
constexpr uint32_t kRfHead = 16u, kRfIsland = 32u, kRfAlone = 1u << 13;

// noise, baseline, clamp, ZLE flags, one 16-byte store of 8 int16 samples + one flag byte
__device__ __forceinline__ void finish_block(int v[kBlk], const TileWin &m, int S, const DeviceConfig &c,
                                             uint4 *out, uint8_t *flag) {
    if (m.channel >= 0) {
        const double *row = c.noise_t + (int64_t)m.channel * c.noise_len;
        int64_t ix = ((int64_t)m.ixbase + S) % c.noise_len;
#pragma unroll
        for (int j = 0; j < kBlk; j++) {
            v[j] = __double2int_rz((double)v[j] + row[ix]);
            if (++ix >= c.noise_len) ix = 0;
        }
    }
    const int nvalid = min(kBlk, m.len - S), thr = m.thr, baseline = c.p.baseline;
    uint32_t flags = 0, h[kBlk];
#pragma unroll
    for (int j = 0; j < kBlk; j++) {
        int x = max(v[j] + baseline, 0);
        if (x < thr) flags |= 1u << j;
        if (j >= nvalid) x = 0;
        h[j] = (uint32_t)(uint16_t)(int16_t)x;
    }
    flags &= (1u << nvalid) - 1u;
    *out = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
    *flag = (uint8_t)flags;
}

// Per-warp working set of one tile.
struct __align__(16) WarpTile {
    uint32_t own[kTileSmpMax];      // per sample: owner index + 1 of layer l in byte l; then the ADC sum
    double gm[kStageCap + 1];
    int T[kStageCap + 1];
    TileWin win[kTileWinMax];
    uint16_t rf[kStageCap + 1];
    uint8_t tb[kTileBlkMax];        // touched blocks, compacted
    uint8_t blkwin[kTileBlkMax];
    uint32_t touch[kTileBlkMax / 32];
};

__global__ void __launch_bounds__(kDigiThreads)
k_digitize(int64_t n_blocks, int tile_blk, int n_tiles, int64_t n_wtot, DeviceConfig c,
           const WinMeta *__restrict__ meta, const uint64_t *__restrict__ win_off,
           const uint32_t *__restrict__ tile_first, const uint32_t *__restrict__ phq,
           const double *__restrict__ gm, const uint32_t *__restrict__ pulse_first,
           const int64_t *__restrict__ group_ix, int16_t *__restrict__ dense,
           uint8_t *__restrict__ flag8, int64_t *scalars) {
    __shared__ double s_tmpl[kTmplMax];
    __shared__ WarpTile s_wt[kDigiWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int dt = c.p.dt, tlen = c.p.template_length;
    const bool noise_on = c.p.enable_noise && c.noise_t != nullptr;
    const double c2a = c.p.current_2_adc;
    const int base_clamped = max(c.p.baseline, 0);
    const uint32_t base_h = (uint32_t)(uint16_t)(int16_t)base_clamped;
    const uint32_t base_pw = base_h | (base_h << 16);
    WarpTile &W = s_wt[warp];
    for (int i = tid; i < dt * tlen; i += kDigiThreads) s_tmpl[i] = c.templates[i];
    for (int i = lane; i < kTileSmpMax; i += 32) W.own[i] = 0;
    __syncthreads();   // the only CTA-wide barrier: from here on every warp works on tiles of its own

    const int n_warps = gridDim.x * kDigiWarps;
    // the window rows of the NEXT tile of this warp are pulled towards the SM while the current tile is
    // worked on: tile -> first window -> window offsets / rows -> photons is a chain of dependent loads
    int64_t w_first = blockIdx.x * kDigiWarps + warp < n_tiles ? tile_first[blockIdx.x * kDigiWarps + warp] : 0;
    for (int tile = blockIdx.x * kDigiWarps + warp; tile < n_tiles; tile += n_warps) {
        const int64_t w_here = w_first;
        if (tile + n_warps < n_tiles) w_first = tile_first[tile + n_warps];
        const int64_t B0 = (int64_t)tile * tile_blk;
        const int nblk = (int)min((int64_t)tile_blk, n_blocks - B0);
        const int nsmp = nblk * kBlk;
        uint4 *out = reinterpret_cast<uint4 *>(dense + B0 * kBlk);
        __syncwarp();
        if (lane < kTileBlkMax / 32) W.touch[lane] = 0;
        // ---- setup: the windows overlapping this tile, in window order ----
        int n = 0;
        for (int64_t base = w_here;; base += 32) {
            const int64_t w = base + lane;
            bool past = w >= n_wtot, take = false;
            int64_t b0 = 0;
            if (!past) {
                b0 = (int64_t)(uint32_t)win_off[w];
                const int64_t b1 = (int64_t)(uint32_t)win_off[w + 1];
                past = b0 >= B0 + nblk;
                take = !past && b1 > b0 && b1 > B0;
            }
            const uint32_t tm = __ballot_sync(0xffffffffu, take);
            const int slot = n + __popc(tm & lt);
            if (take && slot < kTileWinMax) {
                const WinMeta m = meta[w];
                TileWin t;
                t.off = (int32_t)(b0 - B0) * kBlk;
                t.len = m.len;
                t.mult = m.mult;
                t.thr = c.zle_thr[m.channel];
                const bool noisy = noise_on && m.channel < c.noise_nch;
                t.channel = noisy ? m.channel : -1;
                t.ixbase = noisy ? (int32_t)(group_ix[m.group] % c.noise_len) : 0;
                t.ph_lo = m.ph_lo;
                t.ph_cnt = m.mult != 0 ? m.ph_hi - m.ph_lo : 0u;
                t.p0 = m.p0; t.p1 = m.p1;
                t.s_first = m.s_first; t.s_last = m.s_last;
                t.stage0 = 0;
                t.all_touched = (noisy || base_clamped < t.thr) ? 1 : 0;
                W.win[slot] = t;
            }
            n += __popc(tm);
            if (__ballot_sync(0xffffffffu, past)) break;
        }
        if (n > kTileWinMax) {   // cannot happen: the host sizes tile_blk from the minimum window length
            if (lane == 0) scalars[S_ERR] = WFS_E_ARG;
            n = kTileWinMax;
        }
        __syncwarp();
        // exclusive prefix of the photon counts over the slots
        const uint32_t c0 = lane < n ? min(W.win[lane].ph_cnt, 1u << 24) : 0u;
        uint32_t i0 = c0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a0 = __shfl_up_sync(0xffffffffu, i0, o);
            if (lane >= o) i0 += a0;
        }
        if (lane < n) W.win[lane].stage0 = i0 - c0;
        const int total = (int)__shfl_sync(0xffffffffu, i0, 31);
        const bool mult1 = __ballot_sync(0xffffffffu, lane < n && W.win[lane].mult != 1) == 0;
        const int nwin = n;
        bool sparse = total <= kStageCap;
        __syncwarp();
        if (sparse) {
            // ---- stage the photons (one lane each); mark the blocks their templates reach ----
            bool clash = false;
            if (lane < nwin) {
                const TileWin &m = W.win[lane];
                const int blk_lo = max(m.off / kBlk, 0), blk_end = (m.off + m.len + kBlk - 1) / kBlk;
                if (m.all_touched) {
                    for (int b = blk_lo; b < min(blk_end, nblk); b++) {
                        W.blkwin[b] = (uint8_t)lane;
                        atomicOr(&W.touch[b >> 5], 1u << (b & 31));
                    }
                } else if ((m.len & (kBlk - 1)) && blk_end <= nblk) {   // partial last block
                    W.blkwin[blk_end - 1] = (uint8_t)lane;
                    atomicOr(&W.touch[(blk_end - 1) >> 5], 1u << ((blk_end - 1) & 31));
                }
            }
            for (int i = lane; i < total; i += 32) {
                int slot = 0;   // last slot with stage0 <= i
#pragma unroll
                for (int step = kTileWinMax / 2; step > 0; step >>= 1)
                    if (slot + step < nwin && (int)W.win[slot + step].stage0 <= i) slot += step;
                const uint32_t src = W.win[slot].ph_lo + (uint32_t)(i - (int)W.win[slot].stage0);
                const uint32_t v = phq[src];
                const int off = W.win[slot].off;
                const int T0 = off + ph_q(v);
                const uint32_t ord = (v >> kPhOrdShift) & kPhOrdMask;
                W.gm[i] = gm[src];
                W.T[i] = T0;
                W.rf[i] = (uint16_t)((v & 15u) | ((v & kPhRunHead) ? kRfHead : 0u) |
                                     ((v & kPhIsland) ? kRfIsland : 0u) | ((ord & (kLayers - 1)) << 6) |
                                     ((uint32_t)slot << 8) | ((v & kPhAlone) ? kRfAlone : 0u));
                if (ord == kPhOrdMask) clash = true;
                if (v & kPhRunHead) {
                    const int t_hi = min(T0 + tlen - 1, nsmp - 1);
                    for (int b = max(T0, 0) >> 3; b <= (t_hi >> 3); b++) {
                        W.blkwin[b] = (uint8_t)slot;
                        atomicOr(&W.touch[b >> 5], 1u << (b & 31));
                    }
                    // the samples this photon owns: its template's reach minus the previous photon's
                    int lo = T0;
                    if (!(v & kPhIsland)) lo = max(lo, off + ph_q(phq[src - 1]) + tlen);
                    lo = max(lo, 0);
                    const int sh = (int)(ord & (kLayers - 1)) * 8;
                    for (int T = lo; T <= t_hi; T++) {
                        const uint32_t old = atomicOr(&W.own[T], (uint32_t)(i + 1) << sh);
                        clash |= ((old >> sh) & 0xffu) != 0;   // two Pulse calls on one layer overlap
                    }
                }
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, clash)) {   // rare: clear the owner table and take the dense path
                for (int i = lane; i < kTileSmpMax; i += 32) W.own[i] = 0;
                sparse = false;
                __syncwarp();
            }
        }
        if (sparse) {
            if (tile + n_warps < n_tiles && w_first + lane < n_wtot) {   // next tile's window rows and offsets
                asm volatile("prefetch.global.L2 [%0];" ::"l"(meta + w_first + lane));
                if (lane < 5) asm volatile("prefetch.global.L2 [%0];" ::"l"(win_off + w_first + 8 * lane));
            }
            // ---- untouched blocks: constant stores; touched blocks: compacted list ----
            int ntb = 0;
            for (int wd = 0; wd * 32 < nblk; wd++) {
                const uint32_t word = W.touch[wd];
                const int b = wd * 32 + lane;
                if ((word >> lane) & 1u) {
                    W.tb[ntb + __popc(word & lt)] = (uint8_t)b;
                } else if (b < nblk) {          // nothing but the baseline in these 8 samples
                    out[b] = make_uint4(base_pw, base_pw, base_pw, base_pw);
                    flag8[B0 + b] = 0;
                }
                ntb += __popc(word);
            }
            __syncwarp();
            // ---- touched samples (one lane each): gather from the owners, round per Pulse call ----
            for (int e = lane; e < ntb * kBlk; e += 32) {
                const int b = W.tb[e >> 3];
                const int T = b * kBlk + (e & (kBlk - 1));
                uint32_t ow = W.own[T];
                if (!ow) continue;
                const int mult = mult1 ? 1 : W.win[W.blkwin[b]].mult;
                int v = 0;
                do {   // one owner per layer (byte), usually a single one
                    const int sh = (31 - __clz(ow)) & ~7;
                    const int o = (int)((ow >> sh) & 0xffu);
                    ow &= ~(0xffu << sh);
                    const uint32_t r0 = W.rf[o - 1];
                    double cur;
                    if (r0 & kRfAlone) {
                        cur = __dmul_rn(s_tmpl[(r0 & 15u) * tlen + (T - W.T[o - 1])], W.gm[o - 1]);
                    } else {
                        cur = 0.0;
                        for (int k = o - 1; k < total; k++) {
                            const uint32_t rk = W.rf[k];
                            if (k >= o && (rk & kRfIsland)) break;
                            const int tap = T - W.T[k];
                            if (tap < 0) break;
                            if (rk & kRfHead) cur = __dadd_rn(cur, __dmul_rn(s_tmpl[(rk & 15u) * tlen + tap], W.gm[k]));
                        }
                    }
                    v -= __double2int_rn(__dmul_rn(cur, c2a)) * mult;
                } while (ow);
                W.own[T] = (uint32_t)v;
            }
            __syncwarp();
            // ---- touched blocks: finish ----
            for (int j = lane; j < ntb; j += 32) {
                const int b = W.tb[j];
                const TileWin &m = W.win[W.blkwin[b]];
                int4 *acc = reinterpret_cast<int4 *>(&W.own[b * kBlk]);
                const int4 a0 = acc[0], a1 = acc[1];
                acc[0] = acc[1] = make_int4(0, 0, 0, 0);
                int v[kBlk] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                finish_block(v, m, b * kBlk - m.off, c, &out[b], &flag8[B0 + b]);
            }
        } else {
            // ---- dense path ----
            if (lane == 0) atomicAdd((unsigned long long *)&scalars[S_DENSE_TILES], 1ull);
            for (int slot = 0; slot < nwin; slot++) {
                const int lo = max(W.win[slot].off / kBlk, 0);
                const int hi = min((W.win[slot].off + W.win[slot].len + kBlk - 1) / kBlk, nblk);
                for (int b = lo + lane; b < hi; b += 32) W.blkwin[b] = (uint8_t)slot;
            }
            __syncwarp();
            // Instead of a binary search per block and pulse over photons in L2, the warp reads the photons that reach
            // the tile once per (window, pulse), histograms the first block each of them does NOT precede, and a
            // prefix sum gives every block its first photon.  One table row per (window, pulse) in the owner table.
            const int stride = nblk + 1;
            int rows = 0;
            for (int slot = 0; slot < nwin; slot++)
                if (W.win[slot].mult != 0) rows += W.win[slot].p1 - W.win[slot].p0;
            // (tiles that see several windows keep the search: building their rows costs more than it saves -- C2
            // digitize 13.6 ms per 100 heavy events with this condition, 15.1 without, 16.8 with the search everywhere)
            // (tried on top of the table and dropped: lane = sample, 32 samples per round, every photon a broadcast load
            // -- correct, but 2.5x the instructions of a block per lane: C2 digitize 35.6 instead of 13.6 ms)
            const bool tabulated = nwin == 1 && rows > 0 && rows * stride <= kTileSmpMax;
            if (tabulated) {
                uint32_t *tab = W.own;                  // all zero on entry (the sparse path leaves it so)
                int row = 0;
                for (int slot = 0; slot < nwin; slot++) {
                    const TileWin m0 = W.win[slot];
                    __syncwarp();
                    if (lane == 0) W.win[slot].stage0 = (uint32_t)row;      // (the staging offset is not used on this path)
                    if (m0.mult == 0) continue;
                    const int q_lo_t = -m0.off - (tlen - 1), q_hi_t = nblk * kBlk - 1 - m0.off;
                    for (int p = m0.p0; p < m0.p1; p++, row++) {
                        const uint32_t f0 = pulse_first[p], f1 = pulse_first[p + 1];
                        uint32_t a = f0, b = f1;       // lane 0: first photon with q >= q_lo_t; lane 1: first with q > q_hi_t
                        const int key = lane == 0 ? q_lo_t : q_hi_t + 1;
                        if (lane < 2)
                            while (a < b) {
                                const uint32_t mid = (a + b) >> 1;
                                if (ph_q(phq[mid]) < key) a = mid + 1; else b = mid;
                            }
                        const uint32_t A = __shfl_sync(0xffffffffu, a, 0), Z = __shfl_sync(0xffffffffu, a, 1);
                        uint32_t *t = tab + row * stride;
                        for (uint32_t i = A + lane; i < Z; i += 32) {
                            // the photon precedes block b  <=>  q < 8 b - off - (tlen - 1)  <=>  b >= (q + off + tlen - 1) / 8 + 1
                            const int h = (ph_q(phq[i]) + m0.off + tlen - 1) / kBlk + 1;
                            if (h < stride) atomicAdd(&t[h], 1u);
                        }
                        __syncwarp();
                        // inclusive prefix over the blocks, + A: the first photon of the pulse that can reach block b
                        uint32_t carry = A;
                        for (int b0 = 0; b0 < stride; b0 += 32) {
                            const int bb = b0 + lane;
                            uint32_t x = bb < stride ? t[bb] : 0u;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                                if (lane >= o) x += y;
                            }
                            if (bb < stride) t[bb] = carry + x;
                            carry += __shfl_sync(0xffffffffu, x, 31);
                        }
                    }
                }
                __syncwarp();
            }
            for (int b = lane; b < nblk; b += 32) {
                const TileWin &m = W.win[W.blkwin[b]];
                int v[kBlk];
                gather_block(m, b * kBlk - m.off, tlen, c2a, s_tmpl, phq, gm, pulse_first, v,
                             tabulated ? W.own + m.stage0 * stride : nullptr, stride, b);
                finish_block(v, m, b * kBlk - m.off, c, &out[b], &flag8[B0 + b]);
            }
            if (tabulated) {
                __syncwarp();
                for (int i = lane; i < rows * stride; i += 32) W.own[i] = 0;
            }
        }
    }
}

__device__ __forceinline__ void emit_interval(int s, int e, const WinMeta &m, int tw, int64_t dense0,
                                              Interval *itv, uint32_t *itv_nrec, int64_t slot) {
    int l = s - tw, r = e + tw;
    l = max(0, min(l, m.len - 1));
    r = max(0, min(r, m.len - 1));
    l = (l + 1) & ~1;
    r = r & ~1;
    int plen = r - l + 1;
    if (plen < 0) plen = 0;
    Interval it;
    it.left = m.left + l;
    it.src = dense0 + l;
    it.len = plen;
    it.channel = m.channel;
    itv[slot] = it;
    itv_nrec[slot] = (uint32_t)((plen + WFS_SAMPLES_PER_RECORD - 1) / WFS_SAMPLES_PER_RECORD);
}

__global__ void __launch_bounds__(128)
k_zle(int64_t n_wtot, DeviceConfig c, const WinMeta *__restrict__ meta,
      const uint64_t *__restrict__ win_off, const uint8_t *__restrict__ flag8, Interval *itv,
      uint32_t *itv_nrec, uint32_t *group_nitv, int64_t *scalars) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n_emitted = 0, len = 0;
    if (w < n_wtot) {
        const WinMeta m = meta[w];
        len = m.len;
        if (m.len > 0) {
            const uint64_t off = win_off[w], off1 = win_off[w + 1];
            const int64_t blk0 = (int64_t)(uint32_t)off, slot0 = (int64_t)(off >> 32);
            const int cap = (int)((off1 >> 32) - (off >> 32));
            const int64_t dense0 = blk0 * kBlk;
            const int nblk = (m.len + kBlk - 1) / kBlk;
            const int tw = c.p.trigger_window, H = 2 * tw + 1;
            int last = kNeg, start = kNeg;
            const uint8_t *f = flag8 + blk0;
            for (int b = 0; b < nblk; b++) {
                // most flag bytes are zero (baseline): skip them eight at a time
                if (((reinterpret_cast<uintptr_t>(f + b) & 7) == 0) && b + 8 <= nblk) {
                    const unsigned long long word = *reinterpret_cast<const unsigned long long *>(f + b);
                    if (word == 0ull) {
                        b += 7;
                        continue;
                    }
                    // inside a heavy S2 every sample is below threshold: 64 flagged samples that continue the open
                    // interval only move its end (one thread walks the ~40 000 samples of such a window)
                    if (word == ~0ull && last != kNeg && b * kBlk - last <= H) {
                        last = b * kBlk + 8 * kBlk - 1;
                        b += 7;
                        continue;
                    }
                }
                uint32_t byte = f[b];
                while (byte) {
                    const int pos = b * kBlk + __ffs(byte) - 1;
                    if (last == kNeg) start = pos;
                    else if (pos - last > H) {
                        if (n_emitted < cap) emit_interval(start, last, m, tw, dense0, itv, itv_nrec, slot0 + n_emitted);
                        n_emitted++;
                        start = pos;
                    }
                    // consecutive flagged samples never open an interval: jump to the run's end
                    const uint32_t run = byte | (byte - 1);         // fill below the lowest set bit
                    const uint32_t stop = ~run & (run + 1);         // lowest clear bit above the run
                    const int run_end = stop > 0xffu || stop == 0 ? 7 : __ffs(stop) - 2;
                    last = b * kBlk + run_end;
                    byte &= ~((2u << run_end) - 1u);
                }
            }
            if (last != kNeg) {
                if (n_emitted < cap) emit_interval(start, last, m, tw, dense0, itv, itv_nrec, slot0 + n_emitted);
                n_emitted++;
            }
            if (n_emitted > cap) scalars[S_ERR] = WFS_E_ARG;   // cannot happen: cap is an upper bound
            if (n_emitted) atomicAdd(&group_nitv[m.group], (uint32_t)n_emitted);
            for (int k = n_emitted; k < cap; k++) itv_nrec[slot0 + k] = 0;   // unused slots
        }
    }
    // block-level totals
    unsigned long long ni = n_emitted, ns = len;
    for (int o = 16; o > 0; o >>= 1) {
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
        ns += __shfl_xor_sync(0xffffffffu, ns, o);
    }
    if ((threadIdx.x & 31) == 0 && ns) {
        atomicAdd((unsigned long long *)&scalars[S_NITV], ni);
        atomicAdd((unsigned long long *)&scalars[S_NSAMPLES], ns);
    }
}

__device__ __forceinline__ int channel_class(int ch, const wfs_params &p) {
    // strax_interface.py:490-493
    if (p.detector_nt) {
        if (ch < p.he_first) return 0;
        if (ch <= p.he_last) return 1;
        return 2;
    }
    return 0;
}

struct RecDesc {      // everything k_pack needs for one record, 32 bytes
    int64_t time;     // ns
    int64_t src;      // first sample in the dense buffer
    int32_t pulse_length;
    int32_t length;
    int16_t channel;
    int16_t record_i;
    int32_t pad;
};

// Record segments.  Interval slots follow the window order -- all base windows by (group, channel),
// then all high-energy twins by (group, channel) -- so the records of one (data type, group) are a
// contiguous range of the record list.  When the digitisation windows of consecutive groups do not
// overlap in time (what the scheduler produces), sorting every segment by (time, channel) IS the
// global (class, time, channel) order and can be done per segment in shared memory.
__global__ void k_group_first_window(int64_t n_win, const uint32_t *__restrict__ win_key, uint32_t *group_win) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_win) return;
    const uint32_t g = win_key[w] >> kChannelBits;
    if (w == 0 || (win_key[w - 1] >> kChannelBits) != g) group_win[g] = (uint32_t)w;
}

__global__ void k_rec_segments(int64_t n_groups, int64_t n_win, DeviceConfig c, const uint32_t *__restrict__ group_win,
                               const uint64_t *__restrict__ win_off, const uint32_t *__restrict__ itv_rec0,
                               const int64_t *__restrict__ group_lr, uint32_t *rec_seg, int64_t *scalars) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s > 2 * n_groups) return;
    auto start_of = [&](int64_t seg) -> uint32_t {
        int64_t w = 2 * n_win;                       // behind the last window: the record total
        if (seg < 2 * n_groups) {
            const int64_t cls = seg / n_groups;
            w = (cls + 1) * n_win;
            for (int64_t g = seg - cls * n_groups; g < n_groups; g++)
                if (group_win[g] != 0xffffffffu) { w = cls * n_win + group_win[g]; break; }
        }
        return itv_rec0[win_off[w] >> 32];
    };
    const uint32_t a = start_of(s);
    rec_seg[s] = a;
    if (s == 2 * n_groups) return;
    const uint32_t e = start_of(s + 1);
    if (e > a) atomicMax((long long *)&scalars[S_RECSEG_MAX], (long long)(e - a));
    if (s < n_groups && group_lr[2 * s] != LLONG_MAX) {     // my window against the next non-empty group's
        for (int64_t g = s + 1; g < n_groups; g++) {
            if (group_lr[2 * g] == LLONG_MAX) continue;
            if (group_lr[2 * s + 1] + c.p.trigger_window >= group_lr[2 * g] - c.p.trigger_window)
                scalars[S_GROUPS_OVERLAP] = 1;
            break;
        }
    }
}

// strax_interface.py:425-436 header fields + the (class, time, channel) sort key of
// strax.sort_by_time, one thread per interval slot.
__global__ void k_rec_keys(int64_t n_slots, DeviceConfig c, const Interval *itv,
                           const uint32_t *itv_nrec, const uint32_t *itv_rec0, int64_t min_sample,
                           int time_bits, uint64_t *rec_keys, uint32_t *rec_vals, RecDesc *desc) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    uint32_t n = itv_nrec[s];
    if (!n) return;
    const Interval it = itv[s];
    const uint32_t r0 = itv_rec0[s];
    const int cls = channel_class(it.channel, c.p);
    const int spr = WFS_SAMPLES_PER_RECORD;
    for (uint32_t i = 0; i < n; i++) {
        const int64_t first = it.left + (int64_t)spr * i;
        uint64_t trel = (uint64_t)(first - min_sample);
        rec_keys[r0 + i] = (uint64_t)it.channel | (trel << kChannelBits) |
                           ((uint64_t)cls << (kChannelBits + time_bits));
        rec_vals[r0 + i] = r0 + i;
        RecDesc d;
        d.time = (int64_t)c.p.dt * first;
        d.src = it.src + (int64_t)spr * i;
        d.pulse_length = it.len;
        d.length = min(it.len, spr * (int)(i + 1)) - spr * (int)i;
        d.channel = (int16_t)it.channel;
        d.record_i = (int16_t)i;
        d.pad = 0;
        desc[r0 + i] = d;
    }
}

// records per data type from the sorted keys (class is the top field of the key)
__global__ void k_class_counts(int64_t n_rec, const uint64_t *keys, int class_shift, int64_t *scalars) {
    if (blockIdx.x != 0 || threadIdx.x >= 2) return;
    const uint64_t want = (uint64_t)(threadIdx.x + 1);
    int64_t lo = 0, hi = n_rec;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if ((keys[mid] >> class_shift) < want) lo = mid + 1; else hi = mid;
    }
    // first index of class 1 and of class 2
    scalars[threadIdx.x == 0 ? S_CLASS1 : S_CLASS2] = lo;
}

// Record packing (strax_interface.py:425-436): every WARP assembles 8 records in its own slice of
// shared memory and writes them as one contiguous, 16-byte-vectorised span at their final sorted
// position (244 B = 61 words: 6 header words + 55 data words; 8 records = 122 x 16 B).  Warps never
// wait for each other (no CTA barrier): the dependent chain rec_vals -> descriptor -> samples of one
// warp hides behind the other warps of the SM.
// kCompact: the records leave in the compact transport form instead (transport.cuh): a 24-byte
// header per record at its final sorted position and only the 4-sample blocks that differ from the
// fill pattern (baseline below `length`, zero behind it), appended to a block stream through one
// atomic per warp (the header carries the offset, so the stream order does not matter).
constexpr int kPackWarpRecs = 8, kPackWarps = 4, kPackRecs = kPackWarpRecs * kPackWarps;
// Assemble the records of one warp in shared memory.  kFull: all kPackWarpRecs records exist (every warp
// but the last one) -- no per-record guards, so the code is straight-line with predicated loads; the
// header word of a lane is picked with selects (an if-chain on the lane compiles to an indirect branch).
template <bool kFull>
__device__ __forceinline__ void pack_assemble(int nhere, int lane, int dt, const RecDesc *s_desc,
                                              const int16_t *__restrict__ dense, uint32_t *s_rec) {
    uint32_t v0[kPackWarpRecs], v1[kPackWarpRecs];
#pragma unroll
    for (int r = 0; r < kPackWarpRecs; r++) {
        v0[r] = v1[r] = 0;
        if (kFull || r < nhere) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(dense + s_desc[r].src);
            const int length = s_desc[r].length;
            if (2 * lane < length) v0[r] = src[lane];
            if (lane + 32 < 55 && 2 * (lane + 32) < length) v1[r] = src[lane + 32];
        }
    }
#pragma unroll
    for (int r = 0; r < kPackWarpRecs; r++) {
        if (!kFull && r >= nhere) continue;
        const RecDesc d = s_desc[r];
        const int length = d.length;
        uint32_t a0 = v0[r], a1 = v1[r];
        a0 = 2 * lane + 1 >= length ? a0 & 0xffffu : a0;
        a1 = 2 * (lane + 32) + 1 >= length ? a1 & 0xffffu : a1;
        uint32_t h = (uint32_t)(uint64_t)d.time;
        h = lane == 1 ? (uint32_t)((uint64_t)d.time >> 32) : h;
        h = lane == 2 ? (uint32_t)length : h;
        h = lane == 3 ? (((uint32_t)(uint16_t)dt) | ((uint32_t)(uint16_t)d.channel << 16)) : h;
        h = lane == 4 ? (uint32_t)d.pulse_length : h;
        h = lane == 5 ? (uint32_t)(uint16_t)d.record_i : h;       // record_i, baseline = 0
        uint32_t *o = s_rec + r * 61;
        if (lane < 6) o[lane] = h;
        o[6 + lane] = a0;
        if (lane + 32 < 55) o[6 + 32 + lane] = a1;
    }
}

// Compact transport form of the 8 records a warp holds in shared memory (61 words each, header word 2 =
// length): mask of the 4-sample blocks that differ from the fill pattern, one atomic per warp for the
// block stream, 24-byte headers at the records' positions.
__device__ __forceinline__ void compact_emit(int nhere, int lane, int64_t j0, int baseline, const uint32_t *s_rec,
                                             uint32_t *s_mask, uint32_t *s_off, uint32_t *__restrict__ chdr,
                                             uint2 *__restrict__ cblk, int64_t *scalars) {
    const uint32_t fill_h = (uint32_t)(uint16_t)(int16_t)max(baseline, 0);
    for (int idx = lane; idx < nhere * kBlocksPerRecord; idx += 32) {
        const int r = idx / kBlocksPerRecord, b = idx - r * kBlocksPerRecord;
        const int length = (int)s_rec[r * 61 + 2];
        const uint32_t *w = s_rec + r * 61 + 6 + 2 * b;
        const int nw = b == kBlocksPerRecord - 1 ? 1 : 2;
        bool diff = false;
        for (int k = 0; k < nw; k++) {
            const int s = 4 * b + 2 * k;
            const uint32_t expect = (s < length ? fill_h : 0u) | ((s + 1 < length ? fill_h : 0u) << 16);
            diff |= w[k] != expect;
        }
        if (diff) atomicOr(&s_mask[r], 1u << b);
    }
    __syncwarp();
    {
        const uint32_t cnt = lane < nhere ? __popc(s_mask[lane]) : 0u;
        uint32_t inc = cnt;
#pragma unroll
        for (int o = 1; o < kPackWarpRecs; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += a;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, inc, kPackWarpRecs - 1);
        unsigned long long base = 0;
        if (lane == 0 && total) base = atomicAdd((unsigned long long *)&scalars[S_NBLOCKS], (unsigned long long)total);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < kPackWarpRecs) s_off[lane] = (uint32_t)base + inc - cnt;
    }
    __syncwarp();
    for (int idx = lane; idx < nhere * 6; idx += 32) {
        const int r = idx / 6, k = idx - r * 6;
        const uint32_t *o = s_rec + r * 61;
        uint32_t h = o[k & 1];                                           // time (selects, not a branch chain)
        h = k == 2 ? o[4] : h;                                           // pulse_length
        h = k == 3 ? ((o[3] >> 16) | ((o[5] & 0xffffu) << 16)) : h;      // channel, record_i
        h = k == 4 ? s_off[r] : h;
        h = k == 5 ? s_mask[r] : h;
        chdr[(j0 + r) * 6 + k] = h;
    }
    for (int idx = lane; idx < nhere * kBlocksPerRecord; idx += 32) {
        const int r = idx / kBlocksPerRecord, b = idx - r * kBlocksPerRecord;
        const uint32_t m = s_mask[r];
        if (!((m >> b) & 1u)) continue;
        const uint32_t *w = s_rec + r * 61 + 6 + 2 * b;
        cblk[s_off[r] + __popc(m & ((1u << b) - 1u))] = make_uint2(w[0], b == kBlocksPerRecord - 1 ? 0u : w[1]);
    }
}

template <bool kCompact>
__global__ void __launch_bounds__(kPackWarps * 32)
k_pack(int64_t n_rec, DeviceConfig c, const uint32_t *__restrict__ rec_vals,
       const RecDesc *__restrict__ desc, const int16_t *__restrict__ dense, uint32_t *__restrict__ out,
       uint32_t *__restrict__ chdr, uint2 *__restrict__ cblk, int64_t *scalars) {
    __shared__ __align__(16) uint32_t s_rec_all[kPackWarps][kPackWarpRecs * 61];
    __shared__ RecDesc s_desc_all[kPackWarps][kPackWarpRecs];
    __shared__ uint32_t s_mask_all[kPackWarps][kPackWarpRecs], s_off_all[kPackWarps][kPackWarpRecs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j0 = ((int64_t)blockIdx.x * kPackWarps + warp) * kPackWarpRecs;
    const int nhere = (int)min((int64_t)kPackWarpRecs, n_rec - j0);
    if (nhere <= 0) return;
    uint32_t *s_rec = s_rec_all[warp];
    RecDesc *s_desc = s_desc_all[warp];
    uint32_t *s_mask = s_mask_all[warp], *s_off = s_off_all[warp];
    if (lane < nhere) s_desc[lane] = desc[rec_vals[j0 + lane]];
    if (kCompact && lane < kPackWarpRecs) s_mask[lane] = 0;
    __syncwarp();
    if (nhere == kPackWarpRecs) pack_assemble<true>(nhere, lane, c.p.dt, s_desc, dense, s_rec);
    else pack_assemble<false>(nhere, lane, c.p.dt, s_desc, dense, s_rec);
    __syncwarp();
    if (!kCompact) {
        // contiguous span of nhere * 244 bytes starting at a 16-byte aligned address (8 * 244 = 122 * 16)
        uint4 *dst = reinterpret_cast<uint4 *>(out + j0 * 61);
        const uint4 *srcv = reinterpret_cast<const uint4 *>(s_rec);
        const int nvec = (nhere * 61) / 4, rem = (nhere * 61) % 4;
        for (int i = lane; i < nvec; i += 32) dst[i] = srcv[i];
        if (lane < rem) out[j0 * 61 + nvec * 4 + lane] = s_rec[nvec * 4 + lane];
        return;
    }
    // ---- compact form ----
    compact_emit(nhere, lane, j0, c.p.baseline, s_rec, s_mask, s_off, chdr, cblk, scalars);
}

// Plain 244-byte records, already at their final position in HBM (the group-resident fused kernel writes
// them there) -> compact transport form.  Every warp streams 8 consecutive records (1952 bytes, 16-byte
// aligned) through its slice of shared memory.
__global__ void __launch_bounds__(kPackWarps * 32)
k_compact_records(int64_t n_rec, int baseline, const uint32_t *__restrict__ rec, uint32_t *__restrict__ chdr,
                  uint2 *__restrict__ cblk, int64_t *scalars) {
    __shared__ __align__(16) uint32_t s_rec_all[kPackWarps][kPackWarpRecs * 61];
    __shared__ uint32_t s_mask_all[kPackWarps][kPackWarpRecs], s_off_all[kPackWarps][kPackWarpRecs];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j0 = ((int64_t)blockIdx.x * kPackWarps + warp) * kPackWarpRecs;
    const int nhere = (int)min((int64_t)kPackWarpRecs, n_rec - j0);
    if (nhere <= 0) return;
    uint32_t *s_rec = s_rec_all[warp];
    if (lane < kPackWarpRecs) s_mask_all[warp][lane] = 0;
    const uint4 *src = reinterpret_cast<const uint4 *>(rec + j0 * 61);
    uint4 *dstv = reinterpret_cast<uint4 *>(s_rec);
    const int nvec = (nhere * 61) / 4, rem = (nhere * 61) % 4;
    for (int i = lane; i < nvec; i += 32) dstv[i] = src[i];
    if (lane < rem) s_rec[nvec * 4 + lane] = rec[j0 * 61 + nvec * 4 + lane];
    __syncwarp();
    compact_emit(nhere, lane, j0, baseline, s_rec, s_mask_all[warp], s_off_all[warp], chdr, cblk, scalars);
}

__global__ void k_group_info(int64_t n_groups, DeviceConfig c, const int64_t *group_lr,
                             const uint32_t *group_nitv, wfs_group_info *out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    wfs_group_info gi;
    int64_t lo = group_lr[2 * g], hi = group_lr[2 * g + 1];
    if (lo == LLONG_MAX) {
        gi.left = 0; gi.right = 0; gi.n_intervals = -1;   // empty group (no pulses)
    } else {
        gi.left = lo - c.p.trigger_window;
        gi.right = hi + c.p.trigger_window;
        if (gi.left % 2 != 0) gi.left -= 1;               // rawdata.py:221-222
        gi.n_intervals = group_nitv[g];
    }
    out[g] = gi;
}

// ---------------------------------------------------------------------------------------------
static int bits_for(uint64_t n_values) {   // bits needed to hold values 0..n_values-1
    int b = 0;
    while ((uint64_t(1) << b) < n_values) b++;
    return b;
}

Backend::Backend(const DeviceConfig *cfg, cudaStream_t stream, LaunchCounter *lc)
    : cfg_(cfg), stream_(stream), lc_(lc) {
    prim_.stream = stream;
    prim_.lc = lc;
    WFS_CUDA_CHECK(cudaEventCreate(&ev0_));
    WFS_CUDA_CHECK(cudaEventCreate(&ev1_));
    for (int i = 0; i < 8; i++) WFS_CUDA_CHECK(cudaEventCreate(&evp_[i]));
    WFS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    for (int i = 0; i < 5; i++) {
        WFS_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_[i], cudaStreamNonBlocking));
        WFS_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join_[i], cudaEventDisableTiming));
    }
    WFS_CUDA_CHECK(cudaHostAlloc((void **)&h_scalars_, sizeof(int64_t) * S_COUNT, cudaHostAllocDefault));
}

Backend::~Backend() {
    release();
    cudaEventDestroy(ev0_);
    cudaEventDestroy(ev1_);
    for (int i = 0; i < 8; i++) cudaEventDestroy(evp_[i]);
    cudaEventDestroy(ev_fork_);
    for (int i = 0; i < 5; i++) { cudaEventDestroy(ev_join_[i]); cudaStreamDestroy(aux_[i]); }
    if (h_scalars_) cudaFreeHost(h_scalars_);
}

void Backend::release() {
    DevBuf *all[] = {&keys_, &vals_, &st_, &sg_, &flags64_, &pulse_first_, &pulse_left_, &pulse_win_,
                     &win_first_pulse_, &win_meta_, &win_scan_, &group_tmin_, &group_lr_, &scalars_,
                     &dense_, &itv_, &itv_nrec_, &itv_rec0_, &rec_keys_, &rec_vals_,
                     &rec_itv_, &group_nitv_, &group_ix_, &pstart_, &flag8_, &cta_first_, &phq_,
                     &group_nvalid_, &group_out_, &group_win_, &rec_seg_, &fused_lists_, &fused_scal_, &fused_records_,
                     &fused_tkey_, &fused_gain_, &fused_desc_};
    for (DevBuf *b : all) b->release();
    prim_.release();
}

#define LAUNCH(kernel, grid, block, ...)                         \
    do {                                                         \
        kernel<<<(grid), (block), 0, stream_>>>(__VA_ARGS__);    \
        lc_->n++;                                                \
    } while (0)

void Backend::run(const PhotonBatch &b, uint8_t *records_out, int64_t cap_records,
                  wfs_group_info *group_info_out, BackendResult &res, const CompactOut *compact,
                  double plain_fraction) {
    const DeviceConfig &c = *cfg_;
    res = BackendResult();
    const int64_t n = b.n, ng = b.n_groups;
    if (ng <= 0) return;
    if (n >= (int64_t(1) << 31)) throw std::runtime_error("photon batch too large (>= 2^31)");
    if (n > 0 && fused_eligible(b)) {
        // small groups: one CTA per group from the photons to the records (fused.cu); for a host destination
        // the records then stream once more through k_compact_records into the compact transport form
        uint8_t *dst = records_out;
        const bool split = compact && plain_fraction > 0.0 && records_out;
        if (compact && !split) {
            fused_records_.reserve((size_t)WFS_RECORD_BYTES * (size_t)std::max<int64_t>(cap_records, 1) + 64);
            dst = fused_records_.as<uint8_t>();
        }
        if (run_fused(b, dst, cap_records, group_info_out, res)) {
            if (compact && !res.error && res.n_records > 0 && res.n_records <= cap_records) {
                // split transport: the leading rows stay as they are (a multiple of four rows keeps the rest 16-byte aligned)
                const int64_t n_plain = split ? (int64_t)(std::min(plain_fraction, 1.0) * (double)res.n_records) & ~int64_t(3) : 0;
                const int64_t n_comp = res.n_records - n_plain;
                res.n_plain = n_plain;
                if ((uint64_t)n_comp * kBlocksPerRecord >= (uint64_t(1) << 32)) { res.error = WFS_E_KEYBITS; return; }
                if (n_comp > 0) {
                    scalars_.reserve(sizeof(int64_t) * S_COUNT);
                    WFS_CUDA_CHECK(cudaMemsetAsync(scalars_.p, 0, sizeof(int64_t) * S_COUNT, stream_));
                    WFS_CUDA_CHECK(cudaEventRecord(evp_[5], stream_));
                    LAUNCH(k_compact_records, div_up(n_comp, kPackRecs), kPackWarps * 32, n_comp, c.p.baseline,
                           reinterpret_cast<const uint32_t *>(dst + (size_t)n_plain * WFS_RECORD_BYTES),
                           reinterpret_cast<uint32_t *>(compact->hdr), compact->blocks, scalars_.as<int64_t>());
                    WFS_CUDA_CHECK(cudaEventRecord(evp_[6], stream_));
                    WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, scalars_.p, sizeof(int64_t) * S_COUNT, cudaMemcpyDeviceToHost, stream_));
                    WFS_CUDA_CHECK(stream_sync(stream_));
                    res.n_blocks = h_scalars_[S_NBLOCKS];
                    float ms_compact = 0.f;
                    cudaEventElapsedTime(&ms_compact, evp_[5], evp_[6]);
                    res.ms_phase[6] += ms_compact;
                }
            }
            return;
        }
        res = BackendResult();
        if (b.trig_dpe_out)      // the fused attempt has added to the trigger counters: start over
            WFS_CUDA_CHECK(cudaMemsetAsync(b.trig_dpe_out, 0, sizeof(int32_t) * 2 * (size_t)b.n_pulse_calls, stream_));
        if (b.pmt_areas)
            WFS_CUDA_CHECK(cudaMemsetAsync(b.pmt_areas, 0, sizeof(int64_t) * (size_t)b.n_pulse_calls * c.p.n_tpc_pmts, stream_));
    }
    KeyLayout kl;
    kl.bits_rank = bits_for((uint64_t)b.max_rank + 1);
    kl.bits_group = bits_for((uint64_t)ng + 1);
    kl.shift_rank = kRelTimeBits;
    kl.shift_ch = kl.shift_rank + kl.bits_rank;
    kl.shift_group = kl.shift_ch + kChannelBits;
    kl.total_bits = kl.shift_group + kl.bits_group;
    if (kl.total_bits > 64 || kl.bits_group + kChannelBits > 32) {
        res.error = WFS_E_KEYBITS;
        return;
    }
    const int T = 256;
    scalars_.reserve(sizeof(int64_t) * S_COUNT);
    group_tmin_.reserve(sizeof(int64_t) * ng);
    group_lr_.reserve(sizeof(int64_t) * 2 * ng);
    group_nitv_.reserve(sizeof(uint32_t) * ng);
    int64_t *scal = scalars_.as<int64_t>();
    LAUNCH(k_init, div_up(std::max<int64_t>(ng, S_COUNT), T), T, group_tmin_.as<int64_t>(),
           group_lr_.as<int64_t>(), group_nitv_.as<uint32_t>(), ng, scal);
    int64_t np = 0, nw = 0;
    WFS_CUDA_CHECK(cudaEventRecord(evp_[0], stream_));
    if (n > 0) {
        keys_.reserve(sizeof(uint64_t) * n);
        vals_.reserve(sizeof(uint32_t) * n);
        st_.reserve(sizeof(int64_t) * n);
        sg_.reserve(sizeof(double) * n);
        flags64_.reserve(sizeof(uint64_t) * (n + 1));
        pstart_.reserve((size_t)n + 1);
        const bool seg_sort_on = !(getenv("WFS_SEGMENT_SORT") && atoi(getenv("WFS_SEGMENT_SORT")) == 0);
        const bool by_group = seg_sort_on && b.group_start != nullptr && b.max_group_photons <= kSegSortMax &&
                              kl.shift_group + 1 + 13 <= 64;
        if (by_group) {
            group_nvalid_.reserve(sizeof(uint32_t) * (ng + 1));
            group_out_.reserve(sizeof(uint32_t) * (ng + 1));
            WFS_CUDA_CHECK(cudaMemsetAsync(group_nvalid_.p, 0, sizeof(uint32_t) * (ng + 1), stream_));
        }
        LAUNCH(k_group_tmin, div_up(n, T), T, b, c, group_tmin_.as<int64_t>(),
               by_group ? group_nvalid_.as<uint32_t>() : nullptr);
        LAUNCH(k_build_keys, div_up(n, T), T, b, c, kl, group_tmin_.as<int64_t>(),
               keys_.as<uint64_t>(), vals_.as<uint32_t>(), scal);
        res.segment_sorted_photons = by_group ? 1 : 0;
        if (by_group) {
            // photons arrive group by group: order each group by (channel, pulse call, time) in
            // shared memory, dropping the invalid ones (dead PMT, no pattern)
            prim_.exclusive_scan_u32(group_nvalid_.as<uint32_t>(), group_out_.as<uint32_t>(), ng, true);
            prim_.sort_keys_alt.reserve((size_t)(n + 1) * sizeof(uint64_t));
            prim_.sort_vals_alt.reserve((size_t)(n + 1) * sizeof(uint32_t));
            const uint64_t invalid_key = (uint64_t)ng << kl.shift_group;
            prim_.segment_sort_pairs(keys_.as<uint64_t>(), vals_.as<uint32_t>(), prim_.sort_keys_alt.as<uint64_t>(),
                                     prim_.sort_vals_alt.as<uint32_t>(), b.group_start, group_out_.as<uint32_t>(),
                                     ng, b.max_group_photons, kl.shift_group, invalid_key, b.group_ranges);
            LAUNCH(k_fill_invalid_tail, div_up(n, T), T, n, group_out_.as<uint32_t>() + ng, invalid_key,
                   prim_.sort_keys_alt.as<uint64_t>());
            std::swap(keys_, prim_.sort_keys_alt);
            std::swap(vals_, prim_.sort_vals_alt);
        } else {
            prim_.sort_pairs(keys_.as<uint64_t>(), vals_.as<uint32_t>(), n, kl.total_bits);
        }
        WFS_CUDA_CHECK(cudaEventRecord(evp_[1], stream_));
        LAUNCH(k_gather_flags, div_up(n, T), T, b, kl, keys_.as<uint64_t>(), vals_.as<uint32_t>(),
               st_.as<int64_t>(), sg_.as<double>(), flags64_.as<uint64_t>(), pstart_.as<uint8_t>(), scal);
        // positions: reuse the (now consumed) alt key buffer of the sort for the scan output
        DevBuf &posbuf = prim_.sort_keys_alt;
        posbuf.reserve(sizeof(uint64_t) * (n + 1));
        prim_.exclusive_scan_u64(flags64_.as<uint64_t>(), posbuf.as<uint64_t>(), n, true);
        // upper bounds: every photon its own pulse/window
        pulse_first_.reserve(sizeof(uint32_t) * (n + 1));
        pulse_win_.reserve(sizeof(uint32_t) * (n + 1));
        win_first_pulse_.reserve(sizeof(uint32_t) * (n + 1));
        DevBuf &winkey = prim_.sort_vals_alt;   // [n] u32, free after the sort
        LAUNCH(k_emit_pulses, div_up(n + 1, T), T, n, kl, keys_.as<uint64_t>(),
               flags64_.as<uint64_t>(), posbuf.as<uint64_t>(), pulse_first_.as<uint32_t>(),
               pulse_win_.as<uint32_t>(), win_first_pulse_.as<uint32_t>(), winkey.as<uint32_t>(), scal);
        WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, scal, sizeof(int64_t) * S_COUNT,
                                       cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(stream_sync(stream_));
        if (h_scalars_[S_ERR]) { res.error = (int)h_scalars_[S_ERR]; return; }
        res.n_valid_photons = h_scalars_[S_NVALID];
        np = h_scalars_[S_NPULSES];
        nw = h_scalars_[S_NWIN];
    }
    res.n_pulses = np;
    const int64_t nwt = 2 * nw;
    res.n_windows = 0;
    int64_t n_tiles = 0, n_slots = 0, min_sample = 0, max_sample = 0, rec_seg_max = 0;
    bool rec_groups_disjoint = false;
    const bool he_rows = c.p.detector_nt != 0;
    static_assert(sizeof(WinMeta) == 64, "WinMeta layout");
    DevBuf &group_ix_buf = group_ix_;
    group_ix_buf.reserve(sizeof(int64_t) * ng);
    if (nw > 0) {
        pulse_left_.reserve(sizeof(int64_t) * 2 * np);
        win_meta_.reserve(sizeof(WinMeta) * nwt);
        win_scan_.reserve(sizeof(uint64_t) * (nwt + 1));
        LAUNCH(k_window_extents, div_up(nw, T), T, nw, c, st_.as<int64_t>(),
               pulse_first_.as<uint32_t>(), win_first_pulse_.as<uint32_t>(),
               prim_.sort_vals_alt.as<uint32_t>(), pulse_left_.as<int64_t>(), win_meta_.as<WinMeta>(),
               win_scan_.as<uint64_t>(), group_lr_.as<int64_t>(), scal, he_rows ? 1 : 0);
        prim_.exclusive_scan_u64(win_scan_.as<uint64_t>(), win_scan_.as<uint64_t>(), nwt, true);
    }
    if (np > 0 && b.trig_dpe_out && b.flags)
        LAUNCH(k_truth_pulses, div_up(np, T), T, np, b, c, vals_.as<uint32_t>(), st_.as<int64_t>(),
               sg_.as<double>(), pulse_first_.as<uint32_t>(), pulse_win_.as<uint32_t>(),
               prim_.sort_vals_alt.as<uint32_t>());
    LAUNCH(k_group_noise, div_up(ng, T), T, ng, c, group_lr_.as<int64_t>(), b.ix_rand, b.seed,
           group_ix_buf.as<int64_t>(), scal);
    if (nw > 0) {
        // total tiles / interval slots live in win_scan[nwt]
        uint64_t tot;
        WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, scal, sizeof(int64_t) * S_COUNT,
                                       cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(cudaMemcpyAsync(&tot, win_scan_.as<uint64_t>() + nwt, sizeof(uint64_t),
                                       cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(stream_sync(stream_));
        if (h_scalars_[S_ERR]) { res.error = (int)h_scalars_[S_ERR]; return; }
        n_tiles = (int64_t)(uint32_t)tot;
        n_slots = (int64_t)(tot >> 32);
        min_sample = h_scalars_[S_MINSAMPLE];
        max_sample = h_scalars_[S_MAXSAMPLE];
        res.n_tiles = n_tiles;
        dense_.reserve(sizeof(int16_t) * n_tiles * kBlk + 64);
        flag8_.reserve((size_t)n_tiles + 64);
        itv_.reserve(sizeof(Interval) * n_slots);
        itv_nrec_.reserve(sizeof(uint32_t) * (n_slots + 1));
        itv_rec0_.reserve(sizeof(uint32_t) * (n_slots + 1));
        WFS_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
        WFS_CUDA_CHECK(cudaEventRecord(evp_[2], stream_));
        // tile size: at most kTileWinMax windows may overlap a tile (every window is at least
        // left margin + right margin + 1 + 2 trigger windows long)
        const int min_blk = std::max(1, (c.p.pulse_left_margin + c.p.pulse_right_margin + 1 + 2 * c.p.trigger_window) / kBlk);
        const int tile_blk = std::min(kTileBlkMax, (kTileWinMax - 4) * min_blk);
        const int n_cta = div_up(n_tiles, tile_blk);
        cta_first_.reserve(sizeof(uint32_t) * (size_t)(n_cta + 1));
        LAUNCH(k_tile_index, div_up(nwt, T), T, nwt, tile_blk, win_scan_.as<uint64_t>(), win_meta_.as<WinMeta>(),
               cta_first_.as<uint32_t>());
        phq_.reserve(sizeof(uint32_t) * (size_t)(n + 1));
        {
            const int64_t nv = res.n_valid_photons;
            if (nv > 0)
                LAUNCH(k_photon_prep, div_up(nv, T), T, nv, c, st_.as<int64_t>(), sg_.as<double>(),
                       pstart_.as<uint8_t>(), prim_.sort_keys_alt.as<uint64_t>(), pulse_win_.as<uint32_t>(),
                       win_meta_.as<WinMeta>(), phq_.as<uint32_t>(), reinterpret_cast<double *>(flags64_.p));
        }
        static const int ctas_per_sm = getenv("WFS_DIGI_CTAS") ? atoi(getenv("WFS_DIGI_CTAS")) : kDigiCtasPerSm;
        LAUNCH(k_digitize, std::min(div_up(n_cta, kDigiWarps), kNumSMs * ctas_per_sm), kDigiThreads, n_tiles, tile_blk, n_cta, nwt, c,
               win_meta_.as<WinMeta>(), win_scan_.as<uint64_t>(), cta_first_.as<uint32_t>(),
               phq_.as<uint32_t>(), reinterpret_cast<const double *>(flags64_.p),
               pulse_first_.as<uint32_t>(), group_ix_buf.as<int64_t>(), dense_.as<int16_t>(),
               flag8_.as<uint8_t>(), scal);
        WFS_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
        WFS_CUDA_CHECK(cudaEventRecord(evp_[3], stream_));
        LAUNCH(k_zle, div_up(nwt, 128), 128, nwt, c, win_meta_.as<WinMeta>(), win_scan_.as<uint64_t>(),
               flag8_.as<uint8_t>(), itv_.as<Interval>(), itv_nrec_.as<uint32_t>(),
               group_nitv_.as<uint32_t>(), scal);
        prim_.exclusive_scan_u32(itv_nrec_.as<uint32_t>(), itv_rec0_.as<uint32_t>(), n_slots, true);
        WFS_CUDA_CHECK(cudaEventRecord(evp_[4], stream_));
        // (data type, group) segments of the record list, their largest size, and whether the groups'
        // windows are disjoint in time
        group_win_.reserve(sizeof(uint32_t) * ng);
        rec_seg_.reserve(sizeof(uint32_t) * (2 * ng + 1));
        WFS_CUDA_CHECK(cudaMemsetAsync(group_win_.p, 0xff, sizeof(uint32_t) * ng, stream_));
        LAUNCH(k_group_first_window, div_up(nw, T), T, nw, prim_.sort_vals_alt.as<uint32_t>(), group_win_.as<uint32_t>());
        LAUNCH(k_rec_segments, div_up(2 * ng + 1, T), T, ng, nw, c, group_win_.as<uint32_t>(),
               win_scan_.as<uint64_t>(), itv_rec0_.as<uint32_t>(), group_lr_.as<int64_t>(),
               rec_seg_.as<uint32_t>(), scal);
        uint32_t nrec32;
        WFS_CUDA_CHECK(cudaMemcpyAsync(&nrec32, itv_rec0_.as<uint32_t>() + n_slots, sizeof(uint32_t),
                                       cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, scal, sizeof(int64_t) * S_COUNT, cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(stream_sync(stream_));
        res.n_records = nrec32;
        rec_seg_max = h_scalars_[S_RECSEG_MAX];
        rec_groups_disjoint = h_scalars_[S_GROUPS_OVERLAP] == 0;
        WFS_CUDA_CHECK(cudaEventElapsedTime(&res.ms_digitize, ev0_, ev1_));
    }
    if (group_info_out)
        LAUNCH(k_group_info, div_up(ng, T), T, ng, c, group_lr_.as<int64_t>(),
               group_nitv_.as<uint32_t>(), group_info_out);
    const int64_t nrec = res.n_records;
    if (nrec > 0 && nrec <= cap_records) {
        int time_bits = bits_for((uint64_t)(max_sample - min_sample + 1));
        int key_bits = kChannelBits + time_bits + 2;
        if (key_bits > 64) { res.error = WFS_E_KEYBITS; return; }
        rec_keys_.reserve(sizeof(uint64_t) * nrec);
        rec_vals_.reserve(sizeof(uint32_t) * nrec);
        rec_itv_.reserve(sizeof(RecDesc) * nrec);
        LAUNCH(k_rec_keys, div_up(n_slots, T), T, n_slots, c, itv_.as<Interval>(),
               itv_nrec_.as<uint32_t>(), itv_rec0_.as<uint32_t>(), min_sample, time_bits,
               rec_keys_.as<uint64_t>(), rec_vals_.as<uint32_t>(), rec_itv_.as<RecDesc>());
        const bool seg_on = !(getenv("WFS_SEGMENT_SORT") && atoi(getenv("WFS_SEGMENT_SORT")) == 0);
        if (seg_on && rec_groups_disjoint && rec_seg_max <= kSegSortMax && kChannelBits + time_bits + 1 + 13 <= 64) {
            // every (data type, group) segment ordered by (time, channel) in shared memory
            prim_.sort_keys_alt.reserve((size_t)nrec * sizeof(uint64_t));
            prim_.sort_vals_alt.reserve((size_t)nrec * sizeof(uint32_t));
            prim_.segment_sort_pairs(rec_keys_.as<uint64_t>(), rec_vals_.as<uint32_t>(),
                                     prim_.sort_keys_alt.as<uint64_t>(), prim_.sort_vals_alt.as<uint32_t>(),
                                     rec_seg_.as<uint32_t>(), nullptr, 2 * ng, rec_seg_max, kChannelBits + time_bits,
                                     ~0ull);
            std::swap(rec_keys_, prim_.sort_keys_alt);
            std::swap(rec_vals_, prim_.sort_vals_alt);
            res.segment_sorted_records = 1;
        } else {
            prim_.sort_pairs(rec_keys_.as<uint64_t>(), rec_vals_.as<uint32_t>(), nrec, key_bits);
        }
        LAUNCH(k_class_counts, 1, 32, nrec, rec_keys_.as<uint64_t>(), kChannelBits + time_bits, scal);
        WFS_CUDA_CHECK(cudaEventRecord(evp_[5], stream_));
        if (compact) {
            if ((uint64_t)nrec * kBlocksPerRecord >= (uint64_t(1) << 32)) { res.error = WFS_E_KEYBITS; return; }
            LAUNCH(k_pack<true>, div_up(nrec, kPackRecs), kPackWarps * 32, nrec, c, rec_vals_.as<uint32_t>(),
                   rec_itv_.as<RecDesc>(), dense_.as<int16_t>(), nullptr,
                   reinterpret_cast<uint32_t *>(compact->hdr), compact->blocks, scal);
        } else {
            LAUNCH(k_pack<false>, div_up(nrec, kPackRecs), kPackWarps * 32, nrec, c, rec_vals_.as<uint32_t>(),
                   rec_itv_.as<RecDesc>(), dense_.as<int16_t>(), reinterpret_cast<uint32_t *>(records_out),
                   nullptr, nullptr, scal);
        }
    }
    WFS_CUDA_CHECK(cudaEventRecord(evp_[6], stream_));
    WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, scal, sizeof(int64_t) * S_COUNT,
                                   cudaMemcpyDeviceToHost, stream_));
    WFS_CUDA_CHECK(stream_sync(stream_));
    WFS_CUDA_CHECK(cudaGetLastError());
    {
        const bool full = nw > 0 && nrec > 0 && nrec <= cap_records;
        auto el = [&](int a, int b) { float ms = 0; cudaEventElapsedTime(&ms, evp_[a], evp_[b]); return ms; };
        if (n > 0 && nw > 0) { res.ms_phase[1] = el(0, 1); res.ms_phase[2] = el(1, 2); res.ms_phase[3] = el(2, 3); res.ms_phase[4] = el(3, 4); }
        if (full) { res.ms_phase[5] = el(4, 5); res.ms_phase[6] = el(5, 6); }
    }
    if (h_scalars_[S_ERR]) res.error = (int)h_scalars_[S_ERR];
    res.n_intervals = h_scalars_[S_NITV];
    res.n_samples = h_scalars_[S_NSAMPLES];
    res.n_blocks = h_scalars_[S_NBLOCKS];
    res.n_dense_tiles = h_scalars_[S_DENSE_TILES];
    res.n_windows = nwt;
    if (nrec > 0 && nrec <= cap_records) {
        const int64_t i1 = h_scalars_[S_CLASS1], i2 = h_scalars_[S_CLASS2];
        res.n_rec_class[0] = i1;
        res.n_rec_class[1] = i2 - i1;
        res.n_rec_class[2] = nrec - i2;
    }
}

}  // namespace wfs
