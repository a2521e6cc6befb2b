// Group-resident back end: ONE CTA owns a digitisation group from its photons to its raw_records.
//
//   photons of the group (generation order, HBM, read once)
//     -> key (channel | pulse call | sample | ns remainder | index | dpe) in shared memory, bucketed by
//        channel (counting sort) and ordered inside the channel by (pulse call, time)
//                                                                   Pulse.__call__   pulse.py:82-144
//     -> phase A, one warp per (group, channel) window: pulse / window extents (pulse.py:118-127,
//        rawdata.py:231-235,258-259), truth counters of every pulse (pulse.py:229-271), template
//        superposition in a REGISTER ring -- lane j holds the fp64 current of sample (base + j); photons are
//        taken in ascending time, equal-ns photons merged first, mul and add unfused: the summation order
//        of Pulse.add_current (pulse.py:276-318) -- one rounding per pulse and sample, -around(current *
//        current_2_adc) (rawdata.py:236-239); baseline, clamp, threshold (rawdata.py:290-296,441-458);
//        the hysteresis interval search (utils.py:13-58, rawdata.py:296-308) straight on the ballot words
//     -> record keys (time, channel) of the group ordered in shared memory (strax.sort_by_time)
//     -> the group's first record index: decoupled look-back over the groups in front (groups are disjoint
//        in time, so group order IS record order)
//     -> phase B, one warp per ZLE interval: the samples again (same arithmetic), 244-byte records
//        assembled by the warp and written ONCE, at their final sorted position
//                                                                   strax_interface.py:425-436
// Nothing but the photons is read from HBM and nothing but the records is written: no sort keys, no dense
// ADC buffer, no flags, no record descriptors.  Used when every group of a batch fits (<= kFusedMaxPhotons
// photons, no noise, no high-energy twin rows); anything else takes the multi-pass back end (backend.cu),
// which stays the reference implementation of the same arithmetic for heavy S2s.
#include "backend.cuh"
#include "fused.cuh"

#include <algorithm>
#include <limits.h>
#include <stdlib.h>

namespace wfs {

namespace {

constexpr int kNegPos = -(1 << 29);
constexpr uint32_t kPadKey = 0xffffffffu;

// look-back status word: [63:62] state, [61:0] record count (aggregate of the group or inclusive prefix)
constexpr uint64_t kStEmpty = 0, kStAgg = 1ull << 62, kStPrefix = 2ull << 62, kStMask = 3ull << 62;

struct FusedShared {          // fixed-size part of the shared memory, the arrays follow
    int64_t origin_q;         // absolute sample index of key sample 0
    int32_t n_valid, n_win, n_itv, n_rec, n_pulses, n_emitted;
    int32_t next_win, next_itv;
    int32_t lo, hi;           // min pulse left / max pulse right of the group, relative to origin_q
    int32_t group;
    uint32_t rec_base;
    int32_t overflow;
    unsigned long long n_samples;
};

struct Layout {               // byte offsets into the dynamic shared memory
    int keys, chan_start, chan_fill, win_ch, tmpl, scratch, tiles, rkey, itv, total;
};

__host__ __device__ inline Layout make_layout(int n_cap, int n_ch, int n_warps, int tmpl_len) {
    Layout L;
    int o = (int)((sizeof(FusedShared) + 15) & ~15u);
    L.keys = o; o += 8 * n_cap;
    L.chan_start = o; o += 4 * (n_ch + 1);
    L.chan_fill = o; o += 4 * (n_ch + 1);
    L.win_ch = o; o += 2 * (n_ch + 2);
    o = (o + 15) & ~15;
    L.tmpl = o; o += 8 * tmpl_len;
    // scratch: unsorted keys while loading; afterwards per-warp sample tiles, record keys, intervals
    L.scratch = o;
    L.tiles = o;
    L.rkey = L.tiles + n_warps * kFusedTile * 4;
    L.itv = L.rkey + kFusedRecCap * 4;
    const int after = L.itv + kFusedItvCap * 8;
    o = after > L.scratch + 8 * n_cap ? after : L.scratch + 8 * n_cap;
    L.total = (o + 15) & ~15;
    return L;
}

__device__ __forceinline__ int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

struct KeyFmt {
    int shift_pc, shift_ch;   // key = ch << shift_ch | relpc << shift_pc | sample << 18 | rem << 14 | idx << 1 | dpe
};
constexpr int kShiftSample = 18, kShiftRem = 14, kSampleBits = 20;

__device__ __forceinline__ int key_sample(uint64_t k) { return (int)((k >> kShiftSample) & ((1u << kSampleBits) - 1u)); }
__device__ __forceinline__ int key_rem(uint64_t k) { return (int)((k >> kShiftRem) & 15u); }
__device__ __forceinline__ uint32_t key_time(uint64_t k) { return (uint32_t)((k >> kShiftRem) & ((1u << (kSampleBits + 4)) - 1u)); }
__device__ __forceinline__ int key_idx(uint64_t k) { return (int)((k >> 1) & 8191u); }
__device__ __forceinline__ int key_dpe(uint64_t k) { return (int)(k & 1u); }

struct Window {               // one (group, channel) window, warp-uniform
    int ch, a, e;             // photons keys[a, e)
    int wl, len;              // first sample relative to origin_q, samples
    int n_pulses;
    int thr, mult;
};

// The photons of pulse [pa, pe) that can reach samples [lo, hi] (relative to origin_q) are superposed in the
// register ring; every finished sample is handed to `sink(sample, value, lane_active)` exactly once, in
// ascending sample order per call.  All lanes run the same control flow.
//   gains: lane j holds the gain of photon keys[a0 + j] (prefetched by the caller, one round trip to L2 per
//   window); photons further back in a long list are fetched one by one through `gain_of`.
template <typename GainOf, typename Sink>
__device__ __forceinline__ void superpose_pulse(const uint64_t *keys, int pa, int pe, int a0, double gpre, GainOf &&gain_of,
                                                const double *s_tmpl, int tlen, int lo, int hi, int lane, Sink &&sink) {
    auto gain_at = [&](int k) -> double {
        const double pre = __shfl_sync(0xffffffffu, gpre, (k - a0) & 31);
        return (k - a0) < 32 ? pre : gain_of(key_idx(keys[k]));
    };
    double acc = 0.0;
    int base = 0;
    bool have = false;
    int k = pa;
    while (k < pe) {
        const uint64_t key = keys[k];
        const uint32_t tk = key_time(key);
        double g = gain_at(k);
        int k2 = k + 1;
        while (k2 < pe && key_time(keys[k2]) == tk) {      // equal-ns photons: gains summed first (pulse.py:301-318)
            g = __dadd_rn(g, gain_at(k2));
            k2++;
        }
        k = k2;
        const int T = key_sample(key);
        if (T + tlen - 1 < lo) continue;
        if (T > hi) break;
        if (have) {
            const int d = T - base;
            if (d > 0) {
                sink(base + lane, acc, lane < d && lane < tlen);
                const double moved = __shfl_down_sync(0xffffffffu, acc, (unsigned)min(d, 31));
                acc = (d < 32 && lane + d < 32) ? moved : 0.0;
            }
        }
        base = T;
        have = true;
        if (lane < tlen) acc = __dadd_rn(acc, __dmul_rn(s_tmpl[key_rem(key) * tlen + lane], g));
    }
    if (have) sink(base + lane, acc, lane < tlen);
}

__device__ __forceinline__ int adc_of(double cur, double c2a, int mult) {
    return -__double2int_rn(__dmul_rn(cur, c2a)) * mult;      // one rounding per pulse and sample: rawdata.py:236-239
}

struct ZleState {
    int last = kNegPos, start = kNegPos;
};

}  // namespace

__global__ void __launch_bounds__(kFusedThreads)
k_group_fused(FusedArgs A) {
    extern __shared__ __align__(16) uint8_t smem[];
    const PhotonBatch &b = A.b;
    const DeviceConfig &c = A.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int n_ch = c.p.n_tpc_pmts, dt = c.p.dt, tlen = c.p.template_length;
    const Layout L = make_layout(A.n_cap, n_ch, n_warps, dt * tlen);
    FusedShared &S = *reinterpret_cast<FusedShared *>(smem);
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem + L.keys);
    int32_t *s_cstart = reinterpret_cast<int32_t *>(smem + L.chan_start);
    int32_t *s_cfill = reinterpret_cast<int32_t *>(smem + L.chan_fill);
    uint16_t *s_winch = reinterpret_cast<uint16_t *>(smem + L.win_ch);
    double *s_tmpl = reinterpret_cast<double *>(smem + L.tmpl);
    uint64_t *s_raw = reinterpret_cast<uint64_t *>(smem + L.scratch);
    int32_t *s_tile = reinterpret_cast<int32_t *>(smem + L.tiles) + warp * kFusedTile;
    uint32_t *s_rkey = reinterpret_cast<uint32_t *>(smem + L.rkey);
    uint64_t *s_itv = reinterpret_cast<uint64_t *>(smem + L.itv);

    const double c2a = c.p.current_2_adc;
    const int LM = c.p.pulse_left_margin, RM = c.p.pulse_right_margin, tw = c.p.trigger_window, H = 2 * tw + 1;
    const int baseline = c.p.baseline;
    const int key_bias = LM + tw + 2;                 // record key time = left relative to origin + bias >= 0
    const KeyFmt kf{kShiftSample + kSampleBits, kShiftSample + kSampleBits + A.relpc_bits};

    for (int i = tid; i < dt * tlen; i += blockDim.x) s_tmpl[i] = c.templates[i];

    for (;;) {
        __syncthreads();                              // everything of the previous group is done
        if (tid == 0) S.group = (int32_t)atomicAdd(A.ticket, 1u);
        __syncthreads();
        const int g = S.group;
        if (g >= (int)b.n_groups) break;
        // ------------------------------------------------------------------ load ----
        if (tid == 0) {
            S.n_valid = S.n_win = S.n_itv = S.n_rec = S.n_pulses = S.n_emitted = 0;
            S.next_win = S.next_itv = 0;
            S.lo = INT_MAX; S.hi = INT_MIN;
            S.overflow = 0;
            S.n_samples = 0;
            S.origin_q = floordiv64(A.group_t0[g], dt);
        }
        for (int i = tid; i <= n_ch; i += blockDim.x) s_cfill[i] = 0;
        __syncthreads();
        const int64_t origin_t = S.origin_q * dt;
        const int32_t run0 = A.group_run0[g];
        uint32_t r_lo[4], r_n[4];
        int n_g = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            r_lo[r] = r_n[r] = 0;
            if (r < b.group_ranges) {
                const uint32_t *gs = b.group_start + (size_t)r * (b.n_groups + 1);
                r_lo[r] = gs[g];
                r_n[r] = gs[g + 1] - gs[g];
            }
            n_g += (int)r_n[r];
        }
        // the four ranges are addressed as one list: list index i -> global photon index
        auto global_index = [&](int i) -> uint32_t {
            if (i < (int)r_n[0]) return r_lo[0] + i;
            i -= r_n[0];
            if (i < (int)r_n[1]) return r_lo[1] + i;
            i -= r_n[1];
            if (i < (int)r_n[2]) return r_lo[2] + i;
            return r_lo[3] + (i - r_n[2]);
        };
        for (int i = tid; i < n_g; i += blockDim.x) {
            const uint32_t gi = global_index(i);
            const int32_t ch = b.channel[gi];
            const int32_t run = b.instr_run[b.pulse_call[gi]];
            const uint8_t fl = b.flags[gi];
            uint64_t key = ~0ull;
            if (ch >= 0 && ch < n_ch && run >= 0 && c.gains[ch] != 0.0) {
                const int64_t rel = b.t[gi] - origin_t;
                const int64_t q = rel / dt;                       // rel >= 0: origin is a lower bound
                const int relpc = 2 * (run - run0) + ((fl >> 1) & 1);
                if (rel < 0 || q >= (1 << kSampleBits) || relpc < 0 || relpc >= (1 << A.relpc_bits)) {
                    A.scalars[FS_OVERFLOW] = 1;        // outside the key range: the multi-pass back end decides
                } else {
                    key = ((uint64_t)ch << kf.shift_ch) | ((uint64_t)relpc << kf.shift_pc) |
                          ((uint64_t)q << kShiftSample) | ((uint64_t)(rel - q * dt) << kShiftRem) |
                          ((uint64_t)i << 1) | (uint64_t)(fl & 1);
                    atomicAdd(&s_cfill[ch], 1);
                }
            }
            s_raw[i] = key;
        }
        __syncthreads();
        // channel offsets (exclusive scan of the counts) by warp 0; list of non-empty channels
        if (warp == 0) {
            int carry = 0, nwin = 0;
            for (int c0 = 0; c0 < n_ch; c0 += 32) {
                const int ch = c0 + lane;
                const int cnt = ch < n_ch ? s_cfill[ch] : 0;
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (ch < n_ch) s_cstart[ch] = carry + inc - cnt;
                const unsigned m = __ballot_sync(0xffffffffu, cnt > 0);
                if (cnt > 0) s_winch[nwin + __popc(m & ((1u << lane) - 1u))] = (uint16_t)ch;
                nwin += __popc(m);
                carry += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) { s_cstart[n_ch] = carry; S.n_valid = carry; S.n_win = nwin; }
        }
        __syncthreads();
        for (int i = tid; i <= n_ch; i += blockDim.x) s_cfill[i] = 0;
        __syncthreads();
        for (int i = tid; i < n_g; i += blockDim.x) {
            const uint64_t key = s_raw[i];
            if (key == ~0ull) continue;
            const int ch = (int)(key >> kf.shift_ch);
            s_keys[s_cstart[ch] + atomicAdd(&s_cfill[ch], 1)] = key;
        }
        __syncthreads();
        // inside a channel: ascending (pulse call, time, index) -- insertion sort by one thread for short
        // lists, rank sort by a warp for long ones
        for (int w = tid; w < S.n_win; w += blockDim.x) {
            const int ch = s_winch[w], a = s_cstart[ch], n = s_cstart[ch + 1] - a;
            if (n > 48) continue;
            for (int i = 1; i < n; i++) {
                const uint64_t x = s_keys[a + i];
                int j = i - 1;
                while (j >= 0 && s_keys[a + j] > x) { s_keys[a + j + 1] = s_keys[a + j]; j--; }
                s_keys[a + j + 1] = x;
            }
        }
        __syncthreads();
        for (int w = warp; w < S.n_win; w += n_warps) {
            const int ch = s_winch[w], a = s_cstart[ch], n = s_cstart[ch + 1] - a;
            if (n <= 48) continue;
            for (int i = lane; i < n; i += 32) {                  // keys are distinct (index bits)
                const uint64_t x = s_keys[a + i];
                int rank = 0;
                for (int j = 0; j < n; j++) rank += s_keys[a + j] < x ? 1 : 0;
                s_raw[a + rank] = x;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) s_keys[a + i] = s_raw[a + i];
        }
        __syncthreads();      // s_raw (the scratch area) is free from here on: tiles, record keys, intervals

        // gains stay in HBM / L2 (8 bytes per photon of shared memory buy a second CTA per SM instead):
        // read by list index, prefetched per window
        auto gain_of = [&](int idx) -> double { return b.gain[global_index(idx)]; };

        // a window of the group, warp-uniform
        auto window_of = [&](int ch) -> Window {
            Window W;
            W.ch = ch;
            W.a = s_cstart[ch];
            W.e = s_cstart[ch + 1];
            W.thr = c.zle_thr[ch];
            W.mult = 1;
            int qmin = INT_MAX, qmax = INT_MIN, np = 0;
            for (int k0 = W.a; k0 < W.e; k0 += 32) {
                const int k = k0 + lane;
                const bool in = k < W.e;
                const uint64_t key = in ? s_keys[k] : 0;
                const uint64_t prev = (in && k > W.a) ? s_keys[k - 1] : ~0ull;
                const bool first = in && (k == W.a || (key >> kf.shift_pc) != (prev >> kf.shift_pc));
                np += __popc(__ballot_sync(0xffffffffu, first));
                const int q = in ? key_sample(key) : INT_MAX;
                qmin = min(qmin, q);
                qmax = max(qmax, in ? q : INT_MIN);
            }
            qmin = __reduce_min_sync(0xffffffffu, qmin);
            qmax = __reduce_max_sync(0xffffffffu, qmax);
            W.n_pulses = np;
            W.wl = qmin - LM - tw;                                 // pulse.py:118-127, rawdata.py:258-259
            W.len = (qmax + RM + tw) - W.wl + 1;
            return W;
        };

        // ------------------------------------------------------------------ phase A ----
        for (;;) {
            int w = 0;
            if (lane == 0) w = atomicAdd(&S.next_win, 1);
            w = __shfl_sync(0xffffffffu, w, 0);
            if (w >= S.n_win) break;
            const Window W = window_of(s_winch[w]);
            if (W.len > kMaxGroupSamples + 1) {
                if (lane == 0) A.scalars[FS_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
                continue;
            }
            const double gpre = W.a + lane < W.e ? gain_of(key_idx(s_keys[W.a + lane])) : 0.0;
            if (lane == 0) {
                atomicMin(&S.lo, W.wl + tw);
                atomicMax(&S.hi, W.wl + W.len - 1 - tw);
                atomicAdd(&S.n_pulses, W.n_pulses);
                atomicAdd(&S.n_samples, (unsigned long long)W.len);
            }
            // ---- truth counters of every pulse (pulse.py:229-271; odd calls are PMT afterpulses: none) ----
            if (b.trig_dpe_out) {
                const double thr_t = (double)(baseline - 1 - W.thr) - 0.5;
                const double gch = c.gains[W.ch];
                const bool per_pmt = b.pmt_counts != nullptr;
                int pa = W.a;
                while (pa < W.e) {
                    const uint64_t pck = s_keys[pa] >> kf.shift_pc;
                    int pe = pa + 1;
                    while (pe < W.e && (s_keys[pe] >> kf.shift_pc) == pck) pe++;
                    const int relpc = (int)(pck & ((1u << A.relpc_bits) - 1u));
                    if (!(relpc & 1)) {
                        int ndpe = 0;
                        for (int k = pa + lane; k < pe; k += 32) ndpe += key_dpe(s_keys[k]);
                        ndpe = __reduce_add_sync(0xffffffffu, ndpe);
                        int trig = 0, n_trig = 0;
                        long long area = 0, area_trig = 0;
                        const int stop = per_pmt ? pe : pa + ndpe;
                        for (int k = pa + lane; k < stop; k += 32) {
                            const uint64_t key = s_keys[k];
                            const double gn = k - W.a == lane ? gpre : gain_of(key_idx(key));
                            const bool above = gn * c.current_max[key_rem(key)] * c2a > thr_t;
                            if (above && k < pa + ndpe) trig++;
                            if (per_pmt) {
                                const long long ar = llrint(gn / gch * 4294967296.0);
                                area += ar;
                                if (above) { n_trig++; area_trig += ar; }
                            }
                        }
                        trig = __reduce_add_sync(0xffffffffu, trig);
                        const int32_t pc = 2 * run0 + relpc;
                        if (per_pmt) {
                            n_trig = __reduce_add_sync(0xffffffffu, n_trig);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) {
                                area += __shfl_xor_sync(0xffffffffu, area, o);
                                area_trig += __shfl_xor_sync(0xffffffffu, area_trig, o);
                            }
                        }
                        if (lane == 0) {
                            if (trig) {
                                atomicAdd(&b.trig_dpe_out[2 * pc], trig);
                                if (W.ch >= c.p.n_top_pmts) atomicAdd(&b.trig_dpe_out[2 * pc + 1], trig);
                            }
                            if (per_pmt) {
                                const int64_t npmt = n_ch;
                                int32_t *cnt = b.pmt_counts + ((int64_t)(pc >> 1) * 4) * npmt + W.ch;
                                int64_t *ar_out = b.pmt_areas + ((int64_t)(pc >> 1) * 2) * npmt + W.ch;
                                cnt[0] = pe - pa;
                                cnt[npmt] = pe - pa + ndpe;
                                cnt[2 * npmt] = n_trig;
                                cnt[3 * npmt] = n_trig + trig;
                                ar_out[0] = area;
                                ar_out[npmt] = area_trig;
                            }
                        }
                    }
                    pa = pe;
                }
            }
            // ---- samples below threshold -> ZLE intervals ----
            ZleState Z;
            int n_emitted = 0;
            auto emit = [&](int s, int e) {            // utils.py:44-52, rawdata.py:303-308; s, e window-local
                int l = s - tw, r = e + tw;
                l = max(0, min(l, W.len - 1));
                r = max(0, min(r, W.len - 1));
                l = (l + 1) & ~1;
                r = r & ~1;
                const int plen = max(r - l + 1, 0);
                const int nrec = (plen + WFS_SAMPLES_PER_RECORD - 1) / WFS_SAMPLES_PER_RECORD;
                n_emitted++;
                if (lane == 0 && nrec > 0) {
                    const int slot = atomicAdd(&S.n_itv, 1);
                    const int r0 = atomicAdd(&S.n_rec, nrec);
                    if (slot < kFusedItvCap && r0 + nrec <= kFusedRecCap) {
                        // interval: left relative to origin (biased, 22 bits) | length (21 bits) | channel (10)
                        s_itv[slot] = ((uint64_t)(uint32_t)(W.wl + l + key_bias) << 32) | ((uint64_t)(uint32_t)plen << 10) |
                                      (uint64_t)W.ch;
                        for (int i = 0; i < nrec; i++)
                            s_rkey[r0 + i] = ((uint32_t)(W.wl + l + key_bias + WFS_SAMPLES_PER_RECORD * i) << 10) | (uint32_t)W.ch;
                    } else {
                        S.overflow = 1;
                    }
                }
            };
            auto feed = [&](int pos0, uint32_t word) {   // flagged samples pos0 + bit, ascending over calls
                while (word) {
                    const int bit = __ffs(word) - 1;
                    const uint32_t run = word | (word - 1);
                    const uint32_t stopb = ~run & (run + 1);
                    const int run_end = stopb == 0 ? 31 : __ffs(stopb) - 2;
                    const int p0 = pos0 + bit, p1 = pos0 + run_end;
                    if (Z.last == kNegPos) Z.start = p0;
                    else if (p0 - Z.last > H) { emit(Z.start, Z.last); Z.start = p0; }
                    Z.last = p1;
                    word = run_end >= 31 ? 0u : (word & ~((2u << run_end) - 1u));
                }
            };
            if (W.n_pulses == 1) {
                // one pulse: a finished sample of the ring is final -- no sample buffer at all
                superpose_pulse(s_keys, W.a, W.e, W.a, gpre, gain_of, s_tmpl, tlen, INT_MIN / 2, INT_MAX / 2, lane,
                                [&](int s, double cur, bool on) {
                                    bool flag = false;
                                    if (on && cur != 0.0) flag = max(adc_of(cur, c2a, 1) + baseline, 0) < W.thr;
                                    const uint32_t word = __ballot_sync(0xffffffffu, flag);
                                    if (word) feed(s - lane - W.wl, word);
                                });
            } else {
                // several pulse calls on the channel: integer sum over the pulses in a sample tile
                for (int t0 = 0; t0 < W.len; t0 += kFusedTile) {
                    const int n = min(kFusedTile, W.len - t0);
                    for (int i = lane; i < n; i += 32) s_tile[i] = 0;
                    __syncwarp();
                    int pa = W.a;
                    while (pa < W.e) {
                        const uint64_t pck = s_keys[pa] >> kf.shift_pc;
                        int pe = pa + 1;
                        while (pe < W.e && (s_keys[pe] >> kf.shift_pc) == pck) pe++;
                        superpose_pulse(s_keys, pa, pe, W.a, gpre, gain_of, s_tmpl, tlen, W.wl + t0, W.wl + t0 + n - 1, lane,
                                        [&](int s, double cur, bool on) {
                                            const int i = s - W.wl - t0;
                                            if (on && cur != 0.0 && i >= 0 && i < n) s_tile[i] += adc_of(cur, c2a, 1);
                                        });
                        __syncwarp();
                        pa = pe;
                    }
                    for (int i0 = 0; i0 < n; i0 += 32) {
                        const int i = i0 + lane;
                        const bool flag = i < n && max(s_tile[i] + baseline, 0) < W.thr;
                        const uint32_t word = __ballot_sync(0xffffffffu, flag);
                        if (word) feed(t0 + i0, word);
                    }
                    __syncwarp();
                }
            }
            if (Z.last != kNegPos) emit(Z.start, Z.last);
            if (lane == 0 && n_emitted) atomicAdd(&S.n_emitted, n_emitted);
        }
        __syncthreads();
        // ------------------------------------------------------------------ record order ----
        const bool overflow = S.overflow != 0;
        const int n_rec = overflow ? 0 : S.n_rec, n_itv = overflow ? 0 : S.n_itv;
        if (overflow && tid == 0) A.scalars[FS_OVERFLOW] = 1;
        int n_sort = 32;
        while (n_sort < n_rec) n_sort <<= 1;
        for (int i = n_rec + tid; i < n_sort; i += blockDim.x) s_rkey[i] = kPadKey;
        __syncthreads();
        for (int k = 2; k <= n_sort; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < n_sort; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t x = s_rkey[i], y = s_rkey[ixj];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { s_rkey[i] = y; s_rkey[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        }
        // ------------------------------------------------------------------ first record of the group ----
        if (warp == 0) {
            volatile uint64_t *st = A.status;
            uint64_t base = 0;
            if (g == 0) {
                if (lane == 0) { __threadfence(); st[0] = kStPrefix | (uint64_t)n_rec; }
            } else {
                if (lane == 0) { __threadfence(); st[g] = kStAgg | (uint64_t)n_rec; }
                int pos = g - 1;
                for (;;) {
                    const int p = pos - lane;
                    uint64_t v = kStPrefix;                        // in front of group 0: prefix 0
                    if (p >= 0) {
                        do { v = st[p]; } while ((v & kStMask) == kStEmpty);
                    }
                    const unsigned is_prefix = __ballot_sync(0xffffffffu, (v & kStMask) == kStPrefix);
                    const int first = __ffs(is_prefix) - 1;        // nearest group with an inclusive prefix
                    uint64_t add = (first < 0 || lane <= first) ? (v & ~kStMask) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) add += __shfl_xor_sync(0xffffffffu, add, o);
                    base += add;
                    if (first >= 0) break;
                    pos -= 32;
                }
                if (lane == 0) { __threadfence(); st[g] = kStPrefix | (base + (uint64_t)n_rec); }
            }
            if (lane == 0) {
                S.rec_base = (uint32_t)base;
                if (g == (int)b.n_groups - 1) A.scalars[FS_NREC] = (int64_t)(base + (uint64_t)n_rec);
                // group bookkeeping for the chunker (rawdata.py:215-222)
                wfs_group_info gi;
                if (S.lo == INT_MAX) {
                    gi.left = 0; gi.right = 0; gi.n_intervals = -1;
                } else {
                    gi.left = S.origin_q + S.lo - tw;
                    gi.right = S.origin_q + S.hi + tw;
                    if (gi.right - gi.left >= kMaxGroupSamples) A.scalars[FS_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
                    if (gi.left % 2 != 0) gi.left -= 1;
                    gi.n_intervals = S.n_emitted;
                }
                if (A.group_info) A.group_info[g] = gi;
                atomicAdd((unsigned long long *)&A.scalars[FS_NVALID], (unsigned long long)S.n_valid);
                atomicAdd((unsigned long long *)&A.scalars[FS_NPULSES], (unsigned long long)S.n_pulses);
                atomicAdd((unsigned long long *)&A.scalars[FS_NWIN], (unsigned long long)S.n_win);
                atomicAdd((unsigned long long *)&A.scalars[FS_NITV], (unsigned long long)S.n_emitted);
                atomicAdd((unsigned long long *)&A.scalars[FS_NSAMPLES], S.n_samples);
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ phase B ----
        const int64_t rec_base = S.rec_base;
        for (;;) {
            int it = 0;
            if (lane == 0) it = atomicAdd(&S.next_itv, 1);
            it = __shfl_sync(0xffffffffu, it, 0);
            if (it >= n_itv) break;
            const uint64_t iv = s_itv[it];
            const int ch = (int)(iv & 1023u), plen = (int)((iv >> 10) & ((1u << 21) - 1u));
            const int left = (int)(uint32_t)(iv >> 32) - key_bias;          // relative to origin_q
            const int a = s_cstart[ch], e = s_cstart[ch + 1];
            const double gpre = a + lane < e ? gain_of(key_idx(s_keys[a + lane])) : 0.0;
            constexpr int kChunkRecs = kFusedTile / WFS_SAMPLES_PER_RECORD;
            const int n_recs = (plen + WFS_SAMPLES_PER_RECORD - 1) / WFS_SAMPLES_PER_RECORD;
            for (int r0 = 0; r0 < n_recs; r0 += kChunkRecs) {
                const int tl = left + r0 * WFS_SAMPLES_PER_RECORD;
                const int n = min(kChunkRecs * WFS_SAMPLES_PER_RECORD, plen - r0 * WFS_SAMPLES_PER_RECORD);
                for (int i = lane; i < n; i += 32) s_tile[i] = 0;
                __syncwarp();
                int pa = a;
                while (pa < e) {
                    const uint64_t pck = s_keys[pa] >> kf.shift_pc;
                    int pe = pa + 1;
                    while (pe < e && (s_keys[pe] >> kf.shift_pc) == pck) pe++;
                    superpose_pulse(s_keys, pa, pe, a, gpre, gain_of, s_tmpl, tlen, tl, tl + n - 1, lane,
                                    [&](int s, double cur, bool on) {
                                        const int i = s - tl;
                                        if (on && cur != 0.0 && i >= 0 && i < n) s_tile[i] += adc_of(cur, c2a, 1);
                                    });
                    __syncwarp();
                    pa = pe;
                }
                const int here = min(kChunkRecs, n_recs - r0);
                for (int r = 0; r < here; r++) {
                    const int rec_i = r0 + r;
                    const int first = tl + r * WFS_SAMPLES_PER_RECORD;
                    const uint32_t key = ((uint32_t)(first + key_bias) << 10) | (uint32_t)ch;
                    int lo = 0, hi = n_rec;                        // rank of the record in the group
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (s_rkey[mid] < key) lo = mid + 1; else hi = mid;
                    }
                    const int64_t dest = rec_base + lo;
                    if (dest >= A.cap_records) continue;
                    const int length = min(plen, WFS_SAMPLES_PER_RECORD * (rec_i + 1)) - WFS_SAMPLES_PER_RECORD * rec_i;
                    const int64_t time = (int64_t)dt * (S.origin_q + first);
                    uint32_t *out = reinterpret_cast<uint32_t *>(A.records_out + dest * WFS_RECORD_BYTES);
                    // strax_interface.py:425-436: time, length, dt, channel, pulse_length, record_i, baseline = 0
                    uint32_t h = (uint32_t)(uint64_t)time;
                    h = lane == 1 ? (uint32_t)((uint64_t)time >> 32) : h;
                    h = lane == 2 ? (uint32_t)length : h;
                    h = lane == 3 ? (((uint32_t)(uint16_t)dt) | ((uint32_t)(uint16_t)ch << 16)) : h;
                    h = lane == 4 ? (uint32_t)plen : h;
                    h = lane == 5 ? (uint32_t)(uint16_t)rec_i : h;
                    const int base_i = r * WFS_SAMPLES_PER_RECORD;
                    auto sample = [&](int j) -> uint32_t {         // record-local sample j -> int16 ADC
                        if (j >= length) return 0u;
                        return (uint32_t)(uint16_t)(int16_t)max(s_tile[base_i + j] + baseline, 0);
                    };
                    // word w of the record: lanes 0..5 header, data word d = w - 6 holds samples 2d, 2d + 1
                    uint32_t w0 = h;
                    if (lane >= 6) { const int d = lane - 6; w0 = sample(2 * d) | (sample(2 * d + 1) << 16); }
                    out[lane] = w0;
                    if (lane + 32 < 61) { const int d = lane + 26; out[lane + 32] = sample(2 * d) | (sample(2 * d + 1) << 16); }
                }
                __syncwarp();
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
bool Backend::fused_eligible(const PhotonBatch &b) const {
    const char *env = getenv("WFS_FUSED");      // WFS_FUSED=0: always the multi-pass back end (A/B tests)
    const bool on = !(env && atoi(env) == 0);
    const DeviceConfig &c = *cfg_;
    if (!on || !b.group_start || !b.group_t0 || !b.group_run0 || !b.instr_run || !b.flags) return false;
    if (b.group_ranges < 1 || b.group_ranges > 4) return false;
    if (b.max_group_photons > kFusedMaxPhotons || b.max_group_photons < 0) return false;
    if (b.relpc_bits < 1 || kShiftSample + kSampleBits + b.relpc_bits + kChannelBits > 64) return false;
    if (c.p.enable_noise && c.noise_t) return false;          // every sample carries noise: dense path
    if (c.he_rows_possible || c.thr_above_baseline) return false;
    if (c.p.dt > 16 || c.p.template_length > 32) return false;
    return true;
}

bool Backend::run_fused(const PhotonBatch &b, uint8_t *records_out, int64_t cap_records,
                        wfs_group_info *group_info_out, BackendResult &res) {
    const DeviceConfig &c = *cfg_;
    const int64_t ng = b.n_groups;
    int n_cap = (int)std::max<int64_t>(1024, (b.max_group_photons + 1023) / 1024 * 1024);
    // 8 warps and two CTAs per SM while the photons of a group leave room for it, else one CTA
    const int threads = kFusedThreads;
    const Layout L = make_layout(n_cap, c.p.n_tpc_pmts, threads / 32, c.p.dt * c.p.template_length);
    if (L.total > 227 * 1024) return false;
    if (!fused_attr_set_ || L.total > fused_smem_set_) {
        WFS_CUDA_CHECK(cudaFuncSetAttribute(k_group_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        fused_attr_set_ = true;
        fused_smem_set_ = L.total;
    }
    int ctas_per_sm = 1;
    WFS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_group_fused, threads, L.total));
    if (ctas_per_sm < 1) return false;
    fused_status_.reserve(sizeof(uint64_t) * (size_t)(ng + 1) + 64);
    uint32_t *ticket = reinterpret_cast<uint32_t *>(fused_status_.as<uint64_t>() + ng);
    fused_scal_.reserve(sizeof(int64_t) * FS_COUNT);
    WFS_CUDA_CHECK(cudaMemsetAsync(fused_status_.p, 0, sizeof(uint64_t) * (size_t)(ng + 1), stream_));
    WFS_CUDA_CHECK(cudaMemsetAsync(fused_scal_.p, 0, sizeof(int64_t) * FS_COUNT, stream_));
    FusedArgs A;
    A.b = b;
    A.c = c;
    A.n_cap = n_cap;
    A.relpc_bits = b.relpc_bits;
    A.group_t0 = b.group_t0;
    A.group_run0 = b.group_run0;
    A.status = fused_status_.as<uint64_t>();
    A.ticket = ticket;
    A.scalars = fused_scal_.as<int64_t>();
    A.group_nitv = nullptr;
    A.records_out = records_out;
    A.cap_records = records_out ? cap_records : 0;
    A.group_info = group_info_out;
    WFS_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
    const int grid = (int)std::min<int64_t>(ng, (int64_t)kNumSMs * ctas_per_sm);
    k_group_fused<<<grid, threads, L.total, stream_>>>(A);
    lc_->n++;
    WFS_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
    WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, A.scalars, sizeof(int64_t) * FS_COUNT, cudaMemcpyDeviceToHost, stream_));
    WFS_CUDA_CHECK(stream_sync(stream_));
    WFS_CUDA_CHECK(cudaGetLastError());
    if (h_scalars_[FS_OVERFLOW]) return false;               // a group outgrew the shared-memory lists
    res = BackendResult();
    if (h_scalars_[FS_ERR]) { res.error = (int)h_scalars_[FS_ERR]; return true; }
    res.fused = 1;
    res.n_valid_photons = h_scalars_[FS_NVALID];
    res.n_pulses = h_scalars_[FS_NPULSES];
    res.n_windows = 2 * h_scalars_[FS_NWIN];
    res.n_intervals = h_scalars_[FS_NITV];
    res.n_samples = h_scalars_[FS_NSAMPLES];
    res.n_records = h_scalars_[FS_NREC];
    res.n_rec_class[0] = res.n_records <= cap_records ? res.n_records : 0;
    WFS_CUDA_CHECK(cudaEventElapsedTime(&res.ms_digitize, ev0_, ev1_));
    res.ms_phase[3] = res.ms_digitize;
    res.segment_sorted_photons = res.segment_sorted_records = 1;
    return true;
}

}  // namespace wfs
