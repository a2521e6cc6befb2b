// Group-resident back end: ONE CTA owns a digitisation group from its photons to its ZLE intervals and record
// order; a second kernel writes the records.  Two launches, count then fill, no pass over dense samples:
//
// k_group_analyse (one CTA per group; groups are binned by photon count so that small groups run many CTAs per SM;
// the classes up to 512 threads run a second build of the body with fewer registers, k_group_analyse_small)
//   photons of the group (generation order, HBM, read once)
//     -> key (channel | pulse call | sample | ns remainder | index | dpe) in shared memory, bucketed by
//        channel (counting sort) and ranked inside the channel by (pulse call, time): a photon's place is the number
//        of smaller keys on its channel                             Pulse.__call__   pulse.py:82-144
//     -> per PHOTON (one thread each): gain into shared memory, the trigger bit of the truth counters
//        (pulse.py:229-271), equal-ns photons merged (pulse.py:301-318), and its class: lone (nothing else reaches
//        its template), plain (only its own pulse call reaches the samples it owns), mixed (another pulse call of the
//        channel does) -- listed apart, so that a warp runs one loop
//     -> the samples the photon OWNS (from its first sample to the next photon's first sample on the channel)
//        evaluated with every photon that overlaps them -- fp64, ascending time, mul and add unfused: the summation
//        order of Pulse.add_current (pulse.py:276-318), one rounding per pulse and sample, -around(current *
//        current_2_adc) (rawdata.py:236-239), baseline, clamp, threshold (rawdata.py:290-296,441-458).  The lists
//        (and the channels where photons pile up, which a warp takes, lane = sample) are handed out warp by warp
//        through a counter.  The ADC values go to the photon's 64-byte slot in HBM; only the first and the last
//        owned sample below threshold are kept (5 + 5 bits in the key): a template is shorter than the ZLE
//        hold-off, so nothing else can change an interval.
//     -> per WINDOW (one lane each; a warp for the rare window with several pulse calls): window extents
//        (pulse.py:118-127, rawdata.py:231-235,258-259), truth counters, the hysteresis interval search
//        (utils.py:13-58, rawdata.py:296-308) over the photons' flagged runs
//     -> record order (time, channel) of the group: counting sort over log-spaced time bins in shared memory, rank
//        inside a bin by counting smaller keys (strax.sort_by_time)
//     -> to HBM: 4 bytes per photon (sample | ns remainder | samples owned) in channel order, 16 bytes per
//        record (channel, first sample, pulse length, record_i, the photons that reach it) at the record's rank
//        in the group, the group's record count
// exclusive scan of the record counts: groups are disjoint in time, so group order IS record order
// k_group_records (one warp per tile of 16 records): a gather -- baseline fill, headers, the owned samples of the
//   photons that reach a record out of their slots; 244-byte records written ONCE, at their final sorted position
//                                                                   strax_interface.py:425-436
// No sort keys, no dense ADC buffer, no flags travel through HBM.  Used when every group of a batch fits
// (<= kFusedMaxPhotons photons, no noise, no high-energy twin rows); anything else takes the multi-pass back
// end (backend.cu), which stays the reference implementation of the same arithmetic for heavy S2s.
#include "backend.cuh"
#include "fused.cuh"

#include <algorithm>
#include <limits.h>
#include <stdlib.h>
#include <string.h>

namespace wfs {

namespace {

constexpr int kNegPos = -(1 << 29);
constexpr int kDenseAhead = 3;                 // photon k + 3 starts inside the template of photon k: a dense channel
constexpr int kNoFlag = 31;                     // f0 field: no owned sample below threshold
constexpr int kRankSortMax = 32;                // photons on a channel ordered by counting smaller keys; more: a warp's bitonic network

struct FusedShared {          // fixed-size part of the shared memory, the arrays follow
    int64_t origin_q;         // absolute sample index of key sample 0
    unsigned long long n_samples;
    int32_t n_valid, n_win, n_itv, n_rec, n_pulses, n_emitted;
    int32_t n_multi, n_slow, n_cw, n_alone, n_mixed, n_long;
    int32_t work;             // next item of the evaluation phase
    int32_t lo, hi;           // min pulse left / max pulse right of the group, relative to origin_q
    int32_t tmax_q;           // largest photon sample of the group, relative to origin_q
    int32_t max_bin;          // most records in one time bin
    int32_t group;
    uint32_t desc_off;        // the group's first record descriptor in the pool
    int32_t overflow;
    int32_t warp_sums[32];
    uint8_t chunk_nz[32], chunk_long[32];      // non-empty / long channels of every chunk of 32 channels
    int32_t trig[2 * kFusedTrigSlots];
};

// bins of the record order for 2^m bins per octave (they double as per-channel flags: at least one per channel)
__host__ __device__ inline int fused_bins(int m, int n_ch) {
    const int n = (22 - m) << m;
    return n > n_ch + 1 ? n : n_ch + 1;
}

struct Layout {               // byte offsets into the dynamic shared memory
    int keys, gains, chan_start, chan_fill, win_ch, multi, tmpl, cmax, itv, rkey, order, bins, total;
};

__host__ __device__ inline Layout make_layout(int n_cap, int itv_cap, int rec_cap, int n_ch, int tmpl_len, int dt, int bin_bits) {
    Layout L;
    int o = (int)((sizeof(FusedShared) + 15) & ~15u);
    L.keys = o; o += 8 * n_cap;
    L.gains = o; o += 8 * n_cap;                 // unsorted keys while loading, gains afterwards
    L.itv = o; o += 8 * itv_cap > ((2 * n_cap + 15) & ~15) ? 8 * itv_cap : ((2 * n_cap + 15) & ~15);      // intervals; before them the list of lone photons (u16 each)
    L.tmpl = o; o += 8 * tmpl_len;
    L.cmax = o; o += 8 * dt;
    L.chan_start = o; o += 4 * (n_ch + 1);
    L.chan_fill = o; o += 4 * (n_ch + 1);
    L.rkey = o; o += 4 * rec_cap;                // record keys, then record ranks
    L.bins = o; o += 4 * fused_bins(bin_bits, n_ch);
    L.order = o; o += 2 * (rec_cap > n_cap ? rec_cap : n_cap);   // photons with neighbours, then record slots
    L.win_ch = o; o += 2 * (n_ch + 2);
    L.multi = o; o += 2 * (n_ch + 2);
    L.total = (o + 15) & ~15;
    return L;
}

__device__ __forceinline__ int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// key = ch << shift_ch | relpc << shift_pc | sample << 18 | rem << 14 | field << 1 | dpe
//   field (13 bits): the photon's index in the group while the channel lists are ordered; afterwards
//   bit 0 above trigger threshold, bit 1 follower of an equal-ns photon, bits 2-6 / 7-11 first / last owned
//   sample below the ZLE threshold (offset from the photon's own sample; first = 31: none), bit 12 first
//   photon of a pulse (channel, pulse call)
constexpr int kShiftSample = 18, kShiftRem = 14, kSampleBits = 20;
constexpr uint64_t kFieldMask = 0x3ffeull;

__device__ __forceinline__ int key_sample(uint64_t k) { return (int)((k >> kShiftSample) & ((1u << kSampleBits) - 1u)); }
__device__ __forceinline__ int key_rem(uint64_t k) { return (int)((k >> kShiftRem) & 15u); }
__device__ __forceinline__ int key_idx(uint64_t k) { return (int)((k >> 1) & 8191u); }
__device__ __forceinline__ int key_dpe(uint64_t k) { return (int)(k & 1u); }
__device__ __forceinline__ int key_above(uint64_t k) { return (int)((k >> 1) & 1u); }
__device__ __forceinline__ bool key_follower(uint64_t k) { return ((k >> 2) & 1u) != 0; }
__device__ __forceinline__ int key_f0(uint64_t k) { return (int)((k >> 3) & 31u); }
__device__ __forceinline__ int key_f1(uint64_t k) { return (int)((k >> 8) & 31u); }
__device__ __forceinline__ uint32_t key_pulse_start(uint64_t k) { return (uint32_t)((k >> 13) & 1u); }

__device__ __forceinline__ int adc_of(double cur, double c2a) {
    return -__double2int_rn(__dmul_rn(cur, c2a));      // one rounding per pulse and sample: rawdata.py:236-239
}

// Time bins of the record order.  The records of a group crowd behind its first photon (key time 0) and thin out
// towards late PMT afterpulses: bins one sample wide up to 2^m, then 2^m bins per octave (m = bin_bits of the class:
// 6 -> 1024 bins for the 2^21 samples of the key time, 8 -> 3584).
constexpr int kTimeKeyBits = 21;
constexpr int kBinMax = 96;                     // records of one bin ranked by counting; more: bitonic network over the group
__device__ __forceinline__ int time_bin(uint32_t t, int m) {
    if (t < (1u << m)) return (int)t;
    const int k = 31 - __clz(t);
    return ((k - m + 1) << m) + (int)((t >> (k - m)) & ((1u << m) - 1u));
}

// interval in shared memory: left + bias (21 bits) | samples (20) | channel (10) | first record slot (13)
__device__ __forceinline__ uint64_t pack_itv(uint32_t lb, uint32_t plen, uint32_t ch, uint32_t r0) {
    return ((uint64_t)lb << 43) | ((uint64_t)plen << 23) | ((uint64_t)ch << 13) | (uint64_t)r0;
}
// photon in HBM for the record kernel: sample << 4 | ns remainder, bits 27-31: samples it owns; their ADC values at
// adc_slots[(first photon of the group + photon) * kSlotVecs]
// record descriptor in HBM (uint4): x = first sample of the record + bias (21) << 10 | channel; y = pulse length |
// record_i (low 12 bits) << 20; z = first photon that reaches the record (13) | photons (14) << 13 |
// record_i >> 12 (2 bits) << 27


// ADC values of the samples a photon owns: a slot of kSlotVecs 16-byte vectors per photon, filled through four
// registers (sample j -> half (j & 1) of word (j >> 1) & 3 of vector j >> 3)
constexpr int kSlotVecs = 4;                    // 32 samples: template_length <= 30
struct SlotWriter {
    uint4 *slot;
    uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
    __device__ __forceinline__ void put(int j, int v) {
        const uint32_t hv = (uint32_t)(uint16_t)(int16_t)v << (16 * (j & 1));
        const int wi = (j >> 1) & 3;
        w0 |= wi == 0 ? hv : 0u; w1 |= wi == 1 ? hv : 0u; w2 |= wi == 2 ? hv : 0u; w3 |= wi == 3 ? hv : 0u;
        if ((j & 7) == 7) flush(j);
    }
    __device__ __forceinline__ void flush(int j) {      // j: the last sample put
        slot[j >> 3] = make_uint4(w0, w1, w2, w3);
        w0 = w1 = w2 = w3 = 0;
    }
    __device__ __forceinline__ void finish(int n) {     // n samples were put
        if (n > 0 && (n & 7) != 0) flush(n - 1);
    }
};

// all-ascending bitonic network on items [0, n): every compare-exchange leaves the smaller item at the lower
// index, so the virtual +inf padding behind n never moves and pairs that reach into it are skipped.
// `first` / `step`: this thread's share of the items; `sync()` separates the stages.
template <typename Less, typename Swap, typename Sync>
__device__ __forceinline__ void bitonic_ascending(int n, int first, int step, Less &&less, Swap &&swap, Sync &&sync) {
    for (int k = 2; (k >> 1) < n; k <<= 1) {
        for (int i = first; i < n; i += step) {
            const int p = i ^ (k - 1);
            if (p > i && p < n && less(p, i)) swap(i, p);
        }
        sync();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int i = first; i < n; i += step) {
                const int p = i ^ j;
                if (p > i && p < n && less(p, i)) swap(i, p);
            }
            sync();
        }
    }
}

}  // namespace

__device__ __forceinline__ void group_analyse(const FusedArgs &A, const FusedClass &K) {
    extern __shared__ __align__(16) uint8_t smem[];
    const PhotonBatch &b = A.b;
    const DeviceConfig &c = A.c;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n_warps = blockDim.x >> 5;
    const int n_ch = c.p.n_tpc_pmts, dt = c.p.dt, tlen = c.p.template_length;
    const int bin_bits = K.bin_bits;
    const Layout L = make_layout(K.n_cap, K.itv_cap, K.rec_cap, n_ch, dt * tlen, dt, bin_bits);
    FusedShared &S = *reinterpret_cast<FusedShared *>(smem);
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem + L.keys);
    double *s_gain = reinterpret_cast<double *>(smem + L.gains);
    uint64_t *s_raw = reinterpret_cast<uint64_t *>(smem + L.keys);       // keys in generation order while loading
    uint64_t *s_buck = reinterpret_cast<uint64_t *>(smem + L.gains);     // bucketed by channel, then ranked back into s_keys
    uint64_t *s_itv = reinterpret_cast<uint64_t *>(smem + L.itv);
    uint16_t *s_alone = reinterpret_cast<uint16_t *>(smem + L.itv);      // lone photons, until the intervals are made
    uint32_t *s_keyb = reinterpret_cast<uint32_t *>(smem + L.gains);     // record keys in bin order (the gains are dead by then)
    double *s_tmpl = reinterpret_cast<double *>(smem + L.tmpl);
    double *s_cmax = reinterpret_cast<double *>(smem + L.cmax);
    int32_t *s_cstart = reinterpret_cast<int32_t *>(smem + L.chan_start);
    int32_t *s_cfill = reinterpret_cast<int32_t *>(smem + L.chan_fill);
    uint32_t *s_rkey = reinterpret_cast<uint32_t *>(smem + L.rkey);
    int32_t *s_bin = reinterpret_cast<int32_t *>(smem + L.bins);
    uint16_t *s_order = reinterpret_cast<uint16_t *>(smem + L.order);
    uint16_t *s_winch = reinterpret_cast<uint16_t *>(smem + L.win_ch);
    uint16_t *s_multi = reinterpret_cast<uint16_t *>(smem + L.multi);

    const double c2a = c.p.current_2_adc;
    const int LM = c.p.pulse_left_margin, RM = c.p.pulse_right_margin, tw = c.p.trigger_window, H = 2 * tw + 1;
    const int baseline = c.p.baseline;
    const int key_bias = LM + tw + 2;                 // record key time = left relative to origin + bias >= 0
    const int shift_pc = kShiftSample + kSampleBits, shift_ch = shift_pc + A.relpc_bits;
    const uint32_t pc_mask = (1u << A.relpc_bits) - 1u;
    const bool smem_trig = (2 << A.relpc_bits) <= 2 * kFusedTrigSlots;     // 2 counters per pulse call of the group
    const bool per_pmt = b.pmt_counts != nullptr;

    for (int i = tid; i < dt * tlen; i += blockDim.x) s_tmpl[i] = c.templates[i];
    for (int i = tid; i < dt; i += blockDim.x) s_cmax[i] = c.current_max[i];
    for (int i = tid; i < fused_bins(bin_bits, n_ch); i += blockDim.x) s_bin[i] = 0;     // (the bins double as per-channel flags below)

    // (the ticket of a group is drawn one group ahead: the round trip of the atomic is over when the group starts)
    uint32_t ticket_next = 0;
    if (tid == 0) ticket_next = atomicAdd(K.ticket, 1u);
    for (;;) {
        __syncthreads();                              // everything of the previous group is done
        if (tid == 0) {
            S.group = ticket_next < K.n_list ? (int32_t)K.list[ticket_next] : -1;
            if (ticket_next < K.n_list) ticket_next = atomicAdd(K.ticket, 1u);
        }
        __syncthreads();
        const int g = S.group;
        if (g < 0) break;
        // ------------------------------------------------------------------ load ----
        if (tid == 0) {
            S.n_valid = S.n_win = S.n_itv = S.n_rec = S.n_pulses = S.n_emitted = 0;
            S.n_multi = S.n_slow = S.n_cw = S.n_alone = S.n_mixed = S.n_long = 0;
            S.work = 0;
            S.lo = INT_MAX; S.hi = INT_MIN;
            S.tmax_q = 0; S.max_bin = 0;
            S.overflow = 0;
            S.n_samples = 0;
            S.origin_q = floordiv64(A.group_t0[g], dt);
        }
        for (int i = tid; i <= n_ch; i += blockDim.x) s_cfill[i] = 0;
        if (smem_trig)
            for (int i = tid; i < (2 << A.relpc_bits); i += blockDim.x) S.trig[i] = 0;
        __syncthreads();
        const int64_t origin_t = S.origin_q * dt;
        const int32_t run0 = A.group_run0[g];
        uint32_t r_lo[4], r_n[4], pbase = 0;
        int n_g = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            r_lo[r] = r_n[r] = 0;
            if (r < b.group_ranges) {
                const uint32_t *gs = b.group_start + (size_t)r * (b.n_groups + 1);
                r_lo[r] = gs[g];
                r_n[r] = gs[g + 1] - gs[g];
                pbase += gs[g] - gs[0];               // photons of the groups in front: this group's place in the sorted arrays
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
        // ADC values of the samples every photon owns
        uint4 *adc_out = A.adc_slots + (size_t)pbase * kSlotVecs;
        if (n_g > K.n_cap) {                           // (the host bins the groups: cannot happen)
            if (tid == 0) A.scalars[FS_OVERFLOW] = 2;
            continue;
        }
        {
            int qmax = 0;
            for (int i = tid; i < n_g; i += blockDim.x) {
                const uint32_t gi = global_index(i);
                const int32_t ch = b.channel[gi];
                const int32_t run = b.instr_run[b.pulse_call[gi]];
                const uint8_t fl = b.flags[gi];
                uint64_t key = ~0ull;
                if (ch >= 0 && ch < n_ch && run >= 0 && c.gains[ch] != 0.0) {
                    const int64_t rel = b.t[gi] - origin_t;           // rel >= 0: origin is a lower bound
                    const uint32_t rel32 = (uint32_t)rel;
                    const uint32_t q = rel32 / (uint32_t)dt;          // (32-bit division: rel is checked against 2^31 below)
                    const int relpc = 2 * (run - run0) + ((fl >> 1) & 1);
                    if (rel < 0 || rel >= ((int64_t)1 << 31) || (int)q >= (1 << kSampleBits) - 2 * (RM + tw + key_bias) ||
                        relpc < 0 || relpc >= (1 << A.relpc_bits) || i >= 8192) {
                        A.scalars[FS_OVERFLOW] = 2;        // outside the key range: the multi-pass back end decides
                    } else {
                        key = ((uint64_t)ch << shift_ch) | ((uint64_t)relpc << shift_pc) |
                              ((uint64_t)q << kShiftSample) | ((uint64_t)(rel32 - q * (uint32_t)dt) << kShiftRem) |
                              ((uint64_t)i << 1) | (uint64_t)(fl & 1);
                        atomicAdd(&s_cfill[ch], 1);
                        qmax = max(qmax, (int)q);
                    }
                }
                s_raw[i] = key;
            }
            qmax = __reduce_max_sync(0xffffffffu, qmax);
            if (lane == 0 && qmax > 0) atomicMax(&S.tmax_q, qmax);
        }
        __syncthreads();
        // time bins of the record order (time_bin: one sample wide at the start of the group, where the records are)
        const int t_key_max = S.tmax_q + RM + tw + key_bias + 1;      // (< 2^kTimeKeyBits: the photon samples are checked above)
        const int n_bins = time_bin((uint32_t)t_key_max, bin_bits) + 1;
        // channel offsets (exclusive scan of the counts), list of non-empty channels, list of long channels: every warp
        // scans chunks of 32 channels, the chunk totals are scanned by every warp for itself
        const int n_chunks = (n_ch + 31) >> 5;            // <= 32 (n_tpc_pmts <= 1023)
        for (int cx = warp; cx < n_chunks; cx += n_warps) {
            const int ch = cx * 32 + lane;
            const int cnt = ch < n_ch ? s_cfill[ch] : 0;
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (ch < n_ch) s_cstart[ch] = inc - cnt;          // offset inside the chunk
            const unsigned m = __ballot_sync(0xffffffffu, cnt > 0);
            const unsigned ml = __ballot_sync(0xffffffffu, cnt > kRankSortMax);
            if (lane == 31) {
                S.warp_sums[cx] = inc;
                S.chunk_nz[cx] = (uint8_t)__popc(m);
                S.chunk_long[cx] = (uint8_t)__popc(ml);
            }
        }
        // (the bins double as per-channel flags below: the previous group's bin offsets must not be read as flags)
        for (int i = tid; i < max(n_bins, n_ch); i += blockDim.x) s_bin[i] = 0;
        __syncthreads();
        {
            const int tot = lane < n_chunks ? S.warp_sums[lane] : 0;
            const int nz = lane < n_chunks ? (int)S.chunk_nz[lane] : 0, nl = lane < n_chunks ? (int)S.chunk_long[lane] : 0;
            int packed_lo = tot, packed_hi = nz | (nl << 16);       // two scans: photons; non-empty | long channels
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, packed_lo, o), u = __shfl_up_sync(0xffffffffu, packed_hi, o);
                if (lane >= o) { packed_lo += v; packed_hi += u; }
            }
            const int ex_lo = packed_lo - tot, ex_hi = packed_hi - (nz | (nl << 16));
            for (int cx = warp; cx < n_chunks; cx += n_warps) {
                const int base = __shfl_sync(0xffffffffu, ex_lo, cx), bh = __shfl_sync(0xffffffffu, ex_hi, cx);
                const int ch = cx * 32 + lane;
                const int cnt = ch < n_ch ? s_cfill[ch] : 0;
                const unsigned m = __ballot_sync(0xffffffffu, cnt > 0);
                const unsigned ml = __ballot_sync(0xffffffffu, cnt > kRankSortMax);      // ordered by a warp below
                if (ch < n_ch) { const int st = base + s_cstart[ch]; s_cstart[ch] = st; s_cfill[ch] = st; }
                if (cnt > 0) s_winch[(bh & 0xffff) + __popc(m & ((1u << lane) - 1u))] = (uint16_t)ch;
                if (cnt > kRankSortMax) s_multi[(bh >> 16) + __popc(ml & ((1u << lane) - 1u))] = (uint16_t)ch;
            }
            if (warp == 0 && lane == 31) {
                s_cstart[n_ch] = packed_lo; S.n_valid = packed_lo; S.n_win = packed_hi & 0xffff; S.n_long = packed_hi >> 16;
            }
        }
        __syncthreads();
        for (int i = tid; i < n_g; i += blockDim.x) {
            const uint64_t key = s_raw[i];
            if (key == ~0ull) continue;
            const int ch = (int)(key >> shift_ch);
            s_buck[atomicAdd(&s_cfill[ch], 1)] = key;
        }
        __syncthreads();
        const int n_valid = S.n_valid, n_win = S.n_win;
        // inside a channel: ascending (pulse call, time, index).  The keys are distinct (they carry the photon's index), so
        // a photon's place is the number of smaller keys on its channel: one thread per photon, from the bucketed copy
        // back into s_keys.  Long lists are ordered in place by a warp (bitonic network) and copied.
        for (int w = warp; w < S.n_long; w += n_warps) {
            const int ch = s_multi[w], a = s_cstart[ch], n = s_cstart[ch + 1] - a;
            uint64_t *k = s_buck + a;
            bitonic_ascending(n, lane, 32, [&](int x, int y) { return k[x] < k[y]; },
                              [&](int x, int y) { const uint64_t t = k[x]; k[x] = k[y]; k[y] = t; },
                              [&]() { __syncwarp(); });
            for (int i = lane; i < n; i += 32) s_keys[a + i] = k[i];
        }
        for (int k = tid; k < n_valid; k += blockDim.x) {
            const uint64_t x = s_buck[k];
            const int ch = (int)(x >> shift_ch), a = s_cstart[ch], e = s_cstart[ch + 1];
            if (e - a > kRankSortMax) continue;
            int r = a;
            for (int j = a; j < e; j++) r += s_buck[j] < x ? 1 : 0;
            s_keys[r] = x;
        }
        __syncthreads();      // the bucketed copy is dead from here on: the area now takes the gains
        // ------------------------------------------------------------------ per photon: gain, trigger bit ----
        for (int k = tid; k < n_valid; k += blockDim.x) {
            const uint64_t key = s_keys[k];
            const int ch = (int)(key >> shift_ch);
            const double gn = b.gain[global_index(key_idx(key))];
            s_gain[k] = gn;
            uint64_t field = (uint64_t)kNoFlag << 2;
            // equal ns on the same pulse: the gains are summed first (pulse.py:301-318) -- by the first of them
            const uint64_t prev = k > s_cstart[ch] ? s_keys[k - 1] : ~key;
            if ((prev >> kShiftRem) == (key >> kShiftRem)) field |= 2u;
            if ((prev >> shift_pc) != (key >> shift_pc)) field |= 1u << 12;
            if (b.trig_dpe_out) {        // pulse.py:229-271
                const double thr_t = (double)(baseline - 1 - c.zle_thr[ch]) - 0.5;
                const bool above = gn * s_cmax[key_rem(key)] * c2a > thr_t;
                if (above) field |= 1u;
                if (per_pmt && !(((uint32_t)(key >> shift_pc) & pc_mask) & 1u)) {
                    const int32_t pc = 2 * run0 + (int)((uint32_t)(key >> shift_pc) & pc_mask);
                    const long long ar = llrint(gn / c.gains[ch] * 4294967296.0);
                    unsigned long long *ar_out = reinterpret_cast<unsigned long long *>(
                        b.pmt_areas + ((int64_t)(pc >> 1) * 2) * (int64_t)n_ch + ch);
                    atomicAdd(ar_out, (unsigned long long)ar);
                    if (above) atomicAdd(ar_out + n_ch, (unsigned long long)ar);
                }
            }
            s_keys[k] = (key & ~kFieldMask) | (field << 1);
        }
        __syncthreads();
        // Leaders take the sum of their followers, followers add nothing any more.  Every leader is classed by what can
        // reach the samples it owns (from its first sample to the next photon's of the same pulse call, at most a template):
        //   lone   -- no other photon of its pulse call within a template, no photon of another call of the channel
        //             reaches it: the template alone (listed here)
        //   plain  -- only photons of its own pulse call reach its samples: summed in time order (listed below, unless a
        //             warp takes the whole channel)
        //   mixed  -- photons of another pulse call of the channel reach its samples (PMT afterpulses on top of the
        //             signal): one rounding per pulse call, integer sum (rawdata.py:236-239)
        // field f1 = 1 (plain) / 2 (mixed) with no first sample marks "to be evaluated".
        for (int kb = warp * 32; kb < n_valid; kb += blockDim.x) {
            const int k = kb + lane;
            const bool valid = k < n_valid;
            const uint64_t key = valid ? s_keys[k] : 0;
            const int ch = (int)(key >> shift_ch);
            bool alone = false;
            if (valid && !key_follower(key)) {
                const int a = s_cstart[ch], e = s_cstart[ch + 1];
                const uint64_t pck = key >> shift_pc;
                const int T = key_sample(key);
                int kn = k + 1;
                if (kn < e && key_follower(s_keys[kn])) {
                    double gsum = s_gain[k];
                    for (; kn < e && key_follower(s_keys[kn]); kn++) gsum = __dadd_rn(gsum, s_gain[kn]);
                    s_gain[k] = gsum;
                }
                const bool single = (s_keys[a] >> shift_pc) == (s_keys[e - 1] >> shift_pc);
                const int Tp = (k > a && (s_keys[k - 1] >> shift_pc) == pck) ? key_sample(s_keys[k - 1]) : kNegPos;
                const int Tn = (kn < e && (s_keys[kn] >> shift_pc) == pck) ? key_sample(s_keys[kn]) : INT_MAX;
                bool clean = true;
                if (!single) {
                    const int s_end = min(T + tlen, Tn);
                    for (int j = a; j < e; j++) {
                        const uint64_t kj = s_keys[j];
                        const int Tj = key_sample(kj);
                        if ((kj >> shift_pc) != pck && Tj > T - tlen && Tj < s_end) { clean = false; break; }
                    }
                }
                alone = clean && kn == k + 1 && Tp <= T - tlen && Tn >= T + tlen;
                if (!alone) s_keys[k] = key | ((uint64_t)(clean ? 1 : 2) << 8);
                // many photons on top of each other: a warp takes the channel below
                if (single && k + kDenseAhead < e && key_sample(s_keys[k + kDenseAhead]) < T + tlen) s_bin[ch] = 1;
            }
            const unsigned m = __ballot_sync(0xffffffffu, alone);
            int base = 0;
            if (m) {
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(&S.n_alone, __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
            }
            if (alone) s_alone[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
        }
        __syncthreads();
        // photons still to be evaluated: listed for one thread each unless a warp takes their channel -- plain ones from
        // the bottom of the list, mixed ones from its top, so that the lanes of a warp run the same loop
        for (int kb = warp * 32; kb < n_valid; kb += blockDim.x) {
            const int k = kb + lane;
            const uint64_t key = k < n_valid ? s_keys[k] : 0;
            if (k < n_valid && key_follower(key)) s_gain[k] = 0.0;
            const bool todo = k < n_valid && key_f0(key) == kNoFlag && !key_follower(key);
            const bool plain = todo && key_f1(key) == 1 && !s_bin[(int)(key >> shift_ch)];
            const bool mixed = todo && key_f1(key) == 2;
            const unsigned m = __ballot_sync(0xffffffffu, plain), m2 = __ballot_sync(0xffffffffu, mixed);
            int base = 0, base2 = 0;
            if (m) {
                const int leader = __ffs(m) - 1;
                if (lane == leader) base = atomicAdd(&S.n_slow, __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
            }
            if (m2) {
                const int leader = __ffs(m2) - 1;
                if (lane == leader) base2 = atomicAdd(&S.n_mixed, __popc(m2));
                base2 = __shfl_sync(0xffffffffu, base2, leader);
            }
            if (plain) s_order[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)k;
            if (mixed) s_order[K.n_cap - 1 - (base2 + __popc(m2 & ((1u << lane) - 1u)))] = (uint16_t)k;
        }
        for (int w = tid; w < n_win; w += blockDim.x) {
            const int ch = s_winch[w];
            if (s_bin[ch]) s_multi[atomicAdd(&S.n_cw, 1)] = (uint16_t)ch;
        }
        __syncthreads();
        // ------------------------------------------------------------------ overlapping photons of one pulse ----
        // One warp per channel.  A cluster = photons whose templates chain (each starts before the previous one ends);
        // its samples are evaluated 32 at a time, lane = sample, summed over the photons that reach it in list (time)
        // order.  Every sample belongs to the last photon that starts at or before it: value into that photon's
        // slot, first / last sample below threshold into its key.
        auto eval_dense_channel = [&](int w) {
            const int ch = s_multi[w], a = s_cstart[ch], e = s_cstart[ch + 1];
            const int thr = c.zle_thr[ch];
            int ks = a;
            while (ks < e) {
                int ke = ks;                                // last photon of the cluster that starts at ks
                for (;;) {
                    const int k = ke + lane;
                    const bool ends = k < e && (k + 1 == e || key_sample(s_keys[k + 1]) - key_sample(s_keys[k]) >= tlen);
                    const unsigned m = __ballot_sync(0xffffffffu, ends);
                    if (m) { ke += __ffs(m) - 1; break; }
                    ke += 32;
                }
                if (ke > ks) {
                    const int s1 = key_sample(s_keys[ke]) + tlen;
                    int jlo = ks;
                    for (int c0 = key_sample(s_keys[ks]); c0 < s1; c0 += 32) {
                        const int sm = c0 + lane;
                        while (key_sample(s_keys[jlo]) + tlen <= c0) jlo++;       // (the cluster reaches s1 > c0: jlo <= ke)
                        double acc = 0.0;
                        int owner = -1, owner_t = 0;
                        for (int j = jlo; j <= ke; j++) {
                            const uint64_t kj = s_keys[j];
                            const int Tj = key_sample(kj);
                            if (Tj > c0 + 31) break;
                            const unsigned d = (unsigned)(sm - Tj);
                            if (d < (unsigned)tlen) {
                                acc = __dadd_rn(acc, __dmul_rn(s_tmpl[key_rem(kj) * tlen + (int)d], s_gain[j]));
                                if (!key_follower(kj)) { owner = j; owner_t = Tj; }
                            }
                        }
                        const bool in = sm < s1 && owner >= 0;
                        const int v = max(adc_of(acc, c2a) + baseline, 0);
                        if (in) reinterpret_cast<uint16_t *>(adc_out + (size_t)owner * kSlotVecs)[sm - owner_t] = (uint16_t)(int16_t)v;
                        const unsigned same = __match_any_sync(0xffffffffu, in ? owner : -1 - lane);
                        const unsigned below = __ballot_sync(0xffffffffu, in && v < thr) & same;
                        if (in && below && lane == __ffs(same) - 1) {
                            const uint64_t key = s_keys[owner];
                            const int f0 = key_f0(key) == kNoFlag ? c0 + (__ffs(below) - 1) - owner_t : key_f0(key);
                            const int f1 = c0 + (31 - __clz(below)) - owner_t;
                            s_keys[owner] = (key & ~((uint64_t)0x3ff << 3)) | ((uint64_t)f0 << 3) | ((uint64_t)f1 << 8);
                        }
                        __syncwarp();
                    }
                }
                ks = ke + 1;
            }
        };
        // ------------------------------------------------------------------ lone photons ----
        // one photon, one pulse: every sample of the template against the threshold
        auto eval_lone = [&](int q) {
            const int k = s_alone[q];
            const uint64_t key = s_keys[k];
            const double gn = s_gain[k];
            const double *tm = s_tmpl + key_rem(key) * tlen;
            const int thr = c.zle_thr[(int)(key >> shift_ch)];
            int f0 = kNoFlag, f1 = 0;
            uint4 *slot = adc_out + (size_t)k * kSlotVecs;
            for (int j0 = 0; j0 < tlen; j0 += 8) {             // eight samples = one 16-byte vector of the slot
                uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const int j = j0 + u;
                    if (j < tlen) {
                        const int v = max(adc_of(__dmul_rn(tm[j], gn), c2a) + baseline, 0);
                        w[u >> 1] |= (uint32_t)(uint16_t)(int16_t)v << (16 * (u & 1));
                        if (v < thr) {
                            f0 = f0 == kNoFlag ? j : f0;
                            f1 = j;
                        }
                    }
                }
                slot[j0 >> 3] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (f0 != kNoFlag)
                s_keys[k] = (key & ~((uint64_t)0x3ff << 3)) | ((uint64_t)f0 << 3) | ((uint64_t)f1 << 8);
        };
        // ------------------------------------------------------------------ plain photons: owned samples, one pulse call ----
        auto eval_plain = [&](int q) {
            const int k = s_order[q];
            const uint64_t key = s_keys[k];
            const int ch = (int)(key >> shift_ch);
            const int a = s_cstart[ch], e = s_cstart[ch + 1];
            const uint64_t pck = key >> shift_pc;
            const int T = key_sample(key);
            const int thr = c.zle_thr[ch];
            int kn = k + 1;
            while (kn < e && key_follower(s_keys[kn])) kn++;
            const int Tn = (kn < e && (s_keys[kn] >> shift_pc) == pck) ? key_sample(s_keys[kn]) : INT_MAX;
            const int s_end = min(T + tlen, Tn);           // owned samples [T, s_end)
            int f0 = kNoFlag, f1 = 0;
            uint4 *slot = adc_out + (size_t)k * kSlotVecs;
            int klo = k;
            while (klo > a && (s_keys[klo - 1] >> shift_pc) == pck && key_sample(s_keys[klo - 1]) > T - tlen) klo--;
            // four owned samples at a time: key, gain and template row of a contributor are read once per four
            // samples; every sample still sums its contributors in list (time) order
            uint32_t w0 = 0, w1 = 0;
            for (int s0 = T; s0 < s_end; s0 += 4) {
                while (key_sample(s_keys[klo]) <= s0 - tlen) klo++;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                for (int j = klo; j <= k; j++) {
                    const uint64_t kj = s_keys[j];
                    const double gj = s_gain[j];
                    const int d = s0 - key_sample(kj);              // >= 0: the contributors start at or before T
                    const double *row = s_tmpl + key_rem(kj) * tlen + d;
                    if (d < tlen) a0 = __dadd_rn(a0, __dmul_rn(row[0], gj));
                    if (d + 1 < tlen) a1 = __dadd_rn(a1, __dmul_rn(row[1], gj));
                    if (d + 2 < tlen) a2 = __dadd_rn(a2, __dmul_rn(row[2], gj));
                    if (d + 3 < tlen) a3 = __dadd_rn(a3, __dmul_rn(row[3], gj));
                }
                const double acc4[4] = {a0, a1, a2, a3};
                uint32_t h[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    h[u] = 0u;
                    if (s0 + u < s_end) {
                        const int v = max(adc_of(acc4[u], c2a) + baseline, 0);
                        h[u] = (uint32_t)(uint16_t)(int16_t)v;
                        if (v < thr) {
                            f0 = f0 == kNoFlag ? s0 + u - T : f0;
                            f1 = s0 + u - T;
                        }
                    }
                }
                // sample j = s0 - T + u sits in half (j & 1) of word (j >> 1) & 3 of vector j >> 3; s0 - T is a multiple of 4
                const int j0 = s0 - T;
                if (j0 & 4) slot[j0 >> 3] = make_uint4(w0, w1, h[0] | (h[1] << 16), h[2] | (h[3] << 16));
                else {
                    w0 = h[0] | (h[1] << 16); w1 = h[2] | (h[3] << 16);
                    if (s0 + 4 >= s_end) slot[j0 >> 3] = make_uint4(w0, w1, 0u, 0u);
                }
            }
            if (f0 != kNoFlag)
                s_keys[k] = (key & ~((uint64_t)0x3ff << 3)) | ((uint64_t)f0 << 3) | ((uint64_t)f1 << 8);
        };
        // ------------------------------------------------------------------ mixed photons: several pulse calls reach the samples ----
        auto eval_mixed = [&](int q) {
            const int k = s_order[K.n_cap - 1 - q];
            const uint64_t key = s_keys[k];
            const int ch = (int)(key >> shift_ch);
            const int a = s_cstart[ch], e = s_cstart[ch + 1];
            const uint64_t pck = key >> shift_pc;
            const int T = key_sample(key);
            const int thr = c.zle_thr[ch];
            int kn = k + 1;
            while (kn < e && key_follower(s_keys[kn])) kn++;
            const int Tn = (kn < e && (s_keys[kn] >> shift_pc) == pck) ? key_sample(s_keys[kn]) : INT_MAX;
            const int s_end = min(T + tlen, Tn);           // owned samples [T, s_end)
            int f0 = kNoFlag, f1 = 0;
            SlotWriter sw{adc_out + (size_t)k * kSlotVecs};
            // one rounding per pulse, integer sum over the pulses
            for (int s = T; s < s_end; s++) {
                int adc = 0;
                double acc = 0.0;
                uint64_t cur = ~0ull;
                for (int j = a; j < e; j++) {
                    const uint64_t kj = s_keys[j];
                    const unsigned d = (unsigned)(s - key_sample(kj));
                    if (d >= (unsigned)tlen) continue;
                    if ((kj >> shift_pc) != cur) { adc += adc_of(acc, c2a); acc = 0.0; cur = kj >> shift_pc; }
                    acc = __dadd_rn(acc, __dmul_rn(s_tmpl[key_rem(kj) * tlen + (int)d], s_gain[j]));
                }
                adc += adc_of(acc, c2a);
                const int v = max(adc + baseline, 0);
                sw.put(s - T, v);
                if (v < thr) {
                    if (f0 == kNoFlag) f0 = s - T;
                    f1 = s - T;
                }
            }
            sw.finish(max(s_end - T, 0));
            if (f0 != kNoFlag)
                s_keys[k] = (key & ~((uint64_t)0x3ff << 3)) | ((uint64_t)f0 << 3) | ((uint64_t)f1 << 8);
        };
        // The work of this phase is handed out warp by warp through a counter, the long items first (a piled-up channel,
        // then 32 mixed, plain, lone photons at a time): a warp that is done takes the next item instead of waiting at
        // the barrier for the warp that drew a piled-up channel.
        {
            const int n_cw = S.n_cw, n_mixed = S.n_mixed, n_plain = S.n_slow, n_lone = S.n_alone;
            const int w1 = n_cw + ((n_mixed + 31) >> 5), w2 = w1 + ((n_plain + 31) >> 5), w3 = w2 + ((n_lone + 31) >> 5);
            for (;;) {
                int w = 0;
                if (lane == 0) w = atomicAdd(&S.work, 1);
                w = __shfl_sync(0xffffffffu, w, 0);
                if (w >= w3) break;
                if (w < n_cw) eval_dense_channel(w);
                else if (w < w1) { const int q = ((w - n_cw) << 5) + lane; if (q < n_mixed) eval_mixed(q); }
                else if (w < w2) { const int q = ((w - w1) << 5) + lane; if (q < n_plain) eval_plain(q); }
                else { const int q = ((w - w2) << 5) + lane; if (q < n_lone) eval_lone(q); }
                __syncwarp();
            }
        }
        for (int w = tid; w < S.n_cw; w += blockDim.x) s_bin[s_multi[w]] = 0;      // the flags are bins again
        __syncthreads();

        // ------------------------------------------------------------------ per window: truth, ZLE intervals ----
        // one ZLE interval of a window: utils.py:44-52, rawdata.py:303-308; s, e window-local
        auto emit = [&](int ch, int wl, int wlen, int s, int e) {
            int l = s - tw, r = e + tw;
            l = max(0, min(l, wlen - 1));
            r = max(0, min(r, wlen - 1));
            l = (l + 1) & ~1;
            r = r & ~1;
            const int plen = max(r - l + 1, 0);
            const int nrec = (plen + WFS_SAMPLES_PER_RECORD - 1) / WFS_SAMPLES_PER_RECORD;
            if (nrec <= 0) return;
            const int slot = atomicAdd(&S.n_itv, 1);
            const int r0 = atomicAdd(&S.n_rec, nrec);
            const int lb = wl + l + key_bias;
            if (slot < K.itv_cap && r0 + nrec <= K.rec_cap && lb >= 0 &&
                lb + WFS_SAMPLES_PER_RECORD * (nrec - 1) <= t_key_max && plen < (1 << 20)) {
                s_itv[slot] = pack_itv((uint32_t)lb, (uint32_t)plen, (uint32_t)ch, (uint32_t)r0);
                // (the records are counted into their time bins further down, one thread per record)
                for (int i = 0; i < nrec; i++)
                    s_rkey[r0 + i] = ((uint32_t)(lb + WFS_SAMPLES_PER_RECORD * i) << 10) | (uint32_t)ch;
            } else {
                S.overflow = 1;
            }
        };
        {
            // windows with one pulse call: one lane each
            int my_lo = INT_MAX, my_hi = INT_MIN, my_pulses = 0, my_emitted = 0;
            unsigned long long my_samples = 0;
            for (int w = tid; w < n_win; w += blockDim.x) {
                const int ch = s_winch[w], a = s_cstart[ch], e = s_cstart[ch + 1];
                const uint64_t kfirst = s_keys[a], klast = s_keys[e - 1];
                if ((kfirst >> shift_pc) != (klast >> shift_pc)) {
                    s_multi[atomicAdd(&S.n_multi, 1)] = (uint16_t)ch;
                    continue;
                }
                const int wl = key_sample(kfirst) - LM - tw;                 // pulse.py:118-127, rawdata.py:258-259
                const int wlen = (key_sample(klast) + RM + tw) - wl + 1;
                if (wlen > kMaxGroupSamples + 1) { A.scalars[FS_ERR] = WFS_E_PULSE_CACHE_TOO_LONG; continue; }
                my_lo = min(my_lo, wl + tw);
                my_hi = max(my_hi, wl + wlen - 1 - tw);
                my_pulses++;
                my_samples += (unsigned long long)wlen;
                const int relpc = (int)((uint32_t)(kfirst >> shift_pc) & pc_mask);
                if (b.trig_dpe_out && !(relpc & 1)) {        // odd calls are PMT afterpulses: no truth
                    int ndpe = 0, n_trig = 0, trig = 0;
                    for (int k = a; k < e; k++) { const uint64_t key = s_keys[k]; ndpe += key_dpe(key); n_trig += key_above(key); }
                    for (int k = a; k < a + ndpe; k++) trig += key_above(s_keys[k]);       // the [:n_double_pe] quirk, pulse.py:255
                    const int32_t pc = 2 * run0 + relpc;
                    if (trig && smem_trig) {          // (more pulse calls than counters: added behind the overflow check)
                        atomicAdd(&S.trig[2 * relpc], trig);
                        if (ch >= c.p.n_top_pmts) atomicAdd(&S.trig[2 * relpc + 1], trig);
                    }
                    if (per_pmt) {
                        int32_t *cnt = b.pmt_counts + ((int64_t)(pc >> 1) * 4) * (int64_t)n_ch + ch;
                        cnt[0] = e - a;
                        cnt[n_ch] = e - a + ndpe;
                        cnt[2 * n_ch] = n_trig;
                        cnt[3 * n_ch] = n_trig + trig;
                    }
                }
                int last = kNegPos, start = kNegPos;
                for (int k = a; k < e; k++) {
                    const uint64_t key = s_keys[k];
                    const int f0 = key_f0(key);
                    if (f0 == kNoFlag) continue;
                    const int p0 = key_sample(key) + f0 - wl, p1 = key_sample(key) + key_f1(key) - wl;
                    if (last == kNegPos) start = p0;
                    else if (p0 - last > H) { emit(ch, wl, wlen, start, last); my_emitted++; start = p0; }
                    last = p1;
                }
                if (last != kNegPos) { emit(ch, wl, wlen, start, last); my_emitted++; }
            }
            my_lo = __reduce_min_sync(0xffffffffu, my_lo);
            my_hi = __reduce_max_sync(0xffffffffu, my_hi);
            my_pulses = __reduce_add_sync(0xffffffffu, my_pulses);
            my_emitted = __reduce_add_sync(0xffffffffu, my_emitted);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) my_samples += __shfl_xor_sync(0xffffffffu, my_samples, o);
            if (lane == 0 && my_pulses) {
                atomicMin(&S.lo, my_lo);
                atomicMax(&S.hi, my_hi);
                atomicAdd(&S.n_pulses, my_pulses);
                atomicAdd(&S.n_emitted, my_emitted);
                atomicAdd(&S.n_samples, my_samples);
            }
        }
        __syncthreads();
        // windows with several pulse calls: one warp each
        for (int w = warp; w < S.n_multi; w += n_warps) {
            const int ch = s_multi[w], a = s_cstart[ch], e = s_cstart[ch + 1];
            int qmin = INT_MAX, qmax = INT_MIN, np = 0;
            for (int k0 = a; k0 < e; k0 += 32) {
                const int k = k0 + lane;
                const bool in = k < e;
                const uint64_t key = in ? s_keys[k] : 0;
                np += __popc(__ballot_sync(0xffffffffu, in && key_pulse_start(key)));
                qmin = min(qmin, in ? key_sample(key) : INT_MAX);
                qmax = max(qmax, in ? key_sample(key) : INT_MIN);
            }
            qmin = __reduce_min_sync(0xffffffffu, qmin);
            qmax = __reduce_max_sync(0xffffffffu, qmax);
            const int wl = qmin - LM - tw;
            const int wlen = (qmax + RM + tw) - wl + 1;
            if (wlen > kMaxGroupSamples + 1) {
                if (lane == 0) A.scalars[FS_ERR] = WFS_E_PULSE_CACHE_TOO_LONG;
                continue;
            }
            if (b.trig_dpe_out) {
                int pa = a;
                while (pa < e) {
                    const uint64_t pck = s_keys[pa] >> shift_pc;
                    int pe = pa + 1;
                    while (pe < e && (s_keys[pe] >> shift_pc) == pck) pe++;
                    const int relpc = (int)((uint32_t)pck & pc_mask);
                    if (!(relpc & 1)) {
                        int ndpe = 0, n_trig = 0, trig = 0;
                        for (int k = pa + lane; k < pe; k += 32) { const uint64_t key = s_keys[k]; ndpe += key_dpe(key); n_trig += key_above(key); }
                        ndpe = __reduce_add_sync(0xffffffffu, ndpe);
                        n_trig = __reduce_add_sync(0xffffffffu, n_trig);
                        for (int k = pa + lane; k < pa + ndpe; k += 32) trig += key_above(s_keys[k]);
                        trig = __reduce_add_sync(0xffffffffu, trig);
                        const int32_t pc = 2 * run0 + relpc;
                        if (lane == 0) {
                            if (trig && smem_trig) {
                                atomicAdd(&S.trig[2 * relpc], trig);
                                if (ch >= c.p.n_top_pmts) atomicAdd(&S.trig[2 * relpc + 1], trig);
                            }
                            if (per_pmt) {
                                int32_t *cnt = b.pmt_counts + ((int64_t)(pc >> 1) * 4) * (int64_t)n_ch + ch;
                                cnt[0] = pe - pa;
                                cnt[n_ch] = pe - pa + ndpe;
                                cnt[2 * n_ch] = n_trig;
                                cnt[3 * n_ch] = n_trig + trig;
                            }
                        }
                    }
                    pa = pe;
                }
            }
            // flagged runs in ascending (first sample, list position): selection by the warp
            int last = kNegPos, start = kNegPos, n_emitted = 0;
            unsigned long long cursor = 0;          // (p0 + 1) << 14 | position + 1 of the last run taken
            for (;;) {
                unsigned long long best = ~0ull;
                for (int k = a + lane; k < e; k += 32) {
                    const uint64_t key = s_keys[k];
                    if (key_f0(key) == kNoFlag) continue;
                    const unsigned long long cand =
                        ((unsigned long long)(key_sample(key) + key_f0(key) - wl + 1) << 14) | (unsigned long long)(k - a + 1);
                    if (cand > cursor && cand < best) best = cand;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
                    best = v < best ? v : best;
                }
                if (best == ~0ull) break;
                cursor = best;
                const uint64_t key = s_keys[a + (int)(best & 16383u) - 1];
                const int p0 = key_sample(key) + key_f0(key) - wl, p1 = key_sample(key) + key_f1(key) - wl;
                if (last == kNegPos) { start = p0; last = p1; }
                else if (p0 - last > H) {
                    if (lane == 0) emit(ch, wl, wlen, start, last);
                    n_emitted++;
                    start = p0; last = p1;
                } else last = max(last, p1);
            }
            if (last != kNegPos) {
                if (lane == 0) emit(ch, wl, wlen, start, last);
                n_emitted++;
            }
            if (lane == 0) {
                atomicMin(&S.lo, wl + tw);
                atomicMax(&S.hi, wl + wlen - 1 - tw);
                atomicAdd(&S.n_pulses, np);
                atomicAdd(&S.n_emitted, n_emitted);
                atomicAdd(&S.n_samples, (unsigned long long)wlen);
            }
        }
        __syncthreads();
        // ------------------------------------------------------------------ record order ----
        if (S.overflow) {
            // the group outgrew the interval / record lists of its class: listed for the largest class
            if (tid == 0) {
                if (K.overflow_list) K.overflow_list[atomicAdd(A.over_count, 1u)] = (uint32_t)g;
                else A.scalars[FS_OVERFLOW] = 1;
            }
            continue;
        }
        const int n_rec = S.n_rec, n_itv = S.n_itv;
        if (tid == 0) {
            // the group's descriptors: a block of the pool (any order; the scan over n_rec orders the records)
            const unsigned long long off = atomicAdd((unsigned long long *)&A.scalars[FS_NREC], (unsigned long long)n_rec);
            S.desc_off = (uint32_t)off;
            if (off + (unsigned long long)n_rec > (unsigned long long)A.cap_records) S.desc_off = 0xffffffffu;
            A.group_nrec[g] = (uint32_t)n_rec;
            A.group_desc[g] = S.desc_off;
        }
        if (smem_trig && b.trig_dpe_out)
            for (int i = tid; i < (2 << A.relpc_bits); i += blockDim.x)
                if (S.trig[i]) atomicAdd(&b.trig_dpe_out[2 * (2 * run0) + i], S.trig[i]);
        if (!smem_trig && b.trig_dpe_out) {
            // more pulse calls in the group than shared-memory counters: the trigger counts go straight to HBM, but
            // only now that the group is known to fit -- a group that is repeated with larger lists must not have
            // counted already (pulse.py:229-271; one thread per channel walks the channel's pulses)
            for (int w = tid; w < n_win; w += blockDim.x) {
                const int ch = s_winch[w], a = s_cstart[ch], e = s_cstart[ch + 1];
                int pa = a;
                while (pa < e) {
                    const uint64_t pck = s_keys[pa] >> shift_pc;
                    int pe = pa + 1, ndpe = key_dpe(s_keys[pa]);
                    while (pe < e && (s_keys[pe] >> shift_pc) == pck) { ndpe += key_dpe(s_keys[pe]); pe++; }
                    const int relpc = (int)((uint32_t)pck & pc_mask);
                    if (!(relpc & 1)) {
                        int trig = 0;
                        for (int k = pa; k < pa + ndpe; k++) trig += key_above(s_keys[k]);      // the [:n_double_pe] quirk
                        if (trig) {
                            const int32_t pc = 2 * run0 + relpc;
                            atomicAdd(&b.trig_dpe_out[2 * pc], trig);
                            if (ch >= c.p.n_top_pmts) atomicAdd(&b.trig_dpe_out[2 * pc + 1], trig);
                        }
                    }
                    pa = pe;
                }
            }
        }
        // the records counted into their time bins
        for (int slot = tid; slot < n_rec; slot += blockDim.x) {
            const int cnt = atomicAdd(&s_bin[time_bin(s_rkey[slot] >> 10, bin_bits)], 1) + 1;
            if (cnt > kBinMax) atomicMax(&S.max_bin, cnt);
        }
        // the photons in channel order for the record kernel; bits 27-31: samples the photon owns
        for (int k = tid; k < n_valid; k += blockDim.x) {
            const uint64_t key = s_keys[k];
            int n_own = 0;
            if (!key_follower(key)) {
                const int e = s_cstart[(int)(key >> shift_ch) + 1], T = key_sample(key);
                int kn = k + 1;
                while (kn < e && key_follower(s_keys[kn])) kn++;
                const int Tn = (kn < e && !key_pulse_start(s_keys[kn])) ? key_sample(s_keys[kn]) : INT_MAX;
                n_own = min(T + tlen, Tn) - T;
            }
            A.tkey[pbase + k] = (uint32_t)((key >> kShiftRem) & 0xffffffu) | ((uint32_t)n_own << 27);
        }
        __syncthreads();
        if (S.max_bin <= kBinMax && K.rec_cap <= 2 * K.n_cap) {       // (the bin-ordered keys take the place of the gains)
            // counting sort over the time bins; inside its bin a record's place is the number of smaller (time, channel)
            // keys: one thread per record
            const int ipt = (n_bins + (int)blockDim.x - 1) / (int)blockDim.x;
            const int b0 = min(tid * ipt, n_bins), b1 = min(b0 + ipt, n_bins);
            int sum = 0;
            for (int i = b0; i < b1; i++) sum += s_bin[i];
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane == 31) S.warp_sums[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int v = lane < n_warps ? S.warp_sums[lane] : 0, iv = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int u = __shfl_up_sync(0xffffffffu, iv, o);
                    if (lane >= o) iv += u;
                }
                S.warp_sums[lane] = iv - v;
            }
            __syncthreads();
            int run = S.warp_sums[warp] + inc - sum;
            for (int i = b0; i < b1; i++) { const int cnt = s_bin[i]; s_bin[i] = run; run += cnt; }
            __syncthreads();
            for (int slot = tid; slot < n_rec; slot += blockDim.x) {
                const uint32_t key = s_rkey[slot];
                const int pos = atomicAdd(&s_bin[time_bin(key >> 10, bin_bits)], 1);      // (s_bin[b]: now the END of bin b)
                s_order[pos] = (uint16_t)slot;
                s_keyb[pos] = key;
            }
            __syncthreads();
            for (int i = tid; i < n_rec; i += blockDim.x) {
                const uint32_t kx = s_keyb[i];
                const int bi = time_bin(kx >> 10, bin_bits);
                const int lo = bi ? s_bin[bi - 1] : 0, hi = s_bin[bi];
                int r = lo;
                for (int j = lo; j < hi; j++) {
                    const uint32_t ky = s_keyb[j];
                    r += (ky < kx || (ky == kx && j < i)) ? 1 : 0;
                }
                s_rkey[s_order[i]] = (uint32_t)r;          // key -> rank (the keys are read from the bin-ordered copy)
            }
        } else {
            // many records in one bin (a long group): bitonic network over all records of the group
            for (int i = tid; i < n_rec; i += blockDim.x) s_order[i] = (uint16_t)i;
            __syncthreads();
            bitonic_ascending(n_rec, tid, blockDim.x, [&](int x, int y) { return s_rkey[s_order[x]] < s_rkey[s_order[y]]; },
                              [&](int x, int y) { const uint16_t t = s_order[x]; s_order[x] = s_order[y]; s_order[y] = t; },
                              [&]() { __syncthreads(); });
            // the network's last barrier is behind us: nobody reads a key any more, ranks replace them
            for (int i = tid; i < n_rec; i += blockDim.x) s_rkey[s_order[i]] = (uint32_t)i;
        }
        __syncthreads();
        // ------------------------------------------------------------------ descriptors, group bookkeeping ----
        if (S.desc_off != 0xffffffffu && A.desc) {
            for (int it = tid; it < n_itv; it += blockDim.x) {
                const uint64_t iv = s_itv[it];
                const int ch = (int)((iv >> 13) & 1023u), plen = (int)((iv >> 23) & ((1u << 20) - 1u));
                const uint32_t lb = (uint32_t)(iv >> 43);
                const int left = (int)lb - key_bias;                        // relative to origin_q
                const int r0 = (int)(iv & 8191u);
                const int a = s_cstart[ch], e = s_cstart[ch + 1];
                const bool single = (s_keys[a] >> shift_pc) == (s_keys[e - 1] >> shift_pc);
                const int n_recs = (plen + WFS_SAMPLES_PER_RECORD - 1) / WFS_SAMPLES_PER_RECORD;
                int ka = a, kb = a;
                for (int r = 0; r < n_recs; r++) {
                    const int first = left + r * WFS_SAMPLES_PER_RECORD;
                    const int length = min(plen - r * WFS_SAMPLES_PER_RECORD, WFS_SAMPLES_PER_RECORD);
                    if (single) {          // time order: the photons that reach the record are one run of the list
                        while (ka < e && key_sample(s_keys[ka]) + tlen <= first) ka++;
                        kb = max(kb, ka);
                        while (kb < e && key_sample(s_keys[kb]) < first + length) kb++;
                    } else { ka = a; kb = e; }
                    uint4 d;
                    d.x = ((lb + (uint32_t)(r * WFS_SAMPLES_PER_RECORD)) << 10) | (uint32_t)ch;
                    d.y = (uint32_t)plen | ((uint32_t)(r & 0xfff) << 20);
                    d.z = (kb > ka ? (uint32_t)ka : 0u) | ((uint32_t)(kb - ka) << 13) | ((uint32_t)(r >> 12) << 27);
                    d.w = 0;
                    A.desc[(size_t)S.desc_off + s_rkey[r0 + r]] = d;
                }
            }
        }
        if (tid == 0) {
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
}

// Two builds of the same body: the classes of large groups run 1024 threads per CTA (64 registers each); the small
// classes (<= 512 threads) are bound by barrier waits, not by registers, and trade registers for 1536 resident threads.
__global__ void __launch_bounds__(kFusedThreads, 1)
k_group_analyse(FusedArgs A, FusedClass K) { group_analyse(A, K); }

__global__ void __launch_bounds__(kFusedSmallThreads, 3)
k_group_analyse_small(FusedArgs A, FusedClass K) { group_analyse(A, K); }

// Records of a group, a tile of 16 consecutive records at a time, one WARP per tile: assembled in the warp's own
// slice of shared memory and streamed out with aligned 16-byte stores -- baseline fill, then headers and zero
// padding behind the pulse (one lane per record), and, one lane per (record, photon that reaches it), the ADC
// values of the samples that photon owns inside the record, as k_group_analyse left them.  No arithmetic on samples
// here: the kernel is a gather.  Warps never wait for each other (no CTA barrier): the dependent loads of one
// tile (descriptor -> photon word -> slot) overlap with the stores of the others.
constexpr int kTileRecs = 16, kTileWords = kTileRecs * 61 + 8, kRecordWarps = kFusedRecordThreads / 32;

__global__ void __launch_bounds__(kFusedRecordThreads)
k_group_records(FusedArgs A) {
    __shared__ __align__(16) uint32_t s_tile_all[kRecordWarps][kTileWords];
    __shared__ uint4 s_desc_all[kRecordWarps][kTileRecs];
    __shared__ int s_pref_all[kRecordWarps][kTileRecs + 1];
    const PhotonBatch &b = A.b;
    const DeviceConfig &c = A.c;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dt = c.p.dt;
    const int baseline = c.p.baseline;
    const int key_bias = c.p.pulse_left_margin + c.p.trigger_window + 2;
    const uint32_t fill = (uint32_t)(uint16_t)(int16_t)max(baseline, 0), fill2 = fill | (fill << 16);
    constexpr int SPR = WFS_SAMPLES_PER_RECORD;
    uint32_t *s_tile = s_tile_all[warp];
    uint4 *s_desc = s_desc_all[warp];
    int *s_pref = s_pref_all[warp];
    uint16_t *s_tile16 = reinterpret_cast<uint16_t *>(s_tile);
    for (int g = blockIdx.x; g < (int)b.n_groups; g += gridDim.x) {
        const uint32_t n_rec = A.group_nrec[g], desc_off = A.group_desc[g];
        if (n_rec == 0 || desc_off == 0xffffffffu) continue;
        const uint32_t first_tile = (blockIdx.y * kRecordWarps + warp) * kTileRecs;
        if (first_tile >= n_rec) continue;
        const int64_t rec_base = (int64_t)A.rec_base[g];
        const int64_t origin_q = floordiv64(A.group_t0[g], dt);
        uint32_t pbase = 0;
        for (int r = 0; r < b.group_ranges; r++) {
            const uint32_t *gs = b.group_start + (size_t)r * (b.n_groups + 1);
            pbase += gs[g] - gs[0];
        }
        const uint32_t *tkey = A.tkey + pbase;
        const uint4 *adc = A.adc_slots + (size_t)pbase * kSlotVecs;
        const uint32_t tile_step = gridDim.y * kRecordWarps * kTileRecs;
        uint4 d_next = make_uint4(0u, 0u, 0u, 0u);
        if (first_tile + lane < n_rec && lane < kTileRecs) d_next = A.desc[(size_t)desc_off + first_tile + lane];
        for (uint32_t t0 = first_tile; t0 < n_rec; t0 += tile_step) {
            const int64_t dest0 = rec_base + t0;
            const int nr = (int)min((int64_t)min(n_rec - t0, (uint32_t)kTileRecs), A.cap_records - dest0);
            if (nr <= 0) break;
            uint4 d = d_next;                             // this tile's descriptors, loaded while the previous tile was assembled
            if (t0 + tile_step + lane < n_rec && lane < kTileRecs) d_next = A.desc[(size_t)desc_off + t0 + tile_step + lane];
            const int hw = (int)(((uint64_t)dest0 * WFS_RECORD_BYTES) & 15u) >> 2;      // words in front of the tile in its first 16-byte vector
            const int n_words = nr * 61, n_vec = (hw + n_words + 3) >> 2;
            __syncwarp();                                 // the previous tile has left
            if (lane < nr) s_desc[lane] = d;
            else d = make_uint4(0u, 0u, 0u, 0u);
            {
                uint4 *v = reinterpret_cast<uint4 *>(s_tile);
                const uint4 f4 = make_uint4(fill2, fill2, fill2, fill2);
                for (int i = lane; i < n_vec; i += 32) v[i] = f4;
            }
            // photons that reach every record: prefix over the tile
            {
                const int n0 = lane < nr ? (int)((d.z >> 13) & 16383u) : 0;
                int i0 = n0;
#pragma unroll
                for (int o = 1; o < kTileRecs; o <<= 1) {
                    const int u0 = __shfl_up_sync(0xffffffffu, i0, o);
                    if (lane >= o) i0 += u0;
                }
                if (lane < kTileRecs) s_pref[lane] = i0 - n0;
                if (lane == kTileRecs - 1) s_pref[kTileRecs] = i0;
            }
            __syncwarp();
            if (lane < nr) {
                // strax_interface.py:425-436: time, length, dt, channel, pulse_length, record_i, baseline = 0
                const int ch = (int)(d.x & 1023u), plen = (int)(d.y & 0xfffffu);
                const int rec_i = (int)(((d.y >> 20) & 0xfffu) | (((d.z >> 27) & 3u) << 12));
                const int64_t time = (int64_t)dt * (origin_q + (int)(d.x >> 10) - key_bias);
                const int length = min(plen - rec_i * SPR, SPR);
                uint32_t *h = s_tile + hw + lane * 61;
                h[0] = (uint32_t)(uint64_t)time;
                h[1] = (uint32_t)((uint64_t)time >> 32);
                h[2] = (uint32_t)length;
                h[3] = ((uint32_t)(uint16_t)dt) | ((uint32_t)(uint16_t)ch << 16);
                h[4] = (uint32_t)plen;
                h[5] = (uint32_t)(uint16_t)rec_i;
                if (length < SPR && (length & 1)) h[6 + (length >> 1)] &= 0xffffu;
            }
            // zeros behind `length` (the last record of a pulse): two lanes per record, a half of the words each, 16-byte
            // stores between the unaligned ends (one lane walking ~45 words was a seventh of the kernel's instructions)
            if ((lane & (kTileRecs - 1)) < nr) {
                const int rl = lane & (kTileRecs - 1);
                const uint4 dd = s_desc[rl];
                const int plen = (int)(dd.y & 0xfffffu);
                const int rec_i = (int)(((dd.y >> 20) & 0xfffu) | (((dd.z >> 27) & 3u) << 12));
                const int length = min(plen - rec_i * SPR, SPR);
                if (length < SPR) {
                    const int w_lo = (length + 1) >> 1, w_mid = (w_lo + SPR / 2) >> 1;
                    uint32_t *p = s_tile + hw + rl * 61 + 6 + (lane < kTileRecs ? w_lo : w_mid);
                    int n = lane < kTileRecs ? w_mid - w_lo : SPR / 2 - w_mid;
                    for (; n > 0 && ((uintptr_t)p & 15u); n--) *p++ = 0u;
                    for (; n >= 4; n -= 4, p += 4) *reinterpret_cast<uint4 *>(p) = make_uint4(0u, 0u, 0u, 0u);
                    for (; n > 0; n--) *p++ = 0u;
                }
            }
            __syncwarp();
            const int n_pairs = s_pref[kTileRecs];
            for (int q = lane; q < n_pairs; q += 32) {
                int r = 0;                                  // the record of pair q: last r with s_pref[r] <= q
#pragma unroll
                for (int o = kTileRecs >> 1; o > 0; o >>= 1)
                    if (r + o < nr && s_pref[r + o] <= q) r += o;
                const uint4 dd = s_desc[r];
                const int ka = (int)(dd.z & 8191u);
                const int k = ka + (q - s_pref[r]);
                // the photon word and its whole slot at once: neither address depends on the other load
                const uint4 *slot = adc + (size_t)k * kSlotVecs;
                const uint32_t tk = __ldg(tkey + k);
                const uint4 q0 = __ldg(slot), q1 = __ldg(slot + 1), q2 = __ldg(slot + 2), q3 = __ldg(slot + 3);
                const int n_own = (int)(tk >> 27);
                if (n_own == 0) continue;                   // its gain went to the first photon of the same ns
                const int plen = (int)(dd.y & 0xfffffu);
                const int rec_i = (int)(((dd.y >> 20) & 0xfffu) | (((dd.z >> 27) & 3u) << 12));
                const int first = (int)(dd.x >> 10) - key_bias;              // relative to origin_q
                const int length = min(plen - rec_i * SPR, SPR);
                const int T = (int)((tk >> 4) & 0xfffffu);
                uint16_t *rec16 = s_tile16 + 2 * (hw + r * 61 + 6);
                const int j0 = max(0, first - T), j1 = min(n_own, first + length - T);
                if (j0 >= j1) continue;
                uint16_t *dst = rec16 + (T - first);
                const uint32_t w[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
                for (int j = 0; j < 8 * kSlotVecs; j++)
                    if (j >= j0 && j < j1) dst[j] = (uint16_t)(w[j >> 1] >> (16 * (j & 1)));
            }
            __syncwarp();
            // whole 16-byte vectors to their final place, single words at the ragged ends
            {
                uint32_t *out_w = reinterpret_cast<uint32_t *>(A.records_out) + dest0 * 61 - hw;      // word 0 of the tile's first vector
                const int v0 = hw ? 1 : 0, v1 = (hw + n_words) >> 2;
                const uint4 *sv = reinterpret_cast<const uint4 *>(s_tile);
                uint4 *ov = reinterpret_cast<uint4 *>(out_w);
                for (int i = v0 + lane; i < v1; i += 32) ov[i] = sv[i];
                if (hw && lane >= hw && lane < 4 && lane < hw + n_words) out_w[lane] = s_tile[lane];
                const int wt = lane + 4 * v1;
                if (v1 >= v0 && lane < 4 && wt < hw + n_words && wt >= hw) out_w[wt] = s_tile[wt];
            }
        }
    }
}

// "Group order IS record order" needs the groups to follow each other in time.  The reference scheduler can hand out
// a group of delayed secondaries that starts before the previous group has ended (rawdata.py:87-98 looks at the
// clusters' start times): then the records of the two interleave and the batch takes the multi-pass back end, which
// orders the records of a batch globally.
__global__ void k_groups_disjoint(int64_t n_groups, const wfs_group_info *__restrict__ info, int64_t *scalars) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups || g == 0 || info[g].n_intervals <= 0) return;
    int64_t p = g - 1;
    while (p >= 0 && info[p].n_intervals <= 0) p--;          // the last group in front that has records
    if (p >= 0 && info[g].left <= info[p].right) scalars[FS_OVERFLOW] = 3;
}

// ---------------------------------------------------------------------------------------------
bool Backend::fused_eligible(const PhotonBatch &b) const {
    const char *env = getenv("WFS_FUSED");      // WFS_FUSED=0: always the multi-pass back end (A/B tests)
    const bool on = !(env && atoi(env) == 0);
    const DeviceConfig &c = *cfg_;
    if (!on || !b.group_start || !b.h_group_start || !b.group_t0 || !b.group_run0 || !b.instr_run || !b.flags) return false;
    if (b.group_ranges < 1 || b.group_ranges > 4) return false;
    if (b.max_group_photons > kFusedMaxPhotons || b.max_group_photons < 0) return false;
    if (b.relpc_bits < 1 || kShiftSample + kSampleBits + b.relpc_bits + kChannelBits > 64) return false;
    if (c.p.enable_noise && c.noise_t) return false;          // every sample carries noise: dense path
    if (c.he_rows_possible || c.thr_above_baseline) return false;
    if (c.p.dt > 16 || c.p.template_length > 30 || c.p.n_tpc_pmts > 1023) return false;
    if (2 * c.p.trigger_window + 1 < c.p.template_length) return false;   // flagged runs of a photon may then split an interval
    return true;
}

bool Backend::run_fused(const PhotonBatch &b, uint8_t *records_out, int64_t cap_records,
                        wfs_group_info *group_info_out, BackendResult &res) {
    const DeviceConfig &c = *cfg_;
    const int64_t ng = b.n_groups;
    const int n_ch = c.p.n_tpc_pmts, tmpl_len = c.p.dt * c.p.template_length;
    // groups by photon count: small groups get small shared-memory lists and many CTAs per SM
    struct ClassDef { int n_cap, itv_cap, rec_cap, threads, bin_bits; };
    ClassDef defs[kFusedMaxClasses] = {{256, 256, 512, 128, 6}, {512, 512, 1024, 256, 7}, {2048, 1024, 3072, 512, 7},
                                       {4096, 2048, 6144, 1024, 8},
                                       {kFusedMaxPhotons, 4096, kFusedMaxRecCap, 1024, 7}};      // (profiles/tools/class_sweep.sh)
    int kFusedClasses = 5;
    if (const char *e = getenv("WFS_FUSED_CLASSES")) {        // experiments: "photons:intervals:records:threads[:bin_bits],..." ascending, the last one catches all
        int n = 0;
        const char *p = e;
        while (n < kFusedMaxClasses) {
            ClassDef d;
            d.bin_bits = 6;
            if (sscanf(p, "%d:%d:%d:%d:%d", &d.n_cap, &d.itv_cap, &d.rec_cap, &d.threads, &d.bin_bits) < 4) break;
            d.bin_bits = std::max(6, std::min(d.bin_bits, 8));
            defs[n++] = d;
            p = strchr(p, ',');
            if (!p) break;
            p++;
        }
        if (n > 0 && defs[n - 1].n_cap == kFusedMaxPhotons && defs[n - 1].rec_cap <= kFusedMaxRecCap) kFusedClasses = n;
        else throw std::runtime_error("WFS_FUSED_CLASSES: the last class must hold 8192 photons");
    }
    if (const char *e = getenv("WFS_FUSED_REC_CAP"))          // tests: small lists, so that groups take the second attempt
        for (int k = 0; k + 1 < kFusedClasses; k++) defs[k].rec_cap = std::max(32, std::min(defs[k].rec_cap, atoi(e)));
    std::vector<uint32_t> list((size_t)ng);
    uint32_t cls_n[kFusedMaxClasses] = {}, cls_off[kFusedMaxClasses + 1];
    std::vector<uint8_t> cls_of((size_t)ng);
    for (int64_t g = 0; g < ng; g++) {
        int64_t n = 0;
        for (int r = 0; r < b.group_ranges; r++) {
            const uint32_t *gs = b.h_group_start + (size_t)r * (ng + 1);
            n += gs[g + 1] - gs[g];
        }
        int k = 0;
        while (k + 1 < kFusedClasses && n > defs[k].n_cap) k++;
        if (n > defs[k].n_cap) return false;
        cls_of[g] = (uint8_t)k;
        cls_n[k]++;
    }
    cls_off[0] = 0;
    for (int k = 0; k < kFusedClasses; k++) cls_off[k + 1] = cls_off[k] + cls_n[k];
    {
        uint32_t fill[kFusedMaxClasses];
        for (int k = 0; k < kFusedClasses; k++) fill[k] = cls_off[k];
        for (int64_t g = 0; g < ng; g++) list[fill[cls_of[g]]++] = (uint32_t)g;
    }
    // workspaces: [ng] group list, [ng] overflow list, [ng + 1] record counts, [ng + 1] record bases, [ng] descriptor offsets
    fused_lists_.reserve(sizeof(uint32_t) * (size_t)(5 * ng + 8));
    uint32_t *d_list = fused_lists_.as<uint32_t>(), *d_over = d_list + ng, *d_nrec = d_over + ng,
             *d_base = d_nrec + ng + 1, *d_desc_off = d_base + ng + 1;
    fused_scal_.reserve(sizeof(int64_t) * (FS_COUNT + 8));       // + overflow count and one ticket per launch
    int64_t *d_scal = fused_scal_.as<int64_t>();
    uint32_t *d_u32 = reinterpret_cast<uint32_t *>(d_scal + FS_COUNT);        // [0] overflow count, [1..] tickets
    fused_tkey_.reserve(sizeof(uint32_t) * (size_t)std::max<int64_t>(b.n, 1));
    fused_gain_.reserve(sizeof(uint4) * kSlotVecs * (size_t)std::max<int64_t>(b.n, 1));
    const bool want = records_out != nullptr && cap_records > 0;
    if (want) fused_desc_.reserve(sizeof(uint4) * (size_t)cap_records);
    WFS_CUDA_CHECK(cudaMemcpyAsync(d_list, list.data(), sizeof(uint32_t) * (size_t)ng, cudaMemcpyHostToDevice, stream_));
    WFS_CUDA_CHECK(cudaMemsetAsync(d_nrec, 0, sizeof(uint32_t) * (size_t)(ng + 1), stream_));
    WFS_CUDA_CHECK(cudaMemsetAsync(d_scal, 0, sizeof(int64_t) * (FS_COUNT + 8), stream_));
    FusedArgs A;
    A.b = b;
    A.c = c;
    A.relpc_bits = b.relpc_bits;
    A.group_t0 = b.group_t0;
    A.group_run0 = b.group_run0;
    A.scalars = d_scal;
    A.over_count = d_u32;
    A.group_nrec = d_nrec;
    A.group_desc = d_desc_off;
    A.rec_base = d_base;
    A.tkey = fused_tkey_.as<uint32_t>();
    A.adc_slots = fused_gain_.as<uint4>();
    A.desc = want ? fused_desc_.as<uint4>() : nullptr;
    A.records_out = records_out;
    A.cap_records = want ? cap_records : 0;
    // (the bookkeeping rows are also what k_groups_disjoint checks: kept in a buffer of our own if the caller wants none)
    if (!group_info_out) fused_ginfo_.reserve(sizeof(wfs_group_info) * (size_t)std::max<int64_t>(ng, 1));
    A.group_info = group_info_out ? group_info_out : fused_ginfo_.as<wfs_group_info>();
    // (rows of groups that are still to be repeated with larger lists read as "no records" in the first check)
    WFS_CUDA_CHECK(cudaMemsetAsync(A.group_info, 0, sizeof(wfs_group_info) * (size_t)ng, stream_));
    auto launch_class = [&](const ClassDef &d, const uint32_t *lst, uint32_t n, uint32_t *ticket, uint32_t *overflow_list,
                            cudaStream_t st) -> bool {
        const Layout L = make_layout(d.n_cap, d.itv_cap, d.rec_cap, n_ch, tmpl_len, c.p.dt, d.bin_bits);
        if (L.total > 227 * 1024) return false;
        static const bool small_build = !(getenv("WFS_FUSED_SMALL") && atoi(getenv("WFS_FUSED_SMALL")) == 0);
        auto *kern = (small_build && d.threads <= kFusedSmallThreads) ? k_group_analyse_small : k_group_analyse;
        if (!fused_smem_set_) {
            // the attribute belongs to the function, not to this back end: every lane sets the same (largest) value
            WFS_CUDA_CHECK(cudaFuncSetAttribute(k_group_analyse, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            WFS_CUDA_CHECK(cudaFuncSetAttribute(k_group_analyse_small, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            fused_smem_set_ = 227 * 1024;
        }
        int ctas_per_sm = 1;
        WFS_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, d.threads, L.total));
        if (ctas_per_sm < 1) return false;
        FusedClass K;
        K.n_cap = d.n_cap; K.itv_cap = d.itv_cap; K.rec_cap = d.rec_cap; K.bin_bits = d.bin_bits;
        K.list = lst; K.n_list = n; K.ticket = ticket; K.overflow_list = overflow_list;
        const int grid = (int)std::min<int64_t>(n, (int64_t)kNumSMs * ctas_per_sm);
        kern<<<grid, d.threads, L.total, st>>>(A, K);
        lc_->n++;
        return true;
    };
    auto records = [&]() {
        k_groups_disjoint<<<(unsigned)div_up(ng, 256), 256, 0, stream_>>>(ng, A.group_info, d_scal);
        lc_->n++;
        if (!want) return;
        prim_.exclusive_scan_u32(d_nrec, d_base, ng, true);
        k_group_records<<<dim3((unsigned)ng, 2), kFusedRecordThreads, 0, stream_>>>(A);
        lc_->n++;
    };
    auto read_scalars = [&]() {
        WFS_CUDA_CHECK(cudaMemcpyAsync(h_scalars_, d_scal, sizeof(int64_t) * (FS_COUNT + 1), cudaMemcpyDeviceToHost, stream_));
        WFS_CUDA_CHECK(cudaEventRecord(ev1_, stream_));
        WFS_CUDA_CHECK(stream_sync(stream_));
        WFS_CUDA_CHECK(cudaGetLastError());
    };
    WFS_CUDA_CHECK(cudaEventRecord(ev0_, stream_));
    // the classes run side by side (one stream each): the small groups fill the tail of the large ones
    WFS_CUDA_CHECK(cudaEventRecord(ev_fork_, stream_));
    for (int k = kFusedClasses - 1; k >= 0; k--) {
        if (!cls_n[k]) continue;
        cudaStream_t st = k == kFusedClasses - 1 ? stream_ : aux_[k];
        if (st != stream_) WFS_CUDA_CHECK(cudaStreamWaitEvent(st, ev_fork_, 0));
        if (!launch_class(defs[k], d_list + cls_off[k], cls_n[k], d_u32 + 1 + k, k + 1 < kFusedClasses ? d_over : nullptr, st))
            return false;
        if (st != stream_) {
            WFS_CUDA_CHECK(cudaEventRecord(ev_join_[k], st));
            WFS_CUDA_CHECK(cudaStreamWaitEvent(stream_, ev_join_[k], 0));
        }
    }
    WFS_CUDA_CHECK(cudaEventRecord(evp_[4], stream_));
    records();        // records that do not fit the buffer are skipped by the kernel: no decision on the host in between
    read_scalars();
    const uint32_t n_over = reinterpret_cast<const uint32_t *>(h_scalars_ + FS_COUNT)[0];
    if (n_over > 0 && !h_scalars_[FS_OVERFLOW] && !h_scalars_[FS_ERR]) {
        // groups that outgrew the lists of their class: once more with the largest lists (their truth counters
        // are untouched -- a group adds them only when it fits -- except the per-PMT areas), then all records again
        if (b.pmt_areas) return false;
        if (!launch_class(defs[kFusedClasses - 1], d_over, n_over, d_u32 + 1 + kFusedClasses, nullptr, stream_)) return false;
        records();
        read_scalars();
    }
    if (h_scalars_[FS_OVERFLOW]) {                           // key range / lists of the largest class: multi-pass back end
        if (getenv("WFS_DEBUG_SEG")) fprintf(stderr, "[wfs] fused back end gives up: overflow code %lld, %u groups retried\n",
                                             (long long)h_scalars_[FS_OVERFLOW], n_over);
        return false;
    }
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
    WFS_CUDA_CHECK(cudaEventElapsedTime(&res.ms_digitize, ev0_, evp_[4]));       // k_group_analyse (all classes)
    res.ms_phase[3] = res.ms_digitize;
    if (!n_over) WFS_CUDA_CHECK(cudaEventElapsedTime(&res.ms_phase[6], evp_[4], ev1_));      // scan + k_group_records
    res.segment_sorted_photons = res.segment_sorted_records = 1;
    return true;
}

}  // namespace wfs
