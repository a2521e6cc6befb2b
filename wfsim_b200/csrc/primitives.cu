// Device-wide exclusive scan and stable LSD radix sort, hand-written for sm_100a.
// These are the plumbing of the hot path: photon ordering by (group, channel, pulse call, time)
// and raw_record ordering by (time, channel) (strax.sort_by_time, strax_interface.py:453).
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>

namespace wfs {

static bool blocking_sync_mode() {
    // Spin by default.  Measured: sleeping on a blocking event costs more than the spinning waits take from the
    // other ranks -- 144 vs 119 ms per step on a 4-GPU box with 8 cores per rank (round 1), 74 vs 72 ms with 8
    // ranks on 32 cores (round 2) -- and next to anything that holds the driver's locks (an nvidia-smi query
    // every 0.2 s per rank) the sleeping waits stall outright: 154 ms per step on every rank
    // (profiles/tools/n8_device_leg.sh).  WFS_BLOCKING_SYNC=1 asks for the sleeping form.
    if (const char *e = getenv("WFS_BLOCKING_SYNC")) return atoi(e) != 0;
    return false;
}

cudaError_t stream_sync(cudaStream_t s) {
    static const bool blocking = blocking_sync_mode();
    if (!blocking) return cudaStreamSynchronize(s);
    // one blocking event per (thread, device)
    thread_local cudaEvent_t ev[16] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 16) return cudaStreamSynchronize(s);
    if (!ev[dev]) {
        e = cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    e = cudaEventRecord(ev[dev], s);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ev[dev]);
}

// ---------------------------------------------------------------------------------------------
// scan: 3-phase (block reduce, recursive scan of block sums, block scan + offset)
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanIPT = 8;
constexpr int kScanTile = kScanThreads * kScanIPT;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += u;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, total in *total.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T *smem /* >= 33 */, T *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = warp_incl_scan(v);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : T(0);
        T wi = warp_incl_scan(w);
        smem[lane] = wi - w;
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    T res = smem[warp] + incl - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const T *in, T *block_sums, int64_t n) {
    __shared__ T sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanIPT; k++) {
        int64_t i = base + k * kScanThreads + threadIdx.x;
        if (i < n) s += in[i];
    }
    T tot;
    block_excl_scan(s, sm, &tot);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const T *in, T *out, const T *block_offsets, int64_t n, int write_total) {
    __shared__ T sm[33];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanIPT;
    T v[kScanIPT];
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanIPT; k++) {
        int64_t i = base + k;
        v[k] = (i < n) ? in[i] : T(0);
        s += v[k];
    }
    T tot;
    T ex = block_excl_scan(s, sm, &tot);
    T run = ex + (block_offsets ? block_offsets[blockIdx.x] : T(0));
#pragma unroll
    for (int k = 0; k < kScanIPT; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
        if (write_total && i == n - 1) out[n] = run;
    }
}

template <typename T>
static void scan_impl(Primitives &P, const T *in, T *out, int64_t n, bool write_total, int depth,
                      size_t tmp_off) {
    if (n <= 0) {
        if (write_total) WFS_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(T), P.stream));
        return;
    }
    int nb = div_up(n, kScanTile);
    if (nb == 1) {
        k_scan_apply<T><<<1, kScanThreads, 0, P.stream>>>(in, out, nullptr, n, write_total);
        P.lc->n++;
        return;
    }
    T *bs = reinterpret_cast<T *>(reinterpret_cast<char *>(P.scan_tmp.p) + tmp_off);
    k_scan_reduce<T><<<nb, kScanThreads, 0, P.stream>>>(in, bs, n);
    P.lc->n++;
    size_t used = ((size_t)nb * sizeof(T) + 255) & ~size_t(255);
    scan_impl<T>(P, bs, bs, nb, false, depth + 1, tmp_off + used);
    k_scan_apply<T><<<nb, kScanThreads, 0, P.stream>>>(in, out, bs, n, write_total);
    P.lc->n++;
}

template <typename T>
static size_t scan_tmp_bytes(int64_t n) {
    size_t tot = 0;
    while (n > kScanTile) {
        n = div_up(n, kScanTile);
        tot += ((size_t)n * sizeof(T) + 255) & ~size_t(255);
    }
    return tot + 256;
}

void Primitives::exclusive_scan_u64(const uint64_t *in, uint64_t *out, int64_t n, bool write_total) {
    scan_tmp.reserve(scan_tmp_bytes<uint64_t>(n));
    scan_impl<uint64_t>(*this, in, out, n, write_total, 0, 0);
    WFS_CUDA_CHECK(cudaGetLastError());
}

void Primitives::exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, bool write_total) {
    scan_tmp.reserve(scan_tmp_bytes<uint32_t>(n));
    scan_impl<uint32_t>(*this, in, out, n, write_total, 0, 0);
    WFS_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// radix sort: 8-bit digits; per pass: upsweep (block digit histograms) -> scan -> downsweep
// (stable in-block ranking via warp match_any, warp-striped item order)
// ---------------------------------------------------------------------------------------------
// Lanes of the warp holding the same 8-bit digit (among the lanes of `valid`).  Eight ballots:
// match.any walks the distinct values of the warp one by one and random digits are almost all
// distinct, which made it the dominant cost of every ranking loop.
__device__ __forceinline__ unsigned match_digit8(uint32_t d, unsigned valid) {
    unsigned m = valid;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        m &= bit ? bal : ~bal;
    }
    return m;
}

constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortIPT = 12;
constexpr int kSortTile = kSortThreads * kSortIPT;

__global__ void __launch_bounds__(kSortThreads)
k_radix_upsweep(const uint64_t *__restrict__ keys, uint32_t *__restrict__ hist, int64_t n,
                int shift, int nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t chunk = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * 32 * kSortIPT;
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        int64_t i = chunk + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? (uint32_t)((keys[i] >> shift) & 255u) : 0u;
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        unsigned m = match_digit8(d, vm);
        if (valid && (m & ((1u << lane) - 1u)) == 0) atomicAdd(&h[d], __popc(m));
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// Downsweep of one pass: the tile's items are ranked by digit (stable: warp-striped item order, match-based
// ranking inside a warp, warps in order), laid out in shared memory in digit order, and written from there, so
// that consecutive threads write consecutive addresses of a digit's run (a tile of 4096 items holds 16 items per
// digit on average: 128-byte runs of keys instead of the single 8-byte writes of a direct scatter).
__global__ void __launch_bounds__(kSortThreads)
k_radix_downsweep(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                  const uint32_t *__restrict__ hist_scanned, int64_t n, int shift, int nblocks) {
    __shared__ uint64_t s_key[kSortTile];
    __shared__ uint32_t s_val[kSortTile];
    __shared__ uint32_t wcount[kSortWarps][256];
    __shared__ uint32_t dstart[256];       // first position of digit d in the tile's digit order
    __shared__ int32_t gdelta[256];        // global position of the digit's first item minus dstart
    __shared__ uint32_t s_wtot[kSortWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile0 = (int64_t)blockIdx.x * kSortTile;
    const int64_t chunk = tile0 + (int64_t)warp * 32 * kSortIPT;
    uint64_t key[kSortIPT];
    uint16_t rank[kSortIPT];
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        const int64_t i = chunk + k * 32 + lane;
        key[k] = i < n ? keys_in[i] : ~0ull;
    }
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        const int64_t i = chunk + k * 32 + lane;
        const bool valid = i < n;
        const uint32_t d = (uint32_t)((key[k] >> shift) & 255u);
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        const unsigned m = match_digit8(d, vm);
        const uint32_t r = __popc(m & ((1u << lane) - 1u));
        const uint32_t old = valid ? wcount[warp][d] : 0u;
        __syncwarp();
        if (valid && r == 0) wcount[warp][d] = old + __popc(m);
        __syncwarp();
        rank[k] = (uint16_t)(old + r);
    }
    __syncthreads();
    {
        // per digit: exclusive offsets of the warps, the digit's count; then an exclusive scan over the digits
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            const uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t base = 0;
        for (int w = 0; w < warp; w++) base += s_wtot[w];
        const uint32_t start = base + inc - run;
        dstart[d] = start;
        gdelta[d] = (int32_t)(hist_scanned[(int64_t)d * nblocks + blockIdx.x] - start);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        const int64_t i = chunk + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)((key[k] >> shift) & 255u);
            const uint32_t pos = dstart[d] + wcount[warp][d] + rank[k];
            s_key[pos] = key[k];
            s_val[pos] = vals_in[i];          // (read here, not kept in registers over the ranking)
        }
    }
    __syncthreads();
    const int n_here = (int)min((int64_t)kSortTile, n - tile0);
    for (int pos = threadIdx.x; pos < n_here; pos += kSortThreads) {
        const uint64_t kk = s_key[pos];
        const uint32_t d = (uint32_t)((kk >> shift) & 255u);
        const uint32_t dst = (uint32_t)((int32_t)pos + gdelta[d]);
        keys_out[dst] = kk;
        vals_out[dst] = s_val[pos];
    }
}

void Primitives::sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int key_bits) {
    if (n <= 1 || key_bits <= 0) return;
    if (n >= (int64_t(1) << 32)) throw std::runtime_error("sort_pairs: n too large for one batch");
    int passes = (key_bits + 7) / 8;
    int nblocks = div_up(n, kSortTile);
    sort_keys_alt.reserve((size_t)n * sizeof(uint64_t));
    sort_vals_alt.reserve((size_t)n * sizeof(uint32_t));
    sort_hist.reserve(((size_t)256 * nblocks + 1) * sizeof(uint32_t));
    uint64_t *k_in = keys, *k_out = sort_keys_alt.as<uint64_t>();
    uint32_t *v_in = vals, *v_out = sort_vals_alt.as<uint32_t>();
    uint32_t *hist = sort_hist.as<uint32_t>();
    for (int p = 0; p < passes; p++) {
        int shift = p * 8;
        k_radix_upsweep<<<nblocks, kSortThreads, 0, stream>>>(k_in, hist, n, shift, nblocks);
        lc->n++;
        exclusive_scan_u32(hist, hist, (int64_t)256 * nblocks, false);
        k_radix_downsweep<<<nblocks, kSortThreads, 0, stream>>>(k_in, v_in, k_out, v_out, hist, n,
                                                               shift, nblocks);
        lc->n++;
        std::swap(k_in, k_out);
        std::swap(v_in, v_out);
    }
    if (k_in != keys) {
        WFS_CUDA_CHECK(cudaMemcpyAsync(keys, k_in, (size_t)n * sizeof(uint64_t),
                                       cudaMemcpyDeviceToDevice, stream));
        WFS_CUDA_CHECK(cudaMemcpyAsync(vals, v_in, (size_t)n * sizeof(uint32_t),
                                       cudaMemcpyDeviceToDevice, stream));
    }
    WFS_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// segmented sort: one CTA per segment, LSD radix on 8-bit digits entirely in shared memory.
// Items are packed as (dropped | key bits | index in segment) in one 64-bit word; digits on which
// all items of the segment agree are skipped (high time bits, pulse-call rank, ... are often constant).
// ---------------------------------------------------------------------------------------------
constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;
constexpr int kSegIdxBits = 13;
static_assert((1 << kSegIdxBits) >= kSegSortMax, "segment index bits");

// Processes the segments with n_lo < n <= cap (the host launches one size class after the other, so
// that the many short segments run at high occupancy).  Shared memory: [2][cap] packed items + [cap] ranks.
template <int kRoundsMax>
__global__ void __launch_bounds__(kSegThreads)
k_segment_sort(int n_seg, const uint32_t *__restrict__ seg_start, const uint32_t *__restrict__ out_start,
               const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
               uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int key_bits,
               uint64_t drop_from, int n_lo, int cap, int n_ranges) {
    extern __shared__ uint64_t s_buf[];                 // [2][cap] items
    __shared__ uint32_t wcount[kSegWarps][256];
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t s_wsum[kSegWarps];
    __shared__ unsigned long long s_diff;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const uint64_t kmask = (uint64_t(1) << key_bits) - 1u;
    for (int seg = blockIdx.x; seg < n_seg; seg += gridDim.x) {
        // a segment is the concatenation of up to four ranges of the input (seg_start is [n_ranges][n_seg + 1])
        uint32_t ra[4] = {0, 0, 0, 0}, rc[4] = {0, 0, 0, 0};
        int n = 0;
        for (int r = 0; r < n_ranges; r++) {
            ra[r] = seg_start[(size_t)r * (n_seg + 1) + seg];
            rc[r] = seg_start[(size_t)r * (n_seg + 1) + seg + 1] - ra[r];
            n += (int)rc[r];
        }
        if (n <= n_lo || n > cap) continue;
        auto src_of = [&](uint32_t i) -> uint32_t {
            if (i < rc[0]) return ra[0] + i;
            i -= rc[0];
            if (i < rc[1]) return ra[1] + i;
            i -= rc[1];
            if (i < rc[2]) return ra[2] + i;
            return ra[3] + (i - rc[2]);
        };
        const uint32_t a = src_of(0);
        const uint32_t o = out_start ? out_start[seg] : a;
        uint64_t *cur = s_buf, *alt = s_buf + cap;
        __syncthreads();
        if (tid == 0) s_diff = 0ull;
        __syncthreads();
        {
            uint64_t diff = 0, x0 = 0;
            {
                const uint64_t k = keys_in[a];
                x0 = ((k >= drop_from ? (uint64_t(1) << key_bits) : 0ull) | (k & kmask)) << kSegIdxBits;
            }
            for (int i = tid; i < n; i += kSegThreads) {
                const uint64_t k = keys_in[src_of((uint32_t)i)];
                const uint64_t x = (((k >= drop_from ? (uint64_t(1) << key_bits) : 0ull) | (k & kmask)) << kSegIdxBits) |
                                   (uint64_t)i;
                cur[i] = x;
                diff |= (x ^ x0) >> kSegIdxBits;
            }
            const uint32_t dlo = __reduce_or_sync(0xffffffffu, (uint32_t)diff);
            const uint32_t dhi = __reduce_or_sync(0xffffffffu, (uint32_t)(diff >> 32));
            if (lane == 0 && (dlo | dhi)) atomicOr(&s_diff, ((unsigned long long)dhi << 32) | dlo);
        }
        __syncthreads();
        const uint64_t diffbits = (uint64_t)s_diff << kSegIdxBits;
        // warp w ranks the contiguous chunk [w * chunk, (w + 1) * chunk) in rounds of 32 items
        const int rounds = (n + kSegThreads - 1) / kSegThreads;
        const int chunk = rounds * 32;
        for (int shift = kSegIdxBits; shift < kSegIdxBits + key_bits + 1; shift += 8) {
            if (((diffbits >> shift) & 255u) == 0) continue;     // every item has the same digit
            for (int i = tid; i < kSegWarps * 256; i += kSegThreads) (&wcount[0][0])[i] = 0;
            __syncthreads();
            // ranks stay in registers; the loop is unrolled so that the shared-memory loads and the
            // ballots of several rounds are in flight while the per-digit counters are updated in order
            uint64_t xs[kRoundsMax];
            uint16_t rank[kRoundsMax];
#pragma unroll
            for (int k = 0; k < kRoundsMax; k++) {
                if (k < rounds) {
                    const int i = warp * chunk + k * 32 + lane;
                    const bool valid = i < n;
                    const uint64_t x = valid ? cur[i] : ~0ull;
                    xs[k] = x;
                    const uint32_t d = (uint32_t)(x >> shift) & 255u;
                    const unsigned vm = __ballot_sync(0xffffffffu, valid);
                    const unsigned m = __match_any_sync(0xffffffffu, d) & vm;
                    const uint32_t r = __popc(m & lt);
                    const uint32_t old = valid ? wcount[warp][d] : 0u;
                    __syncwarp();
                    if (valid && r == 0) wcount[warp][d] = old + __popc(m);
                    __syncwarp();
                    rank[k] = (uint16_t)(old + r);
                }
            }
            __syncthreads();
            {   // digit tid: exclusive prefix over the warps, then over the digits
                uint32_t run = 0;
#pragma unroll
                for (int w = 0; w < kSegWarps; w++) {
                    const uint32_t c = wcount[w][tid];
                    wcount[w][tid] = run;
                    run += c;
                }
                uint32_t inc = run;
#pragma unroll
                for (int s = 1; s < 32; s <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, s);
                    if (lane >= s) inc += u;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                uint32_t wbase = 0;
                for (int w = 0; w < warp; w++) wbase += s_wsum[w];
                dbase[tid] = wbase + inc - run;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < kRoundsMax; k++) {
                if (k < rounds) {
                    const int i = warp * chunk + k * 32 + lane;
                    if (i < n) {
                        const uint64_t x = xs[k];
                        const uint32_t d = (uint32_t)(x >> shift) & 255u;
                        alt[dbase[d] + wcount[warp][d] + rank[k]] = x;
                    }
                }
            }
            __syncthreads();
            uint64_t *t = cur; cur = alt; alt = t;
        }
        for (int i = tid; i < n; i += kSegThreads) {
            const uint64_t x = cur[i];
            if ((x >> (kSegIdxBits + key_bits)) & 1ull) continue;     // dropped (they sort behind the kept ones)
            const uint32_t src = src_of((uint32_t)(x & ((1u << kSegIdxBits) - 1u)));
            keys_out[o + i] = keys_in[src];
            vals_out[o + i] = vals_in[src];
        }
    }
}

void Primitives::segment_sort_pairs(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                                    uint32_t *vals_out, const uint32_t *seg_start, const uint32_t *out_start,
                                    int64_t n_seg, int64_t max_seg, int key_bits, uint64_t drop_from, int n_ranges) {
    if (n_seg <= 0) return;
    if (max_seg > kSegSortMax || key_bits + 1 + kSegIdxBits > 64 || n_ranges < 1 || n_ranges > 4)
        throw std::runtime_error("segment_sort_pairs: segment too large / key too wide");
    auto smem_of = [](int cap) { return (size_t)cap * 2 * sizeof(uint64_t); };
    if (!seg_attr_set) {
        WFS_CUDA_CHECK(cudaFuncSetAttribute(k_segment_sort<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem_of(kSegSortMax)));
        seg_attr_set = true;
    }
    // optionally two size classes (WFS_SEG_SPLIT): measured on B200 (C1 groups, 30-4800 photons) one
    // launch over all sizes is as fast as any split -- the kernel is bound by the dependent
    // rank/count chains of each pass, not by occupancy
    const int split = getenv("WFS_SEG_SPLIT") ? atoi(getenv("WFS_SEG_SPLIT")) : 0;
    const int cap_all = (int)((max_seg + 255) / 256 * 256);
    auto launch = [&](int lo, int cap) {
        const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)200 * 1024 / (smem_of(cap) + 10 * 1024)));
        const int grid = (int)std::min<int64_t>(n_seg, (int64_t)kNumSMs * per_sm);
        if (cap <= 8 * kSegThreads)
            k_segment_sort<8><<<grid, kSegThreads, smem_of(cap), stream>>>((int)n_seg, seg_start, out_start, keys_in,
                                                                         vals_in, keys_out, vals_out, key_bits,
                                                                         drop_from, lo, cap, n_ranges);
        else
            k_segment_sort<32><<<grid, kSegThreads, smem_of(cap), stream>>>((int)n_seg, seg_start, out_start, keys_in,
                                                                          vals_in, keys_out, vals_out, key_bits,
                                                                          drop_from, lo, cap, n_ranges);
        lc->n++;
    };
    if (split >= 256 && split <= 8 * kSegThreads && cap_all > split) {
        launch(0, split);
        launch(split, cap_all);
    } else {
        launch(0, cap_all);
    }
    WFS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace wfs
