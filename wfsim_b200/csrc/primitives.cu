// Device-wide exclusive scan and stable LSD radix sort, hand-written for sm_100a.
// These are the plumbing of the hot path: photon ordering by (group, channel, pulse call, time)
// and raw_record ordering by (time, channel) (strax.sort_by_time, strax_interface.py:453).
#include "common.cuh"

namespace wfs {

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
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortIPT = 8;
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
        unsigned m = __match_any_sync(0xffffffffu, d) & vm;
        if (valid && (m & ((1u << lane) - 1u)) == 0) atomicAdd(&h[d], __popc(m));
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kSortThreads)
k_radix_downsweep(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                  uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                  const uint32_t *__restrict__ hist_scanned, int64_t n, int shift, int nblocks) {
    __shared__ uint32_t wcount[kSortWarps][256];
    __shared__ uint32_t gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    int64_t chunk = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * 32 * kSortIPT;
    uint64_t key[kSortIPT];
    uint32_t val[kSortIPT];
    uint32_t rank[kSortIPT];
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        int64_t i = chunk + k * 32 + lane;
        bool valid = i < n;
        key[k] = valid ? keys_in[i] : ~0ull;
        val[k] = valid ? vals_in[i] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        int64_t i = chunk + k * 32 + lane;
        bool valid = i < n;
        uint32_t d = (uint32_t)((key[k] >> shift) & 255u);
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        unsigned m = __match_any_sync(0xffffffffu, d) & vm;
        uint32_t r = __popc(m & ((1u << lane) - 1u));
        uint32_t old = valid ? wcount[warp][d] : 0u;
        __syncwarp();
        if (valid && r == 0) wcount[warp][d] = old + __popc(m);
        __syncwarp();
        rank[k] = old + r;
    }
    __syncthreads();
    {
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kSortWarps; w++) {
            uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
        gbase[d] = hist_scanned[(int64_t)d * nblocks + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortIPT; k++) {
        int64_t i = chunk + k * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)((key[k] >> shift) & 255u);
            uint32_t dst = gbase[d] + wcount[warp][d] + rank[k];
            keys_out[dst] = key[k];
            vals_out[dst] = val[k];
        }
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

}  // namespace wfs
