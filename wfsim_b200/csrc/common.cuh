// Shared device/host utilities for the wfsim_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <stdexcept>
#include <atomic>

#define WFS_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            char _buf[512];                                                               \
            snprintf(_buf, sizeof(_buf), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), \
                     __FILE__, __LINE__, cudaGetErrorString(_e));                         \
            throw std::runtime_error(_buf);                                               \
        }                                                                                 \
    } while (0)

namespace wfs {

constexpr int kNumSMs = 148;  // B200

// Grow-only device buffer: the workspace of a handle is a set of these, so steady-state calls
// perform no cudaMalloc.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) WFS_CUDA_CHECK(cudaFree(p));
        p = nullptr;
        size_t want = bytes + bytes / 4 + 256;
        WFS_CUDA_CHECK(cudaMalloc(&p, want));
        cap = want;
    }
    // grow, keeping the first `keep_bytes` of the old content
    void reserve_keep(size_t bytes, size_t keep_bytes, cudaStream_t stream) {
        if (bytes <= cap) return;
        void *old = p;
        size_t want = bytes + bytes / 2 + 256;
        void *np_ = nullptr;
        WFS_CUDA_CHECK(cudaMalloc(&np_, want));
        if (old && keep_bytes) {
            WFS_CUDA_CHECK(cudaMemcpyAsync(np_, old, keep_bytes, cudaMemcpyDeviceToDevice, stream));
            WFS_CUDA_CHECK(cudaStreamSynchronize(stream));
        }
        if (old) WFS_CUDA_CHECK(cudaFree(old));
        p = np_;
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct LaunchCounter {
    std::atomic<int64_t> n{0};   // lanes launch from their own host threads
};

static inline int div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Wait for the work queued on a stream: cudaStreamSynchronize (spinning) while every lane thread of
// every rank has a core of its own; a sleep on a blocking event when the lane threads would take more
// than half of this rank's cores (WFS_BLOCKING_SYNC=0/1 overrides).
cudaError_t stream_sync(cudaStream_t s);
int host_cores_per_rank();   // cores of the affinity mask / LOCAL_WORLD_SIZE (transport.cu)
int default_lanes();         // host threads driving device batches side by side (frontend.cu)

constexpr int kSegSortMax = 8192;   // items per segment of Primitives::segment_sort_pairs

// ---------------------------------------------------------------------------------------------
// Device-wide primitives (implemented in primitives.cu)
// ---------------------------------------------------------------------------------------------
struct Primitives {
    cudaStream_t stream = nullptr;
    LaunchCounter *lc = nullptr;
    DevBuf scan_tmp, sort_hist, sort_keys_alt, sort_vals_alt, red_tmp;
    bool seg_attr_set = false;

    // out[i] = sum_{j<i} in[j] (exclusive); out may alias in.  Returns nothing; total is
    // written to out[n] when `write_total` (out must then hold n+1 entries).
    void exclusive_scan_u64(const uint64_t *in, uint64_t *out, int64_t n, bool write_total);
    void exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, bool write_total);

    // Stable LSD radix sort of (key, value) pairs on bits [0, key_bits).  Sorted result ends in
    // keys/vals (ping-pong handled inside).
    void sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int key_bits);

    // Stable sort of (key, value) pairs INSIDE segments: segment s holds items
    // [seg_start[s], seg_start[s+1]) of keys_in/vals_in (device arrays) -- or, with n_ranges > 1, the
    // concatenation of that range of each of the n_ranges tables seg_start[r][n_seg + 1] (at most
    // four) -- each at most kSegSortMax items (max_seg = the largest one, sizes the shared memory).  Ordered by bits [0, key_bits) of
    // the key -- the bits above must be equal inside a segment; items with key >= drop_from are
    // dropped.  Segment s is written to keys_out/vals_out at out_start[s] (or seg_start[s] when
    // out_start is null).  One CTA per segment, everything in shared memory: one read and one
    // write of the data instead of one per digit.
    void segment_sort_pairs(const uint64_t *keys_in, const uint32_t *vals_in, uint64_t *keys_out,
                            uint32_t *vals_out, const uint32_t *seg_start, const uint32_t *out_start,
                            int64_t n_seg, int64_t max_seg, int key_bits, uint64_t drop_from, int n_ranges = 1);

    void release() {
        scan_tmp.release(); sort_hist.release(); sort_keys_alt.release();
        sort_vals_alt.release(); red_tmp.release();
    }
};

}  // namespace wfs
