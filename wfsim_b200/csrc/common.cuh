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

// ---------------------------------------------------------------------------------------------
// Device-wide primitives (implemented in primitives.cu)
// ---------------------------------------------------------------------------------------------
struct Primitives {
    cudaStream_t stream = nullptr;
    LaunchCounter *lc = nullptr;
    DevBuf scan_tmp, sort_hist, sort_keys_alt, sort_vals_alt, red_tmp;

    // out[i] = sum_{j<i} in[j] (exclusive); out may alias in.  Returns nothing; total is
    // written to out[n] when `write_total` (out must then hold n+1 entries).
    void exclusive_scan_u64(const uint64_t *in, uint64_t *out, int64_t n, bool write_total);
    void exclusive_scan_u32(const uint32_t *in, uint32_t *out, int64_t n, bool write_total);

    // Stable LSD radix sort of (key, value) pairs on bits [0, key_bits).  Sorted result ends in
    // keys/vals (ping-pong handled inside).
    void sort_pairs(uint64_t *keys, uint32_t *vals, int64_t n, int key_bits);

    void release() {
        scan_tmp.release(); sort_hist.release(); sort_keys_alt.release();
        sort_vals_alt.release(); red_tmp.release();
    }
};

}  // namespace wfs
