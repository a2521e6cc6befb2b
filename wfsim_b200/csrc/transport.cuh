// Record transport from HBM to the caller's host array.
//
// raw_records are 244 bytes each, most of them the constant baseline (and zero padding behind
// `length`): copying them as they are makes the PCIe link the bottleneck of the whole path
// (30 GB per 1e5 low-energy events).  The pack kernel can therefore emit a COMPACT form instead:
//   CompactHdr[n_rec]   24 B: the header fields of strax_interface.py:425-436 that are neither constant
//                       nor derivable, a 28-bit mask of the 4-sample blocks that differ from the fill
//                       pattern (baseline for samples < length, 0 behind it) and the offset of the first;
//   blocks[n_blocks]    8 B per differing block (block 27 holds 2 samples + 2 pad).
// Both streams are copied to pinned staging memory and a pool of host threads expands them into the
// caller's array (header, fill pattern, patched blocks) with streaming stores, so the destination
// needs neither pinning nor alignment.  Records are bit-identical to what k_pack<false> writes.
#pragma once
#include "common.cuh"
#include "compact_format.h"

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

namespace wfs {

struct CompactOut {      // device pointers handed to Backend::run
    CompactHdr *hdr = nullptr;
    uint2 *blocks = nullptr;
};

// Expands records [j0, j1) of a compact batch into dst (dst = address of record 0 of the batch);
// dispatches to the widest expander the CPU supports (WFS_EXPAND_ISA=sse2|avx2|avx512 overrides).
void expand_records(const CompactHdr *hdr, const uint8_t *blocks, int64_t j0, int64_t j1,
                    uint8_t *dst, int16_t fill, int16_t dt);

class HostPool;

struct TransportStats {          // wall clock, summed over the batches of one call
    std::atomic<int64_t> ns_copy{0};      // ship -> copies complete (waits for the stream included)
    std::atomic<int64_t> ns_expand{0};    // copies complete -> last slice expanded
    std::atomic<int64_t> n_plain{0};      // records that travelled as plain 244-byte rows (split transport)
};

// Split transport (page-locked destinations only): the leading rows of a batch leave HBM as they are (DMA
// straight into the caller's array, no host core involved), the rest as the compact form that the pool expands.
// The fraction follows what the host can do: each batch reports when its DMA and when its expansion finished,
// and the share moves towards the side that finished first.  The records do not depend on it.
struct SplitControl {
    std::atomic<int> permille{-1};        // share of plain rows; -1: not initialised
    bool fixed = false;                   // WFS_PLAIN_FRACTION given
    bool frozen = false;                  // enough expansion threads: everything compact, no feedback
    void init(int expand_threads);
    double fraction() const { return std::max(0, permille.load()) * 1e-3; }
    // t_*: steady-clock ns; expansion / DMA of one batch finished
    void feedback(int64_t t_ship, int64_t t_plain_done, int64_t t_expand_done);
};

// One batch in flight: filled in by the producer, enqueued (from a stream callback) when its D2H
// copies have completed, executed in slices by the pool.
struct ExpandJob {
    HostPool *pool = nullptr;
    const CompactHdr *hdr = nullptr;
    const uint8_t *blocks = nullptr;
    int64_t n_rec = 0;
    uint8_t *dst = nullptr;
    int16_t fill = 0, dt = 0;
    int slices = 0;
    TransportStats *stats = nullptr;
    int64_t t_ship = 0, t_callback = 0;
    // split transport: the plain rows of the same batch travel behind the compact streams
    SplitControl *split = nullptr;
    std::atomic<int64_t> t_plain_done{0}, t_expand_done{0};
    std::atomic<int> arrivals{0};
    void arrive();               // DMA done / expansion done: the second one reports to the controller
    static void CUDART_CB plain_done_callback(void *job);
    // completion
    std::mutex mu;
    std::condition_variable cv;
    int remaining = 0;
    bool pending = false;

    void arm(int n_slices);      // before the stream callback is queued
    void wait();                 // until every slice has run (no-op if nothing is pending)
    void abandon();              // the stream failed before the callback could run: nothing will come
};

class HostPool {
public:
    explicit HostPool(int n_threads);
    ~HostPool();
    int size() const { return (int)threads_.size(); }
    void enqueue(ExpandJob *job);            // callable from a CUDA host function
    static void CUDART_CB stream_callback(void *job);   // cudaLaunchHostFunc target
    static int default_threads();

private:
    struct Task { ExpandJob *job; int k; };
    void worker();
    std::vector<std::thread> threads_;
    std::deque<Task> queue_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// Grow-only pinned host buffer.
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes) {
        if (bytes <= cap) return;
        if (p) WFS_CUDA_CHECK(cudaFreeHost(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        WFS_CUDA_CHECK(cudaHostAlloc(&p, want, cudaHostAllocDefault));
        cap = want;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

// Device buffers + pinned staging + job of one compact batch in flight.
struct CompactStage {
    DevBuf d_hdr, d_blk;
    PinnedBuf h_hdr, h_blk;
    ExpandJob job;
    int64_t cap_records() const {
        return (int64_t)std::min(d_hdr.cap / sizeof(CompactHdr), d_blk.cap / ((size_t)kBlockBytes * kBlocksPerRecord));
    }
    void reserve_device(int64_t n_rec) {
        d_hdr.reserve(sizeof(CompactHdr) * (size_t)n_rec);
        d_blk.reserve((size_t)kBlockBytes * kBlocksPerRecord * (size_t)n_rec);
    }
    CompactOut out() { return CompactOut{d_hdr.as<CompactHdr>(), d_blk.as<uint2>()}; }
    // Queues the D2H copies of (n_rec headers, n_blocks blocks) on `copy_stream` and, behind them,
    // the expansion into dst.  The caller must job.wait() before touching the stage again.
    // `split`: plain rows of the same batch follow on the stream; the caller queues them and then
    // ExpandJob::plain_done_callback(&job).
    void ship(HostPool *pool, cudaStream_t copy_stream, int64_t n_rec, int64_t n_blocks, uint8_t *dst,
              int16_t fill, int16_t dt, TransportStats *stats = nullptr, SplitControl *split = nullptr);
    void release() {
        d_hdr.release(); d_blk.release(); h_hdr.release(); h_blk.release();
    }
};

}  // namespace wfs
