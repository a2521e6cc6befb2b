// Host side of the compact record transport (see transport.cuh): pinned staging, the expansion
// thread pool and the record expander.  The reference builds each record with numpy slicing
// (strax_interface.py:425-436); here the same 244 bytes are produced from the compact form.
#include "transport.cuh"
#include "../../include/wfsim_b200.h"

#include <algorithm>
#include <chrono>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

namespace wfs {

static_assert(WFS_RECORD_BYTES == 244 && WFS_SAMPLES_PER_RECORD == 110, "record layout");

typedef void (*ExpandFn)(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);

static ExpandFn pick_expander() {
    const char *e = getenv("WFS_EXPAND_ISA");
    const std::string want = e ? e : "";
    __builtin_cpu_init();
    const bool avx2 = __builtin_cpu_supports("avx2");
    const bool avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
                        __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    if (want == "sse2") return expand_records_sse2;
    if (want == "avx2" && avx2) return expand_records_avx2;
    if (want == "avx512" && avx512) return expand_records_avx512;
    return avx512 ? expand_records_avx512 : avx2 ? expand_records_avx2 : expand_records_sse2;
}

void expand_records(const CompactHdr *hdr, const uint8_t *blocks, int64_t j0, int64_t j1,
                    uint8_t *dst_base, int16_t fill, int16_t dt) {
    static const ExpandFn fn = pick_expander();
    fn(hdr, blocks, j0, j1, dst_base, fill, dt);
}

// ---------------------------------------------------------------------------------------------
void ExpandJob::arm(int n_slices) {
    std::lock_guard<std::mutex> lk(mu);
    slices = n_slices;
    remaining = n_slices;
    pending = true;
}

void ExpandJob::wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !pending; });
}

void ExpandJob::abandon() {
    {
        std::lock_guard<std::mutex> lk(mu);
        pending = false;
    }
    cv.notify_all();
}

void SplitControl::init(int expand_threads) {
    if (const char *e = getenv("WFS_PLAIN_FRACTION")) {      // fixed share (tests, measurements); read at every call
        fixed = true;
        permille = (int)std::max(0.0, std::min(1000.0, atof(e) * 1000.0));
        return;
    }
    if (fixed) { fixed = false; permille = -1; }
    if (permille.load() >= 0) return;
    frozen = false;
    // Measured on the B200 boxes of this pool (profiles/r2w_split_transport.md): one expansion thread writes
    // ~9.8 GB/s of records, eight or more reach what the host memory takes (~80 GB/s) and plain rows on top of
    // that gain nothing; the link moves P = 47 GB/s; compact / plain bytes c = 0.28 for single-photon records.
    // Few threads (several ranks sharing the cores of one box): both sides finish together at
    // f = (1/E - c/P) / ((1 - c)/P + 1/E), then the feedback below takes over.
    // ... and measured with 8 ranks on one 32-core box (profiles/r3f_n8_*.json): 8 x 0.7 x 30 GB of DMA rows per step
    // into one host memory system took 4.7 s per step where the compact form alone takes 2.6 s -- the cores are not
    // the bound there either, the memory system is, and it takes DMA writes from eight devices worse than the
    // expanders' streaming stores.  So the split is opt-in: WFS_PLAIN_ADAPTIVE=1 (or a fixed WFS_PLAIN_FRACTION).
    const char *adaptive = getenv("WFS_PLAIN_ADAPTIVE");
    if (expand_threads >= 8 || !(adaptive && atoi(adaptive) != 0)) {
        frozen = true;
        permille = 0;
        return;
    }
    const double E = 9.8 * std::max(1, expand_threads), P = 47.0, c = 0.28;
    const double f = (1.0 / E - c / P) / ((1.0 - c) / P + 1.0 / E);
    permille = (int)std::max(0.0, std::min(950.0, f * 1000.0));
}

void SplitControl::feedback(int64_t t_ship, int64_t t_plain_done, int64_t t_expand_done) {
    if (fixed || frozen) return;
    const double dp = (double)(t_plain_done - t_ship), de = (double)(t_expand_done - t_ship);
    const double span = std::max(std::max(dp, de), 1.0);
    const double err = (de - dp) / span;                  // > 0: the expansion finished last -> more plain rows
    int cur = permille.load(), next;
    do {
        next = (int)std::max(0.0, std::min(950.0, cur + 120.0 * err));
    } while (!permille.compare_exchange_weak(cur, next));
}

void ExpandJob::arrive() {
    if (arrivals.fetch_add(1) == 1 && split) split->feedback(t_ship, t_plain_done.load(), t_expand_done.load());
}

void CUDART_CB ExpandJob::plain_done_callback(void *p) {
    ExpandJob *job = reinterpret_cast<ExpandJob *>(p);
    job->t_plain_done = std::chrono::duration_cast<std::chrono::nanoseconds>(
                            std::chrono::steady_clock::now().time_since_epoch()).count();
    job->arrive();
}

int host_cores_per_rank() {
    // the cores this process may run on (cgroup / taskset aware, unlike hardware_concurrency), shared
    // with the LOCAL_WORLD_SIZE - 1 other per-GPU processes of the box (torchrun sets it)
    cpu_set_t set;
    CPU_ZERO(&set);
    int hw = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : 0;
    if (hw <= 0) hw = (int)std::thread::hardware_concurrency();
    if (const char *w = getenv("LOCAL_WORLD_SIZE")) hw /= std::max(1, atoi(w));
    return std::max(1, hw);
}

int HostPool::default_threads() {
    if (const char *e = getenv("WFS_EXPAND_THREADS")) return std::max(1, atoi(e));
    return std::max(1, std::min(16, host_cores_per_rank() - 2));
}

HostPool::HostPool(int n_threads) {
    for (int i = 0; i < std::max(1, n_threads); i++) threads_.emplace_back([this] { worker(); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
}

void HostPool::enqueue(ExpandJob *job) {
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (int k = 0; k < job->slices; k++) queue_.push_back(Task{job, k});
    }
    cv_.notify_all();
}

static int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void CUDART_CB HostPool::stream_callback(void *p) {
    ExpandJob *job = reinterpret_cast<ExpandJob *>(p);
    job->t_callback = now_ns();
    job->pool->enqueue(job);
}

void HostPool::worker() {
    for (;;) {
        Task t;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || !queue_.empty(); });
            if (queue_.empty()) return;     // stop requested and nothing left
            t = queue_.front();
            queue_.pop_front();
        }
        ExpandJob *job = t.job;
        const int64_t per = (job->n_rec + job->slices - 1) / job->slices;
        const int64_t j0 = std::min(job->n_rec, per * t.k), j1 = std::min(job->n_rec, per * (t.k + 1));
        if (j1 > j0) expand_records(job->hdr, job->blocks, j0, j1, job->dst, job->fill, job->dt);
        bool last;
        {
            std::lock_guard<std::mutex> lk(job->mu);
            last = --job->remaining == 0;
            if (last) {
                const int64_t t_done = now_ns();
                if (job->stats) {
                    job->stats->ns_copy += job->t_callback - job->t_ship;
                    job->stats->ns_expand += t_done - job->t_callback;
                }
                if (job->split) {
                    job->t_expand_done = t_done;
                    job->arrive();
                }
                job->pending = false;
            }
        }
        if (last) job->cv.notify_all();
    }
}

void CompactStage::ship(HostPool *pool, cudaStream_t copy_stream, int64_t n_rec, int64_t n_blocks,
                        uint8_t *dst, int16_t fill, int16_t dt, TransportStats *stats, SplitControl *split) {
    const size_t hdr_bytes = sizeof(CompactHdr) * (size_t)n_rec, blk_bytes = (size_t)kBlockBytes * (size_t)n_blocks;
    job.wait();
    h_hdr.reserve(hdr_bytes);
    h_blk.reserve(blk_bytes + 64);
    // in pieces: the copy engine serves streams in FIFO order, and small count readbacks of other
    // lanes must not queue behind hundreds of megabytes
    const size_t piece = size_t(8) << 20;
    for (size_t o = 0; o < hdr_bytes; o += piece)
        WFS_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)h_hdr.p + o, d_hdr.as<uint8_t>() + o, std::min(piece, hdr_bytes - o),
                                       cudaMemcpyDeviceToHost, copy_stream));
    for (size_t o = 0; o < blk_bytes; o += piece)
        WFS_CUDA_CHECK(cudaMemcpyAsync((uint8_t *)h_blk.p + o, d_blk.as<uint8_t>() + o, std::min(piece, blk_bytes - o),
                                       cudaMemcpyDeviceToHost, copy_stream));
    job.pool = pool;
    job.stats = stats;
    job.split = split;
    job.arrivals = 0;
    job.t_ship = now_ns();
    job.hdr = reinterpret_cast<const CompactHdr *>(h_hdr.p);
    job.blocks = reinterpret_cast<const uint8_t *>(h_blk.p);
    job.n_rec = n_rec;
    job.dst = dst;
    job.fill = fill;
    job.dt = dt;
    // slices of at least 16k records: small batches are not worth a wake-up per thread
    job.arm((int)std::max<int64_t>(1, std::min<int64_t>(4 * pool->size(), n_rec / 16384)));
    if (cudaLaunchHostFunc(copy_stream, HostPool::stream_callback, &job) != cudaSuccess) {
        {
            std::lock_guard<std::mutex> lk(job.mu);
            job.pending = false;
        }
        throw std::runtime_error("CUDA error: cudaLaunchHostFunc failed");
    }
}

}  // namespace wfs
