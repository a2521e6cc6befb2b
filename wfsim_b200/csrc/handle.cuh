// Per-GPU simulator handle: device tables, streams, workspaces, back end + front end.
#pragma once
#include "backend.cuh"
#include <mutex>

namespace wfs {

struct Frontend;   // sampling stages + host scheduler (frontend.cu)
struct Lane;       // stream + workspaces + host thread processing every n-th device batch (frontend.cu)

struct Handle {
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_a, ev_b, ev_c, ev_d;
    DeviceConfig cfg;
    std::vector<void *> owned;          // device tables freed at destroy
    std::vector<double> h_gains;
    LaunchCounter launches;
    Backend *backend = nullptr;
    Frontend *frontend = nullptr;
    std::vector<Lane *> lanes;          // lane 0 = (stream, backend, frontend) above
    std::string last_error;
    void *staged_plan = nullptr;        // Plan of wfs_stage_instructions (frontend.cu)
    // staging for host-pointer calls
    DevBuf d_t, d_ch, d_gain, d_pc, d_pc_group, d_pc_rank, d_ix, d_records, d_groups;
    DevBuf d_opt_ch, d_opt_t;           // externally supplied photons of the current wfs_simulate call
    int64_t opt_cutoff = 0;
    // compact record transport (transport.cuh): expansion threads + the stage of wfs_simulate_photons
    HostPool *pool = nullptr;
    CompactStage cstage;
    TransportStats tstats;
    SplitControl split;                 // share of the records that travels as plain rows (page-locked destinations)
    static bool page_locked(const void *p) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return attr.type == cudaMemoryTypeHost;
    }
    int compact_mode = 1;               // WFS_COMPACT: 0 never, 1 (default) see use_compact, 2 always

    // Compact transport + host expansion, or a plain DMA of the 244-byte rows?  With the noise on
    // nothing is compressible (the compact form is 248 B per record), but a plain copy is only fast
    // into page-locked memory: into an ordinary numpy array the driver stages it at a few GB/s, while
    // the expansion threads write it at the host's memory bandwidth.
    bool use_compact(const void *dst) const {
        if (compact_mode != 1) return compact_mode == 2;
        if (!(cfg.p.enable_noise && cfg.noise_t)) return true;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, dst) != cudaSuccess) { cudaGetLastError(); return true; }
        return attr.type != cudaMemoryTypeHost;      // pinned destination: plain DMA straight into it
    }
    std::once_flag pool_once;
    HostPool *host_pool() {            // lanes call this from their own host threads
        std::call_once(pool_once, [this] { pool = new HostPool(HostPool::default_threads()); });
        return pool;
    }
    int16_t record_fill() const { return (int16_t)std::max(cfg.p.baseline, 0); }

    Handle(const wfs_params &p, const wfs_tables &t, int dev);
    ~Handle();
    void frontend_init(const wfs_tables &t);
    void frontend_release();
    void lanes_release();
};

void pulse_call_ranks(const int32_t *group_of, int64_t n_pc, int64_t n_groups,
                      std::vector<int32_t> &rank, int32_t &max_rank);

}  // namespace wfs
