// C-ABI of libwfsim_b200.so (declared in include/wfsim_b200.h).
#include "handle.cuh"
#include <stdlib.h>

#include <string.h>
#include <algorithm>
#include <map>

using namespace wfs;

static thread_local std::string g_create_error;

namespace wfs {

template <typename T>
static T *upload(const T *host, size_t count, std::vector<void *> &owned) {
    if (!host || count == 0) return nullptr;
    T *d = nullptr;
    WFS_CUDA_CHECK(cudaMalloc((void **)&d, count * sizeof(T)));
    WFS_CUDA_CHECK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    owned.push_back(d);
    return d;
}

Handle::Handle(const wfs_params &p, const wfs_tables &t, int dev) : device(dev) {
    WFS_CUDA_CHECK(cudaSetDevice(device));
    WFS_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    WFS_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    WFS_CUDA_CHECK(cudaEventCreate(&ev_a));
    WFS_CUDA_CHECK(cudaEventCreate(&ev_b));
    WFS_CUDA_CHECK(cudaEventCreate(&ev_c));
    WFS_CUDA_CHECK(cudaEventCreate(&ev_d));
    cfg.p = p;
    if (p.dt <= 0 || p.dt > 16 || p.template_length <= 0 || p.template_length > 32 || p.dt * p.template_length > 352)
        throw std::runtime_error("unsupported sample_duration / template length");
    if (p.n_tpc_pmts <= 0 || p.n_tpc_pmts >= (1 << kChannelBits) || p.n_rows > (1 << kChannelBits))
        throw std::runtime_error("unsupported channel count");
    // Pulse.add_current writes template_length samples from (photon sample - pulse left); the pulse
    // array (pulse.py:118-128) only holds them if the right margin covers the template
    if (p.template_length > p.pulse_right_margin + 1 || p.pulse_left_margin < 0 || p.trigger_window < 0)
        throw std::runtime_error("template longer than the stored pulse margin (samples_to_store_after)");
    if (!t.templates || !t.gains || !t.zle_thresholds)
        throw std::runtime_error("templates, gains and zle_thresholds tables are required");
    cfg.templates = upload(t.templates, (size_t)p.dt * p.template_length, owned);
    cfg.gains = upload(t.gains, (size_t)p.n_tpc_pmts, owned);
    cfg.zle_thr = upload(t.zle_thresholds, (size_t)p.n_rows, owned);
    h_gains.assign(t.gains, t.gains + p.n_tpc_pmts);
    {
        std::vector<double> cmax((size_t)p.dt, 0.0);
        for (int r = 0; r < p.dt; r++)
            for (int k = 0; k < p.template_length; k++)
                cmax[r] = std::max(cmax[r], t.templates[(size_t)r * p.template_length + k]);
        cfg.current_max = upload(cmax.data(), cmax.size(), owned);
    }
    if (p.enable_noise && t.noise && t.noise_len > 0 && t.noise_nch > 0) {
        // transpose to [channel][sample]: a window reads consecutive samples of one channel
        std::vector<double> tr((size_t)t.noise_len * t.noise_nch);
        for (int64_t i = 0; i < t.noise_len; i++)
            for (int ch = 0; ch < t.noise_nch; ch++)
                tr[(size_t)ch * t.noise_len + i] = t.noise[(size_t)i * t.noise_nch + ch];
        cfg.noise_t = upload(tr.data(), tr.size(), owned);
        cfg.noise_len = t.noise_len;
        cfg.noise_nch = t.noise_nch;
    }
    {
        const int base = std::max(p.baseline, 0);
        const bool noisy = p.enable_noise && cfg.noise_t;
        for (int ch = 0; ch < p.n_tpc_pmts; ch++) {
            if (t.zle_thresholds[ch] > base) cfg.thr_above_baseline = true;
            if (p.detector_nt && ch < p.n_top_pmts) {
                const int hch = p.he_first + ch;
                if (p.he_mult != 0 || (noisy && hch < cfg.noise_nch) || (hch < p.n_rows && t.zle_thresholds[hch] > p.baseline))
                    cfg.he_rows_possible = true;
            }
        }
    }
    if (const char *e = getenv("WFS_COMPACT")) compact_mode = atoi(e);
    backend = new Backend(&cfg, stream, &launches);
    frontend_init(t);
}

Handle::~Handle() {
    cudaSetDevice(device);
    frontend_release();
    cstage.release();
    delete pool;
    delete backend;
    for (void *p : owned) cudaFree(p);
    DevBuf *bufs[] = {&d_t, &d_ch, &d_gain, &d_pc, &d_pc_group, &d_pc_rank, &d_ix, &d_records, &d_groups,
                      &d_opt_ch, &d_opt_t};
    for (DevBuf *b : bufs) b->release();
    cudaEventDestroy(ev_a); cudaEventDestroy(ev_b); cudaEventDestroy(ev_c); cudaEventDestroy(ev_d);
    cudaStreamDestroy(stream);
    cudaStreamDestroy(copy_stream);
}

// rank of every pulse call inside its group (host side; n_pulse_calls is small)
void pulse_call_ranks(const int32_t *group_of, int64_t n_pc, int64_t n_groups,
                      std::vector<int32_t> &rank, int32_t &max_rank) {
    std::vector<int32_t> cnt((size_t)std::max<int64_t>(n_groups, 1), 0);
    rank.resize((size_t)n_pc);
    max_rank = 0;
    for (int64_t i = 0; i < n_pc; i++) {
        int32_t g = group_of[i];
        if (g < 0 || g >= n_groups) throw std::runtime_error("group_of entry out of range");
        rank[i] = cnt[g]++;
        max_rank = std::max(max_rank, rank[i]);
    }
}

}  // namespace wfs

#define API_BEGIN(h)                      \
    Handle *H = reinterpret_cast<Handle *>(h); \
    if (!H) return WFS_E_ARG;             \
    try {                                 \
        WFS_CUDA_CHECK(cudaSetDevice(H->device));

#define API_END                                                        \
    } catch (const std::exception &e) {                                \
        H->last_error = e.what();                                      \
        cudaGetLastError();                                            \
        return H->last_error.find("CUDA") != std::string::npos ? WFS_E_CUDA : WFS_E_ARG; \
    }

extern "C" {

int wfs_abi_version(void) { return WFS_ABI_VERSION; }

void wfs_struct_sizes(int64_t *out) {
    out[0] = sizeof(wfs_params);
    out[1] = sizeof(wfs_tables);
    out[2] = sizeof(wfs_instr_maps);
    out[3] = sizeof(wfs_counts);
    out[4] = sizeof(wfs_group_info);
    out[5] = sizeof(wfs_outputs);
}

int wfs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *wfs_last_error(void *handle) {
    if (!handle) return g_create_error.c_str();
    return reinterpret_cast<Handle *>(handle)->last_error.c_str();
}

int wfs_create(const wfs_params *params, const wfs_tables *tables, int device, void **handle) {
    if (!params || !tables || !handle) { g_create_error = "null argument"; return WFS_E_ARG; }
    if (params->abi_version != WFS_ABI_VERSION) { g_create_error = "ABI version mismatch"; return WFS_E_ARG; }
    *handle = nullptr;
    try {
        int n = 0;
        WFS_CUDA_CHECK(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) throw std::runtime_error("no such CUDA device (the B200 path has no CPU fallback)");
        *handle = new Handle(*params, *tables, device);
    } catch (const std::exception &e) {
        g_create_error = e.what();
        cudaGetLastError();
        return g_create_error.find("CUDA") != std::string::npos ? WFS_E_CUDA : WFS_E_ARG;
    }
    return 0;
}

void wfs_destroy(void *handle) { delete reinterpret_cast<Handle *>(handle); }

void *wfs_host_alloc(int64_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void wfs_host_free(void *p) { if (p) cudaFreeHost(p); }

int wfs_host_register(void *p, int64_t bytes) {
    if (!p || bytes <= 0) return WFS_E_ARG;
    if (cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return WFS_E_CUDA; }
    return 0;
}

int wfs_host_unregister(void *p) {
    if (!p) return WFS_E_ARG;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return WFS_E_CUDA; }
    return 0;
}

int wfs_expand_compact(const void *hdr, const void *blocks, int64_t n_records, uint8_t *records,
                       int fill, int dt, int n_threads) {
    if (n_records < 0 || (n_records > 0 && (!hdr || !records))) return WFS_E_ARG;
    if (n_records == 0) return 0;
    const CompactHdr *h = reinterpret_cast<const CompactHdr *>(hdr);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(blocks);
    if (n_threads <= 1) {
        expand_records(h, b, 0, n_records, records, (int16_t)fill, (int16_t)dt);
        return 0;
    }
    HostPool pool(n_threads);
    ExpandJob job;
    job.pool = &pool;
    job.hdr = h; job.blocks = b; job.n_rec = n_records; job.dst = records;
    job.fill = (int16_t)fill; job.dt = (int16_t)dt;
    job.arm((int)std::min<int64_t>(4 * n_threads, n_records));
    pool.enqueue(&job);
    job.wait();
    return 0;
}

int wfs_simulate_photons(void *handle, int64_t n_photons, const int64_t *t_ns, const int32_t *channel,
                         const double *gain, const int32_t *pulse_call, int64_t n_pulse_calls,
                         const int32_t *group_of, int64_t n_groups, const int64_t *ix_rand,
                         uint64_t seed, int on_device, uint8_t *records, int64_t cap_records,
                         wfs_group_info *groups, wfs_counts *counts) {
    API_BEGIN(handle)
    if (!counts) throw std::runtime_error("counts is required");
    memset(counts, 0, sizeof(*counts));
    if (n_photons < 0 || n_pulse_calls < 0 || n_groups < 0) throw std::runtime_error("negative size");
    if (n_photons > 0 && (!t_ns || !channel || !gain || !pulse_call)) throw std::runtime_error("null photon array");
    if (n_pulse_calls > 0 && !group_of) throw std::runtime_error("null group_of");
    const int64_t launches0 = H->launches.n;
    std::vector<int32_t> rank;
    int32_t max_rank = 0;
    pulse_call_ranks(group_of, n_pulse_calls, n_groups, rank, max_rank);
    cudaStream_t s = H->stream;
    PhotonBatch b;
    b.n = n_photons;
    b.n_pulse_calls = n_pulse_calls;
    b.n_groups = n_groups;
    b.max_rank = max_rank;
    b.seed = seed;
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_a, s));
    H->d_pc_group.reserve(sizeof(int32_t) * std::max<int64_t>(n_pulse_calls, 1));
    H->d_pc_rank.reserve(sizeof(int32_t) * std::max<int64_t>(n_pulse_calls, 1));
    WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_pc_group.p, group_of, sizeof(int32_t) * n_pulse_calls, cudaMemcpyHostToDevice, s));
    WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_pc_rank.p, rank.data(), sizeof(int32_t) * n_pulse_calls, cudaMemcpyHostToDevice, s));
    b.pc_group = H->d_pc_group.as<int32_t>();
    b.pc_rank = H->d_pc_rank.as<int32_t>();
    if (ix_rand) {
        H->d_ix.reserve(sizeof(int64_t) * std::max<int64_t>(n_groups, 1));
        WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_ix.p, ix_rand, sizeof(int64_t) * n_groups, cudaMemcpyHostToDevice, s));
        b.ix_rand = H->d_ix.as<int64_t>();
    }
    uint8_t *d_rec = records;
    // host destination: the records cross PCIe in the compact form and are expanded by host threads
    const bool compact = !on_device && records && H->use_compact(records);
    if (on_device) {
        b.t = t_ns; b.channel = channel; b.gain = gain; b.pulse_call = pulse_call;
    } else {
        H->d_t.reserve(sizeof(int64_t) * std::max<int64_t>(n_photons, 1));
        H->d_ch.reserve(sizeof(int32_t) * std::max<int64_t>(n_photons, 1));
        H->d_gain.reserve(sizeof(double) * std::max<int64_t>(n_photons, 1));
        H->d_pc.reserve(sizeof(int32_t) * std::max<int64_t>(n_photons, 1));
        WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_t.p, t_ns, sizeof(int64_t) * n_photons, cudaMemcpyHostToDevice, s));
        WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_ch.p, channel, sizeof(int32_t) * n_photons, cudaMemcpyHostToDevice, s));
        WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_gain.p, gain, sizeof(double) * n_photons, cudaMemcpyHostToDevice, s));
        WFS_CUDA_CHECK(cudaMemcpyAsync(H->d_pc.p, pulse_call, sizeof(int32_t) * n_photons, cudaMemcpyHostToDevice, s));
        b.t = H->d_t.as<int64_t>(); b.channel = H->d_ch.as<int32_t>();
        b.gain = H->d_gain.as<double>(); b.pulse_call = H->d_pc.as<int32_t>();
        if (!compact) {
            H->d_records.reserve((size_t)WFS_RECORD_BYTES * std::max<int64_t>(cap_records, 1));
            d_rec = H->d_records.as<uint8_t>();
        }
    }
    CompactOut co;
    if (compact) {
        H->cstage.job.wait();
        H->cstage.reserve_device(std::max<int64_t>(cap_records, 1));
        co = H->cstage.out();
    }
    wfs_group_info *d_groups = nullptr;
    if (groups && n_groups > 0) {
        H->d_groups.reserve(sizeof(wfs_group_info) * n_groups);
        d_groups = H->d_groups.as<wfs_group_info>();
    }
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_b, s));
    BackendResult r;
    H->backend->run(b, d_rec, cap_records, d_groups, r, compact ? &co : nullptr);
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_c, s));
    if (r.error) {
        H->last_error = r.error == WFS_E_PULSE_CACHE_TOO_LONG ? "Pulse cache too long" :
                        r.error == WFS_E_KEYBITS ? "too many groups / pulse calls per group for one batch" :
                        "back end error";
        return r.error;
    }
    int rc = 0;
    counts->need_records = r.n_records;
    if (r.n_records > cap_records) {
        rc = WFS_E_CAPACITY;
    } else if (compact && r.n_records > 0) {
        H->cstage.ship(H->host_pool(), s, r.n_records, r.n_blocks, records, H->record_fill(), (int16_t)H->cfg.p.dt);
        counts->d2h_bytes = (int64_t)sizeof(CompactHdr) * r.n_records + kBlockBytes * r.n_blocks;
    } else if (!on_device && r.n_records > 0) {
        WFS_CUDA_CHECK(cudaMemcpyAsync(records, d_rec, (size_t)WFS_RECORD_BYTES * r.n_records, cudaMemcpyDeviceToHost, s));
        counts->d2h_bytes = (int64_t)WFS_RECORD_BYTES * r.n_records;
    }
    if (d_groups)
        WFS_CUDA_CHECK(cudaMemcpyAsync(groups, d_groups, sizeof(wfs_group_info) * n_groups, cudaMemcpyDeviceToHost, s));
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_d, s));
    WFS_CUDA_CHECK(stream_sync(s));
    H->cstage.job.wait();
    float ms;
    WFS_CUDA_CHECK(cudaEventElapsedTime(&ms, H->ev_a, H->ev_b)); counts->ms_h2d = ms;
    WFS_CUDA_CHECK(cudaEventElapsedTime(&ms, H->ev_b, H->ev_c)); counts->ms_total = ms;
    WFS_CUDA_CHECK(cudaEventElapsedTime(&ms, H->ev_c, H->ev_d)); counts->ms_d2h = ms;
    counts->ms_digitize = r.ms_digitize;
    for (int k = 1; k <= 6; k++) counts->ms_phase[k] = r.ms_phase[k];
    for (int k = 0; k < 3; k++) counts->n_records[k] = rc ? 0 : r.n_rec_class[k];
    counts->n_records_total = rc ? 0 : r.n_records;
    counts->n_photons = r.n_valid_photons;
    counts->n_pulses = r.n_pulses;
    counts->n_windows = r.n_windows;
    counts->n_intervals = r.n_intervals;
    counts->n_samples = r.n_samples;
    counts->n_groups = n_groups;
    counts->n_pulse_calls = n_pulse_calls;
    counts->n_batches = 1;
    counts->gpu_launches = H->launches.n - launches0;
    return rc;
    API_END
}

}  // extern "C"
