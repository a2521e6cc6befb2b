// Deterministic back end: photons -> pulses -> windows -> ADC samples -> ZLE -> raw_records.
#pragma once
#include "common.cuh"
#include "transport.cuh"
#include "../../include/wfsim_b200.h"

namespace wfs {

constexpr int kRelTimeBits = 24;    // ns inside a digitisation group (< 1e6 samples * 10 ns)
constexpr int kChannelBits = 10;
constexpr int64_t kMaxGroupSamples = 1000000;  // rawdata.py:219

// Device-resident tables + scalars shared by all kernels.
struct DeviceConfig {
    wfs_params p;
    double *templates = nullptr;     // [dt][template_length]
    double *gains = nullptr;         // [n_tpc_pmts]
    int32_t *zle_thr = nullptr;      // [n_rows]
    double *noise_t = nullptr;       // transposed: [noise_nch][noise_len]
    double *current_max = nullptr;   // [dt] max of each template row (pulse.py:32)
    int64_t noise_len = 0;
    int32_t noise_nch = 0;
    // host-side facts about the tables (set at wfs_create): can a high-energy twin row (rawdata.py:241-249)
    // ever hold data, and is any TPC threshold above the clamped baseline (then every sample is flagged)
    bool he_rows_possible = false;
    bool thr_above_baseline = false;
};

// Photons of one batch, resident in HBM, in arbitrary order.
struct PhotonBatch {
    int64_t n = 0;
    const int64_t *t = nullptr;
    const int32_t *channel = nullptr;
    const double *gain = nullptr;
    const int32_t *pulse_call = nullptr;   // batch-local pulse-call id (or instruction index, see instr_run)
    // generate mode: pulse_call[i] is the instruction index of photon i and the pulse-call id is
    // 2 * instr_run[instruction] + is_afterpulse (flags bit 1); flags bit 0 = double-pe photon
    const int32_t *instr_run = nullptr;
    const uint8_t *flags = nullptr;
    int32_t *trig_dpe_out = nullptr;       // [2 * n_pulse_calls] (total, bottom): pulse.py:255 quirk
    int64_t n_pulse_calls = 0;
    const int32_t *pc_group = nullptr;     // [n_pulse_calls] batch-local group id
    const int32_t *pc_rank = nullptr;      // [n_pulse_calls] rank of the pulse call inside its group
    int32_t max_rank = 0;
    int64_t n_groups = 0;
    const int64_t *ix_rand = nullptr;      // [n_groups] or nullptr -> Philox(seed, first sample of the group's window)
    uint64_t seed = 0;
    // Optional: the photons of group g are the ranges [group_start[r][g], group_start[r][g+1]),
    // r < group_ranges, of the arrays above (device array, [group_ranges][n_groups+1]; the generate
    // path has up to four runs: photons of the primaries, of the secondaries, and the PMT-afterpulse
    // children of either) and no group holds more than max_group_photons of them -> the ordering is
    // done per group in shared memory (Primitives::segment_sort_pairs) instead of by the device-wide
    // radix sort.
    const uint32_t *group_start = nullptr;
    const uint32_t *h_group_start = nullptr;   // the same table on the host (fused back end: size classes)
    int group_ranges = 1;
    int64_t max_group_photons = 0;
    // Optional (generate mode), for the group-resident fused back end (fused.cu): per group a lower bound
    // of its photon times and its first Pulse call (the calls of a group are consecutive), and the bits
    // needed for 2 * (calls of the largest group)
    const int64_t *group_t0 = nullptr;
    const int32_t *group_run0 = nullptr;
    int relpc_bits = 0;
    // Optional per-PMT truth (generate mode): per Pulse call r = pulse-call id >> 1 and PMT, written
    // by the thread that owns the (pulse call, channel) pulse -- pulse.py:257-269 with per_pmt_truth.
    int32_t *pmt_counts = nullptr;         // [n_pulse_calls / 2][4][n_tpc_pmts] n_photon, n_pe, n_photon_trigger, n_pe_trigger
    int64_t *pmt_areas = nullptr;          // [n_pulse_calls / 2][2][n_tpc_pmts] raw_area, raw_area_trigger (x 2^32)
};

struct BackendResult {
    int64_t n_valid_photons = 0, n_pulses = 0, n_windows = 0, n_tiles = 0;
    int64_t n_intervals = 0, n_records = 0, n_samples = 0;
    int64_t n_blocks = 0;            // compact transport: 8-byte blocks in the stream
    int64_t n_plain = 0;             // split transport: records [0, n_plain) stay plain rows in records_out, the
                                     // compact streams hold records [n_plain, n_records)
    int64_t n_dense_tiles = 0;       // digitize tiles that took the gather (dense) path
    int64_t n_rec_class[3] = {0, 0, 0};
    float ms_digitize = 0.f;
    float ms_phase[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // 1 sort, 2 windows, 3 digitize, 4 zle, 5 rec sort, 6 pack
    int error = 0;
    int segment_sorted_photons = 0, segment_sorted_records = 0;   // 1: ordered per group in shared memory
    int fused = 0;                   // 1: the batch went through the group-resident fused kernel
};

class Backend {
public:
    Backend(const DeviceConfig *cfg, cudaStream_t stream, LaunchCounter *lc);
    ~Backend();
    // Runs the whole back end on a device-resident batch.  Records are produced in
    // `records_out` (device pointer, capacity cap_records rows): [tpc | he | aqmon] segments each
    // sorted by (time, channel).  group_info_out: device pointer [n_groups] or nullptr.
    // If cap_records is too small, res.n_records holds the need and nothing is written.
    // With `compact` the records leave in the compact transport form (transport.cuh) instead:
    // headers at compact->hdr[0 .. n_records), res.n_blocks blocks at compact->blocks; cap_records
    // then bounds those buffers and records_out is unused -- unless plain_fraction > 0 (split transport): then
    // records_out (cap_records rows as well) receives the plain rows of the batch, res.n_plain of them are
    // meant to travel as they are and the compact streams start at record res.n_plain.
    void run(const PhotonBatch &b, uint8_t *records_out, int64_t cap_records,
             wfs_group_info *group_info_out, BackendResult &res, const CompactOut *compact = nullptr,
             double plain_fraction = 0.0);
    void release();
    bool fused_eligible(const PhotonBatch &b) const;
    // false: not run / a group outgrew the shared-memory lists -- the caller takes the multi-pass path
    bool run_fused(const PhotonBatch &b, uint8_t *records_out, int64_t cap_records,
                   wfs_group_info *group_info_out, BackendResult &res);

private:
    const DeviceConfig *cfg_;
    cudaStream_t stream_;
    LaunchCounter *lc_;
    Primitives prim_;
    cudaEvent_t ev0_, ev1_;
    cudaEvent_t evp_[8];
    cudaStream_t aux_[5] = {};     // the size classes of the fused back end run side by side
    cudaEvent_t ev_fork_ = nullptr, ev_join_[5] = {};
    int64_t *h_scalars_ = nullptr;   // pinned readback area
    DevBuf keys_, vals_, st_, sg_, flags64_, pulse_first_, pulse_left_, pulse_win_, win_first_pulse_,
        win_meta_, win_scan_, group_tmin_, group_lr_, scalars_, dense_, itv_, itv_nrec_,
        itv_rec0_, rec_keys_, rec_vals_, rec_itv_, group_nitv_, group_ix_, pstart_, flag8_, cta_first_, phq_,
        group_nvalid_, group_out_, group_win_, rec_seg_, fused_lists_, fused_scal_, fused_records_, fused_tkey_,
        fused_gain_, fused_desc_, fused_ginfo_;
    int fused_smem_set_ = 0;
};

}  // namespace wfs
