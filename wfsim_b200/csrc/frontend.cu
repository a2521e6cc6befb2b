// Front end: instruction parsing, batching, Philox sampling kernels, host-side emulation of the
// reference's event scheduler (rawdata.py:38-157), truth rows (rawdata.py:313-375) and the
// wfs_simulate / wfs_run_staged entry points.
#include "frontend_kernels.cuh"

#include <string.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <chrono>
#include <limits>
#include <unordered_map>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <exception>

namespace wfs {

// ---------------------------------------------------------------------------------------------
// instruction rows (wfsim/strax_interface.py:25-42, packed, 70 bytes)
// ---------------------------------------------------------------------------------------------
template <typename T> static inline T rd(const uint8_t *p) { T v; memcpy(&v, p, sizeof(T)); return v; }
template <typename T> static inline void wr(uint8_t *p, T v) { memcpy(p, &v, sizeof(T)); }

static void parse_instructions(const uint8_t *rows, int64_t n, std::vector<HostInstr> &out) {
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const uint8_t *r = rows + i * WFS_INSTRUCTION_BYTES;
        HostInstr &h = out[i];
        h.event_number = rd<int32_t>(r + 0);
        h.type = rd<int8_t>(r + 4);
        h.time = rd<int64_t>(r + 5);
        h.x = rd<float>(r + 13); h.y = rd<float>(r + 17); h.z = rd<float>(r + 21);
        h.amp = rd<int32_t>(r + 25);
        h.recoil = rd<int8_t>(r + 29);
        h.e_dep = rd<float>(r + 30); h.tot_e = rd<float>(r + 34);
        h.g4id = rd<int32_t>(r + 38); h.vol_id = rd<int32_t>(r + 42);
        h.local_field = rd<double>(r + 46);
        h.n_excitons = rd<int32_t>(r + 54);
        h.x_pri = rd<float>(r + 58); h.y_pri = rd<float>(r + 62); h.z_pri = rd<float>(r + 66);
    }
}

// rawdata.py:61 -- evaluated by numpy in float32 (z is a float32 field, v a python scalar)
static inline int64_t signal_time(int64_t time, float z, int type, double v) {
    float q = z / (float)v;
    int k = ((type % 2) + 2) % 2 - 1;   // python modulo: odd -> 0, even -> -1
    q = q * (float)k;
    return time + (int64_t)q;
}

// ---------------------------------------------------------------------------------------------
// host emulation of RawData.__call__ (rawdata.py:66-155): which instructions form a Pulse call
// and which Pulse calls are digitised together.
// ---------------------------------------------------------------------------------------------
struct SchedIn {
    int64_t n_prim = 0;
    std::vector<int64_t> stime;     // signal time per batch-local instruction
    std::vector<int8_t> type;
    std::vector<int32_t> parent;    // primary that spawned the instruction, -1 for primaries
    std::vector<int64_t> pend;      // end of the last pulse of the instruction [ns]; LLONG_MIN = no pulse
    std::vector<std::pair<int32_t, int32_t>> clusters;   // [a, b) ranges over the primaries
    int64_t rext = 100000;
    bool save_full_truth = true;
    double v = 1.0;
};

// n_groups = number of group indices in use.  Only the last one can be without pulses (the counter moves on
// when a non-empty cache is pushed out): it holds the trailing Pulse calls that made nothing (secondaries
// without electrons, S1s without hits) and is reported with n_intervals = -1; *n_real = groups that digitise.
static void schedule(const SchedIn &in, std::vector<Run> &runs, int32_t &n_groups, int32_t *n_real = nullptr) {
    const int64_t n_tot = (int64_t)in.stime.size();
    std::vector<std::vector<int32_t>> children((size_t)in.n_prim);
    for (int64_t i = in.n_prim; i < n_tot; i++) children[in.parent[i]].push_back((int32_t)i);
    std::vector<int32_t> buffer;
    bool have_end = false, cache = false;
    int64_t last_end = 0;
    int32_t group = 0;
    size_t k = 0;
    bool finished = false;
    auto flush = [&]() { if (cache) { group++; cache = false; } };
    auto by_time = [&](int32_t a, int32_t b) { return in.stime[a] < in.stime[b]; };
    runs.clear();
    while (!finished) {
        if (k < in.clusters.size()) {                                          // A)
            for (int32_t i = in.clusters[k].first; i < in.clusters[k].second; i++) buffer.push_back(i);
            k++;
        }
        std::stable_sort(buffer.begin(), buffer.end(), by_time);               // B)
        if (have_end && !buffer.empty() && in.stime[buffer[0]] - last_end > in.rext) flush();   // C)
        std::vector<int32_t> remaining, spawned;
        bool stop = false;
        size_t a = 0;
        while (a < buffer.size()) {                                            // D)
            size_t b = a + 1;
            while (b < buffer.size() && in.stime[buffer[b]] - in.stime[buffer[b - 1]] <= in.rext) b++;
            if (stop) {
                remaining.insert(remaining.end(), buffer.begin() + a, buffer.begin() + b);
                a = b;
                continue;
            }
            static const int kTypes[4] = {1, 2, 4, 6};
            for (int ti = 0; ti < 4; ti++) {
                const int ptype = kTypes[ti];
                std::vector<int32_t> sel;
                for (size_t j = a; j < b; j++) if (in.type[buffer[j]] == ptype) sel.push_back(buffer[j]);
                if (sel.empty()) continue;
                std::vector<std::vector<int32_t>> sets;
                if (ptype == 1 || ptype == 2) {
                    stop = true;
                    if (in.save_full_truth) {
                        for (int32_t s : sel) sets.push_back({s});
                    } else {
                        const int64_t gap = ptype == 1 ? 100 : (int64_t)(0.2 / in.v);
                        sets.push_back({sel[0]});
                        for (size_t j = 1; j < sel.size(); j++) {
                            if (in.stime[sel[j]] - in.stime[sel[j - 1]] > gap) sets.push_back({});
                            sets.back().push_back(sel[j]);
                        }
                    }
                } else {
                    sets.push_back(sel);
                }
                for (auto &set : sets) {
                    Run r;
                    r.type = ptype;
                    r.group = group;
                    r.instr = set;
                    for (int32_t s : set) {
                        if (in.pend[s] != LLONG_MIN) {
                            cache = true;
                            last_end = have_end ? std::max(last_end, in.pend[s]) : in.pend[s];
                            have_end = true;
                        }
                        if (ptype == 2 && s < in.n_prim)
                            spawned.insert(spawned.end(), children[s].begin(), children[s].end());
                    }
                    runs.push_back(std::move(r));
                }
            }
            if (!stop) flush();
            a = b;
        }
        buffer = remaining;
        buffer.insert(buffer.end(), spawned.begin(), spawned.end());
        finished = (k == in.clusters.size()) && buffer.empty();
    }
    flush();
    if (n_real) *n_real = group;
    n_groups = 0;
    for (auto &r : runs) n_groups = std::max(n_groups, r.group + 1);
}

// ---------------------------------------------------------------------------------------------
struct BatchSpec {
    int64_t first_cluster, last_cluster;   // [first, last)
};

struct Plan {
    std::vector<int32_t> gg_lo, gg_hi;   // 'garfield_gas_gap' luminescence rows / fraction per instruction
    std::vector<double> gg_frac;
    std::vector<int64_t> opt_first;   // externally supplied photons: first list index / count per instruction
    std::vector<int32_t> opt_n;       // (empty: the call has none)
    const int32_t *opt_channels = nullptr;
    const int64_t *opt_timings = nullptr;
    int64_t n_opt = 0, opt_cutoff = 0;
    std::vector<HostInstr> instr;            // as given
    std::vector<int64_t> stime;              // signal time per instruction
    std::vector<int64_t> order;              // instruction indices ordered by signal time
    std::vector<int64_t> cluster_start;      // offsets into `order` (+ sentinel)
    std::vector<BatchSpec> batches;
    // map values
    std::vector<double> lce, scg, cy;
    std::vector<double> vd, dl, xo, yo;      // optional (empty = absent)
    std::vector<double> hsr, hsa;            // transverse-diffusion sigmas (radial, azimuthal); empty = off
    std::vector<double> lgap, lgapmax, le0;  // warped gas gap: own, largest of the instruction's S2 call, E0; empty = off
    std::vector<float> pattern;
    std::vector<int32_t> patrow;
    std::vector<uint64_t> rng_id;
    int64_t pattern_rows = 0;
    double scg_default = 0.0;
};

static int64_t env_i64(const char *name, int64_t dflt) {
    const char *s = getenv(name);
    return s ? atoll(s) : dflt;
}

// Signal-time gap behind which nothing of the instructions in front can still arrive: the clustering
// gap (rawdata.py:63) plus the longest delay of a secondary instruction an S2 can spawn (photo-ionisation
// electrons: the end of the coarse delay grid, afterpulse.py:63-80; photo-electric electrons: 6 sigma
// of their delay, :105-115).  Device batches, plugin pieces and GPU shards are all cut at such gaps only.
int default_lanes() { return host_cores_per_rank() >= 12 ? 4 : 3; }

static double quiet_gap_ns(const Handle *H) {
    const wfs_params &p = H->cfg.p;
    const bool secondaries = p.enable_electron_afterpulses && H->frontend && H->frontend->pi_coarse_len > 0;
    double quiet = (double)p.right_raw_extension +
                   (secondaries ? H->frontend->h_pi_coarse_time.back() + 50000.0 : 0.0);
    if (p.enable_gate_afterpulses && p.photoelectric_p > 0.0)
        quiet = std::max(quiet, (double)p.right_raw_extension + p.photoelectric_t_center + p.drift_time_gate +
                                    6.0 * p.photoelectric_t_spread + 50000.0);
    // The reference starts a new digitisation group when the next cluster begins more than rext behind the END of
    // the last pulse (rawdata.py:87-98), not behind its signal time: photons trail the signal time by the S2 width
    // (diffusion, trapping, luminescence, the gate offset: 50 us covers 6 sigma at full drift several times over)
    // and by the longest PMT-afterpulse delay of the tables.  A cut inside that reach could split a group.
    quiet += 50000.0;
    if (p.enable_pmt_afterpulses && H->frontend) quiet += H->frontend->ap_max_delay_ns;
    return quiet;
}

static void make_plan(Handle *H, const uint8_t *rows, int64_t n, const wfs_instr_maps *maps, Plan &P) {
    const wfs_params &p = H->cfg.p;
    parse_instructions(rows, n, P.instr);
    P.stime.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        const HostInstr &h = P.instr[i];
        if (!(h.type == 1 || h.type == 2 || h.type == 4 || h.type == 6))
            throw std::runtime_error("unsupported instruction type (expected 1, 2, 4 or 6)");
        P.stime[i] = signal_time(h.time, h.z, h.type, p.drift_velocity_liquid);
    }
    P.order.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) P.order[i] = i;
    std::stable_sort(P.order.begin(), P.order.end(), [&](int64_t a, int64_t b) { return P.stime[a] < P.stime[b]; });
    P.cluster_start.clear();
    for (int64_t j = 0; j < n; j++)
        if (j == 0 || P.stime[P.order[j]] - P.stime[P.order[j - 1]] > p.right_raw_extension)
            P.cluster_start.push_back(j);
    P.cluster_start.push_back(n);
    // maps
    const int n_ch = p.n_tpc_pmts;
    P.scg_default = maps ? maps->s2_sc_gain_default : 0.0;
    if (P.scg_default == 0.0) P.scg_default = p.s2_secondary_sc_gain / (1.0 + p.p_double_pe_emision);
    P.lce.assign((size_t)n, 1.0);
    P.scg.assign((size_t)n, P.scg_default);
    P.cy.assign((size_t)n, 1.0);
    P.patrow.assign((size_t)n, 0);
    P.gg_lo.clear(); P.gg_hi.clear(); P.gg_frac.clear();
    if (p.s2_luminescence_model == 2) {
        if (!maps || !maps->gg_lo_row || !maps->gg_hi_row || !maps->gg_frac)
            throw std::runtime_error("s2_luminescence_model 'garfield_gas_gap' needs gg_lo_row / gg_hi_row / gg_frac per instruction");
        const int rows = H->frontend ? H->frontend->gg_rows : 0;
        P.gg_lo.assign(maps->gg_lo_row, maps->gg_lo_row + n);
        P.gg_hi.assign(maps->gg_hi_row, maps->gg_hi_row + n);
        P.gg_frac.assign(maps->gg_frac, maps->gg_frac + n);
        for (int64_t i = 0; i < n; i++)
            if (P.gg_lo[i] < 0 || P.gg_lo[i] >= rows || P.gg_hi[i] < 0 || P.gg_hi[i] >= rows)
                throw std::runtime_error("gg_lo_row / gg_hi_row out of range");
    }
    if (maps && maps->opt_first && maps->opt_last && maps->n_opt > 0) {
        if (!maps->opt_channels || !maps->opt_timings) throw std::runtime_error("opt_channels / opt_timings missing");
        P.opt_first.assign((size_t)n, 0);
        P.opt_n.assign((size_t)n, 0);
        for (int64_t i = 0; i < n; i++) {
            const int64_t a = maps->opt_first[i], b = maps->opt_last[i];
            if (b <= a) continue;
            if (a < 0 || b > maps->n_opt || b - a > INT32_MAX) throw std::runtime_error("opt_first / opt_last out of range");
            P.opt_first[i] = a;
            P.opt_n[i] = (int32_t)(b - a);
        }
        P.opt_channels = maps->opt_channels;
        P.opt_timings = maps->opt_timings;
        P.n_opt = maps->n_opt;
        P.opt_cutoff = maps->opt_time_cutoff > 0 ? maps->opt_time_cutoff : 1000000;   // nveto_time_max_cutoff default
    }
    P.rng_id.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) P.rng_id[i] = maps && maps->rng_id ? maps->rng_id[i] : (uint64_t)i;
    if (maps && maps->s1_lce) std::copy(maps->s1_lce, maps->s1_lce + n, P.lce.begin());
    if (maps && maps->s2_sc_gain) std::copy(maps->s2_sc_gain, maps->s2_sc_gain + n, P.scg.begin());
    if (maps && maps->s2_cy_extra) std::copy(maps->s2_cy_extra, maps->s2_cy_extra + n, P.cy.begin());
    P.vd.clear(); P.dl.clear(); P.xo.clear(); P.yo.clear(); P.hsr.clear(); P.hsa.clear();
    if (maps && maps->hdiff_sigma_r && maps->hdiff_sigma_a) {
        P.hsr.assign(maps->hdiff_sigma_r, maps->hdiff_sigma_r + n);
        P.hsa.assign(maps->hdiff_sigma_a, maps->hdiff_sigma_a + n);
    }
    if (maps && maps->drift_velocity) P.vd.assign(maps->drift_velocity, maps->drift_velocity + n);
    if (maps && maps->diffusion_long) P.dl.assign(maps->diffusion_long, maps->diffusion_long + n);
    if (maps && maps->x_obs && maps->y_obs) {
        P.xo.assign(maps->x_obs, maps->x_obs + n);
        P.yo.assign(maps->y_obs, maps->y_obs + n);
    }
    P.lgap.clear(); P.lgapmax.clear(); P.le0.clear();
    if (maps && maps->lum_gap && maps->lum_e0) {
        P.lgap.assign(maps->lum_gap, maps->lum_gap + n);
        P.le0.assign(maps->lum_e0, maps->lum_e0 + n);
        // the reference lays its radial grid from the LARGEST gap of the S2 call down to the wire
        // (s2.py:372-373), i.e. of the type-2 instructions of one cluster (rawdata.py:102-150)
        P.lgapmax.assign((size_t)n, 0.0);
        for (size_t c = 0; c + 1 < P.cluster_start.size(); c++) {
            double mx = 0.0;
            for (int64_t j = P.cluster_start[c]; j < P.cluster_start[c + 1]; j++)
                if (P.instr[P.order[j]].type == 2) mx = std::max(mx, P.lgap[P.order[j]]);
            for (int64_t j = P.cluster_start[c]; j < P.cluster_start[c + 1]; j++)
                P.lgapmax[P.order[j]] = std::max(mx, P.lgap[P.order[j]]);
        }
    } else if (p.s2_luminescence_model == 0 && H->frontend && H->frontend->lum_len <= 0) {
        for (int64_t i = 0; i < n; i++)
            if (P.instr[i].type != 1)
                throw std::runtime_error("s2_luminescence_model 'simple' with enable_gas_gap_warping needs wfs_instr_maps.lum_gap / lum_e0");
    }
    if (p.s1_model_custom)
        for (int64_t i = 0; i < n; i++) {
            const HostInstr &h = P.instr[i];
            if (h.type == 1 && !(h.recoil == 0 || h.recoil == 6 || h.recoil == 20))
                // s1.py:205-217: ER goes to S1.er, which fails in the reference (undefined `units`);
                // anything else raises AttributeError there
                throw std::runtime_error("Recoil type must be ER, NR, alpha or LED (custom S1 model: only NR, alpha and LED are usable)");
        }
    if (maps && maps->pattern && maps->n_pattern_rows > 0) {
        P.pattern_rows = maps->n_pattern_rows;
        P.pattern.assign(maps->pattern, maps->pattern + maps->n_pattern_rows * n_ch);
    } else {
        P.pattern_rows = 1;
        P.pattern.assign((size_t)n_ch, 1.0f);
    }
    if (maps && maps->pattern_row) std::copy(maps->pattern_row, maps->pattern_row + n, P.patrow.begin());
    for (int64_t i = 0; i < n; i++)     // negative: evaluated on the device from the pattern grid
        if (P.patrow[i] >= P.pattern_rows) throw std::runtime_error("pattern_row out of range");
    // batches: contiguous cluster ranges under a photon / sample budget
    const int64_t ph_budget = env_i64("WFS_BATCH_PHOTONS", 48000000);
    const int64_t sample_budget = env_i64("WFS_BATCH_SAMPLES", 1500000000);
    const int64_t instr_budget = env_i64("WFS_BATCH_INSTRUCTIONS", 400000);
    const double quiet = quiet_gap_ns(H);
    const int64_t n_cl = (int64_t)P.cluster_start.size() - 1;
    P.batches.clear();
    int64_t c0 = 0;
    double ph = 0, smp = 0;
    int64_t ni = 0;
    for (int64_t c = 0; c < n_cl; c++) {
        for (int64_t j = P.cluster_start[c]; j < P.cluster_start[c + 1]; j++) {
            const HostInstr &h = P.instr[P.order[j]];
            double e = h.type == 1 ? h.amp * P.lce[P.order[j]] * p.s1_detection_efficiency
                                   : (double)h.amp * P.scg[P.order[j]] * 1.1;
            if (!P.opt_n.empty() && P.opt_n[P.order[j]] > 0) e = (double)P.opt_n[P.order[j]];   // supplied photons
            if (p.enable_pmt_afterpulses) e *= 1.1;
            ph += e;
            smp += std::min(e, (double)n_ch) * 450.0 * (p.detector_nt ? 1.5 : 1.0);
            ni++;
        }
        bool over = ph > ph_budget || smp > sample_budget || ni > instr_budget;
        bool over2 = ph > 2 * ph_budget || smp > 2 * sample_budget || ni > 2 * instr_budget;
        if (over && c + 1 < n_cl) {
            double gap = (double)(P.stime[P.order[P.cluster_start[c + 1]]] - P.stime[P.order[P.cluster_start[c + 1] - 1]]);
            if (gap > quiet || over2) {
                P.batches.push_back({c0, c + 1});
                c0 = c + 1;
                ph = smp = 0;
                ni = 0;
            }
        }
    }
    if (c0 < n_cl) P.batches.push_back({c0, n_cl});
}

// ---------------------------------------------------------------------------------------------
// Lanes: device batches are independent units (closed sets of instruction clusters), so batch k
// runs on lane k % n_lanes -- a stream, a workspace set and a host thread of its own.  While one
// lane's host thread replays the scheduler or waits for a count, the other lane's kernels keep
// the GPU busy.  Output offsets are published in batch order through `Order` (the noise RNG identity of a
// digitisation group is its first sample), so results do not depend on the number of lanes.
struct Lane {
    int id = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_c = nullptr, ev_d = nullptr, ev_done = nullptr;
    Frontend *F = nullptr;
    Backend *B = nullptr;
    int64_t local_batches = 0;
};

struct Order {
    std::mutex mu;
    std::condition_variable cv;
    int64_t out_done = 0;        // batches [0, out_done) have reserved their output ranges
    bool abort = false;
};

struct LaneAborted : std::runtime_error {
    LaneAborted() : std::runtime_error("another lane failed") {}
};

struct PhotonDump {     // wfs_sample_stage
    int stage = 0;
    uint8_t *out = nullptr;
    int64_t cap = 0, n = 0;
    int64_t sec_base = 0;   // secondaries of the batches dumped so far: secondary ids count through the call
};

struct SimOut {
    PhotonDump *dump = nullptr;
    wfs_outputs *out = nullptr;
    wfs_counts *counts = nullptr;
    bool resident = false;
    int64_t n_rec = 0, n_truth = 0, n_groups = 0, n_batches = 0;
    bool overflow = false;
    bool compact = false;        // records travel in the compact transport form (decided once per call)
    bool split = false;          // ... and, the destination being page-locked, partly as plain rows (transport.cuh)
};

static GenCtx make_ctx(Frontend &F, uint64_t seed) {
    GenCtx g;
    memset(&g, 0, sizeof(g));
    g.i_type = F.b_itype.as<int32_t>(); g.i_time = F.b_itime.as<int64_t>();
    g.i_x = F.b_ix.as<float>(); g.i_y = F.b_iy.as<float>(); g.i_z = F.b_iz.as<float>();
    g.i_amp = F.b_iamp.as<int32_t>(); g.i_gidx = F.b_igidx.as<uint64_t>();
    g.i_lce = F.b_ilce.as<double>(); g.i_scg = F.b_iscg.as<double>(); g.i_cy = F.b_icy.as<double>();
    g.i_pat = F.b_ipat.as<int32_t>();
    g.i_vd = F.has_vd ? F.b_ivd.as<double>() : nullptr;
    g.i_dl = F.has_dl ? F.b_idl.as<double>() : nullptr;
    g.i_xo = F.has_xy ? F.b_ixo.as<double>() : nullptr;
    g.i_yo = F.has_xy ? F.b_iyo.as<double>() : nullptr;
    g.i_lgap = F.has_lw ? F.b_ilgap.as<double>() : nullptr;
    g.i_lgapmax = F.has_lw ? F.b_ilgapmax.as<double>() : nullptr;
    g.i_le0 = F.has_lw ? F.b_ile0.as<double>() : nullptr;
    g.i_lavgt = F.has_lw ? F.b_ilavgt.as<double>() : nullptr;
    g.lumw_alpha = F.lumw_alpha; g.lumw_ue = F.lumw_ue; g.lumw_pressure = F.lumw_pressure;
    g.lumw_ra = F.lumw_ra; g.lumw_rw = F.lumw_rw; g.lumw_dr = F.lumw_dr;
    g.i_hsr = F.has_hd ? F.b_ihsr.as<double>() : nullptr;
    g.i_hsa = F.has_hd ? F.b_ihsa.as<double>() : nullptr;
    g.i_recoil = F.b_irecoil.as<int32_t>();
    g.gg_cdf = F.gg_cdf; g.gg_rows = F.gg_rows; g.gg_len = F.gg_len;
    g.i_gglo = F.has_gg ? F.b_igglo.as<int32_t>() : nullptr;
    g.i_gghi = F.has_gg ? F.b_igghi.as<int32_t>() : nullptr;
    g.i_ggfrac = F.has_gg ? F.b_iggfrac.as<double>() : nullptr;
    g.i_ggmean = F.has_gg ? F.b_iggmean.as<double>() : nullptr;
    g.i_optfirst = F.has_opt ? F.b_ioptfirst.as<int64_t>() : nullptr;
    g.i_optn = F.has_opt ? F.b_ioptn.as<int32_t>() : nullptr;
    g.opt_ch = F.H->d_opt_ch.as<int32_t>();
    g.opt_t = F.H->d_opt_t.as<int64_t>();
    g.opt_cutoff = F.H->opt_cutoff;
    g.i_lrow = F.b_ilrow.as<int32_t>();
    g.i_dmean = F.b_dmean.as<double>(); g.i_dspread = F.b_dspread.as<double>();
    g.i_nemit = F.b_nemit.as<uint32_t>(); g.i_emitoff = F.b_emitoff.as<uint32_t>();
    g.i_nhits = F.b_nhits.as<int64_t>(); g.i_acc = F.b_acc.as<int64_t>();
    g.cdf = F.b_cdf.as<double>(); g.cdf_ok = F.b_cdfok.as<int32_t>(); g.cdf_guide = F.b_cdfguide.as<uint16_t>();
    g.e_t = F.b_et.as<int64_t>(); g.e_instr = F.b_einstr.as<int32_t>();
    g.e_nph = F.b_enph.as<uint32_t>(); g.e_phoff = F.b_ephoff.as<uint32_t>();
    g.ph_t = F.b_pht.as<int64_t>(); g.ph_ch = F.b_phch.as<int32_t>(); g.ph_gain = F.b_phgain.as<double>();
    g.ph_instr = F.b_phinstr.as<int32_t>(); g.ph_flags = F.b_phflags.as<uint8_t>();
    g.ph_nap = F.b_phnap.as<uint8_t>(); g.ap_off = F.b_apoff.as<uint32_t>();
    g.spe_ppf = F.spe_ppf; g.spe_row = F.spe_row; g.spe_len = F.spe_len;
    g.lum_cdf = F.lum_cdf; g.lum_t = F.lum_t; g.lum_len = F.lum_len; g.lum_guide = F.lum_guide;
    g.n_ap = F.H->cfg.p.enable_pmt_afterpulses ? F.n_ap : 0;
    for (int e = 0; e < WFS_MAX_AP_ELEMENTS; e++) {
        g.ap_is_uniform[e] = F.ap_is_uniform[e];
        g.ap_delay_cdf[e] = F.ap_delay_cdf[e]; g.ap_delay_len[e] = F.ap_delay_len[e];
        g.ap_delay_bin[e] = F.ap_delay_bin[e];
        g.ap_amp_cdf[e] = F.ap_amp_cdf[e]; g.ap_amp_len[e] = F.ap_amp_len[e];
        g.ap_amp_rows[e] = F.ap_amp_rows[e]; g.ap_amp_bin[e] = F.ap_amp_bin[e];
    }
    g.pi_time = F.pi_coarse_time; g.pi_prob = F.pi_coarse_prob; g.pi_len = F.pi_coarse_len;
    g.s1_op_top = F.s1_op_top; g.s1_op_bottom = F.s1_op_bottom; g.s2_op_top = F.s2_op_top; g.s2_op_bottom = F.s2_op_bottom;
    g.s1_op_nz = F.s1_op_nz; g.s1_op_nu = F.s1_op_nu; g.s2_op_nu = F.s2_op_nu;
    g.s1_op_z0 = F.s1_op_z0; g.s1_op_z1 = F.s1_op_z1; g.s1_op_u0 = F.s1_op_u0; g.s1_op_u1 = F.s1_op_u1;
    g.s2_op_u0 = F.s2_op_u0; g.s2_op_u1 = F.s2_op_u1;
    g.gf_t = F.gf_t; g.gf_x = F.gf_x; g.gf_rows = F.gf_rows; g.gf_cols = F.gf_cols;
    g.seed = seed;
    return g;
}

#define FLAUNCH(kernel, grid, block, ...)                          \
    do {                                                           \
        kernel<<<(grid), (block), 0, s>>>(__VA_ARGS__);            \
        H->launches.n++;                                           \
    } while (0)

// largest number of photons of one instruction in [i0, i1) -> *out (atomicMax; preset to 0)
__global__ void k_max_instr_photons(uint32_t i0, uint32_t i1, const uint32_t *__restrict__ emit_off,
                                    const uint32_t *__restrict__ e_phoff, uint32_t *out) {
    const uint32_t i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = 0;
    if (i < i1) n = e_phoff[emit_off[i + 1]] - e_phoff[emit_off[i]];
    n = __reduce_max_sync(0xffffffffu, n);
    if ((threadIdx.x & 31) == 0 && n) atomicMax(out, n);
}

// first photon of every instruction (photons are laid out instruction by instruction): [n + 1]
// ... and the first PMT-afterpulse child of the instruction's photons (children follow the parent order)
__global__ void k_instr_ph_start(uint32_t n, const uint32_t *__restrict__ emit_off,
                                 const uint32_t *__restrict__ e_phoff, const uint32_t *__restrict__ ap_off,
                                 uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const uint32_t q = e_phoff[emit_off[i]];
    out[i] = q;
    out[n + 1 + i] = ap_off[q];
}

static void grow_instr(Frontend &F, int64_t n_new, int64_t n_old, cudaStream_t s) {
    auto g = [&](DevBuf &b, size_t el) { b.reserve_keep(el * (size_t)n_new, el * (size_t)n_old, s); };
    g(F.b_itype, 4); g(F.b_itime, 8); g(F.b_ix, 4); g(F.b_iy, 4); g(F.b_iz, 4); g(F.b_iamp, 4);
    g(F.b_igidx, 8); g(F.b_ilce, 8); g(F.b_iscg, 8); g(F.b_icy, 8); g(F.b_ipat, 4);
    g(F.b_dmean, 8); g(F.b_dspread, 8); g(F.b_nemit, 4); g(F.b_nhits, 8);
    g(F.b_ivd, 8); g(F.b_idl, 8); g(F.b_ixo, 8); g(F.b_iyo, 8); g(F.b_irecoil, 4); g(F.b_ilrow, 4);
    g(F.b_ioptfirst, 8); g(F.b_ioptn, 4);
    g(F.b_igglo, 4); g(F.b_igghi, 4); g(F.b_iggfrac, 8); g(F.b_iggmean, 8);
    g(F.b_ihsr, 8); g(F.b_ihsa, 8);
    g(F.b_ilgap, 8); g(F.b_ilgapmax, 8); g(F.b_ile0, 8); g(F.b_ilavgt, 8);
    F.b_emitoff.reserve_keep(4 * (size_t)(n_new + 1), 4 * (size_t)(n_old + 1), s);
    F.b_irun.reserve_keep(4 * (size_t)n_new, 0, s);
}

static void grow_photons(Frontend &F, int64_t n_new, int64_t n_old, cudaStream_t s) {
    auto g = [&](DevBuf &b, size_t el) { b.reserve_keep(el * (size_t)n_new, el * (size_t)n_old, s); };
    g(F.b_pht, 8); g(F.b_phch, 4); g(F.b_phgain, 8); g(F.b_phinstr, 4); g(F.b_phflags, 1); g(F.b_phnap, 1);
}

// Generate emitters and photons of instructions [i0, i1); arrays are appended.
static void generate(Handle *H, Frontend &F, cudaStream_t s, uint64_t seed, int64_t i0, int64_t i1,
                     int64_t &n_emit, int64_t &n_ph, int64_t *max_instr_photons = nullptr) {
    const wfs_params &p = H->cfg.p;
    if (i1 <= i0) return;
    GenCtx g = make_ctx(F, seed);
    FLAUNCH(k_instr, div_up(i1 - i0, 128), 128, g, p, (uint32_t)i0, (uint32_t)i1);
    F.prim.exclusive_scan_u32(g.i_nemit, g.i_emitoff, i1, true);
    uint32_t tot;
    WFS_CUDA_CHECK(cudaMemcpyAsync(&tot, g.i_emitoff + i1, 4, cudaMemcpyDeviceToHost, s));
    WFS_CUDA_CHECK(stream_sync(s));
    const int64_t e0 = n_emit, e1 = tot;
    if (e1 >= (int64_t(1) << 31)) throw std::runtime_error("emitter batch too large");
    if (F.has_hd && i0 == 0 && F.n_pattern_rows > F.first_dev_row) {
        // transverse diffusion (s2.py:560-613): the rows of the S2 primaries become the average of the
        // pattern grid over their electrons' displaced positions, now that the electron counts are
        // known; secondaries (pass B) inherit the row of their parent
        if (!F.s2_pat.v || F.s2_pat.nd != 2 || F.s2_pat.npmt > 128 * kDiffuseMaxPerThread)
            throw std::runtime_error("hdiff_sigma_* need a 2-D device-resident S2 pattern grid of at most 1024 PMTs");
        FLAUNCH(k_pattern_diffuse, (unsigned)(i1 - i0), 128, g, (int32_t)F.first_dev_row, F.s2_pat, (int)p.n_tpc_pmts,
                p.tpc_radius, F.b_pattern.as<float>());
        FLAUNCH(k_pattern_cdf, div_up(F.n_pattern_rows, 64), 64, F.n_pattern_rows, (int)p.n_tpc_pmts,
                F.b_pattern.as<float>(), H->cfg.gains, F.b_cdf.as<double>(), F.b_cdfok.as<int32_t>(),
                F.b_cdfguide.as<uint16_t>());
    }
    F.b_et.reserve_keep(8 * (size_t)std::max<int64_t>(e1, 1), 8 * (size_t)e0, s);
    F.b_einstr.reserve_keep(4 * (size_t)std::max<int64_t>(e1, 1), 4 * (size_t)e0, s);
    F.b_enph.reserve_keep(4 * (size_t)std::max<int64_t>(e1, 1), 4 * (size_t)e0, s);
    F.b_ephoff.reserve_keep(4 * (size_t)(e1 + 1), 0, s);
    g = make_ctx(F, seed);
    if (e1 > e0) FLAUNCH(k_emitters, div_up(e1 - e0, 128), 128, g, p, (uint32_t)i1, (uint32_t)e0, (uint32_t)e1);
    F.prim.exclusive_scan_u32(g.e_nph, g.e_phoff, e1, true);
    WFS_CUDA_CHECK(cudaMemcpyAsync(&tot, g.e_phoff + e1, 4, cudaMemcpyDeviceToHost, s));
    uint32_t max_ph = 0;
    if (max_instr_photons) {
        F.b_scal.reserve(64);
        WFS_CUDA_CHECK(cudaMemsetAsync(F.b_scal.p, 0, 4, s));
        FLAUNCH(k_max_instr_photons, div_up(i1 - i0, 256), 256, (uint32_t)i0, (uint32_t)i1, g.i_emitoff, g.e_phoff,
                F.b_scal.as<uint32_t>());
        WFS_CUDA_CHECK(cudaMemcpyAsync(&max_ph, F.b_scal.p, 4, cudaMemcpyDeviceToHost, s));
    }
    WFS_CUDA_CHECK(stream_sync(s));
    if (max_instr_photons) *max_instr_photons = std::max<int64_t>(*max_instr_photons, max_ph);
    const int64_t p0 = n_ph, p1 = tot;
    if (p1 >= (int64_t(1) << 30)) throw std::runtime_error("photon batch too large; lower WFS_BATCH_PHOTONS");
    grow_photons(F, std::max<int64_t>(p1, 1), p0, s);
    g = make_ctx(F, seed);
    if (p1 > p0 && p.s2_luminescence_model == 2 && F.has_gg) {
        // 'garfield_gas_gap': the mean excitation time of every instruction is subtracted (s2.py:450-451)
        const int ny = (int)std::max<int64_t>(1, std::min<int64_t>(64, (p1 - p0) / (std::max<int64_t>(i1 - i0, 1) * 4096)));
        F.b_ggpartial.reserve(sizeof(double) * (size_t)(i1 - i0) * ny);
        FLAUNCH(k_gg_sum, dim3((unsigned)(i1 - i0), (unsigned)ny), 128, g, (uint32_t)i0, (uint32_t)i1, F.b_ggpartial.as<double>());
        FLAUNCH(k_gg_mean, div_up(i1 - i0, 128), 128, g, (uint32_t)i0, (uint32_t)i1, ny, F.b_ggpartial.as<double>());
    }
    if (p1 > p0) {
        FLAUNCH(k_photon_emitter, div_up(e1 - e0, 256), 256, g, (uint32_t)e0, (uint32_t)e1);
        FLAUNCH(k_photons, div_up(p1 - p0, 256), 256, g, p, H->cfg.gains, p.n_tpc_pmts, (uint32_t)e1,
                (uint32_t)p0, (uint32_t)p1);
    }
    n_emit = e1;
    n_ph = p1;
    WFS_CUDA_CHECK(cudaGetLastError());
}

static void write_truth_row(uint8_t *row, const HostInstr &h0, int run_type, int64_t time, float x, float y,
                            float z, int32_t amp, const int64_t *acc_sum /*A_COUNT summed*/,
                            long double ph_mean, long double ph_sigma, long double e_mean,
                            long double e_sigma, int32_t trig_dpe, int32_t trig_dpe_b, const wfs_params &p,
                            float x_mean_e, float y_mean_e) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    memset(row, 0, WFS_TRUTH_BYTES);
    wr<int32_t>(row + 0, h0.event_number);
    wr<int8_t>(row + 4, (int8_t)run_type);
    wr<int64_t>(row + 5, time);
    wr<float>(row + 13, x); wr<float>(row + 17, y); wr<float>(row + 21, z);
    wr<int32_t>(row + 25, amp);
    wr<int8_t>(row + 29, h0.recoil);
    wr<float>(row + 30, h0.e_dep); wr<float>(row + 34, h0.tot_e);
    wr<int32_t>(row + 38, h0.g4id); wr<int32_t>(row + 42, h0.vol_id);
    wr<double>(row + 46, h0.local_field);
    wr<int32_t>(row + 54, h0.n_excitons);
    wr<float>(row + 58, h0.x_pri); wr<float>(row + 62, h0.y_pri); wr<float>(row + 66, h0.z_pri);
    const bool has_ph = acc_sum[A_NPHALL] > 0, has_e = acc_sum[A_NE] > 0;
    const double t_last = has_ph ? (double)acc_sum[A_TMAX] : nan;
    int64_t endtime = time;
    if (has_ph) endtime = (int64_t)(t_last + (double)((p.template_length + 1) * p.dt));   // rawdata.py:347-352
    wr<int64_t>(row + 70, endtime);
    wr<int32_t>(row + 78, (int32_t)acc_sum[A_NE]);
    wr<int32_t>(row + 82, (int32_t)acc_sum[A_NPH]);
    wr<int32_t>(row + 86, (int32_t)(acc_sum[A_NPH] + acc_sum[A_NDPE]));
    wr<int32_t>(row + 90, (int32_t)acc_sum[A_NTRIG]);
    wr<int32_t>(row + 94, (int32_t)(acc_sum[A_NTRIG] + trig_dpe));
    wr<double>(row + 98, (double)acc_sum[A_AREA] / kAreaScale);
    wr<double>(row + 106, (double)acc_sum[A_AREA_TRIG] / kAreaScale);
    wr<int32_t>(row + 114, (int32_t)acc_sum[A_NPH_B]);
    wr<int32_t>(row + 118, (int32_t)(acc_sum[A_NPH_B] + acc_sum[A_NDPE_B]));
    wr<int32_t>(row + 122, (int32_t)acc_sum[A_NTRIG_B]);
    wr<int32_t>(row + 126, (int32_t)(acc_sum[A_NTRIG_B] + trig_dpe_b));
    wr<double>(row + 130, (double)acc_sum[A_AREA_B] / kAreaScale);
    wr<double>(row + 138, (double)acc_sum[A_AREA_TRIG_B] / kAreaScale);
    wr<double>(row + 146, has_ph ? (double)acc_sum[A_TMIN] : nan);
    wr<double>(row + 154, t_last);
    wr<double>(row + 162, has_ph ? (double)ph_mean : nan);
    wr<double>(row + 170, has_ph ? (double)ph_sigma : nan);
    wr<float>(row + 178, x_mean_e);      // rawdata.py:377-390
    wr<float>(row + 182, y_mean_e);
    wr<double>(row + 186, has_e ? (double)acc_sum[A_ETMIN] : nan);
    wr<double>(row + 194, has_e ? (double)acc_sum[A_ETMAX] : nan);
    wr<double>(row + 202, has_e ? (double)e_mean : nan);
    wr<double>(row + 210, has_e ? (double)e_sigma : nan);
}

// np.mean of a float32 array as numpy evaluates it (rawdata.py:362-364 `tb[field] = np.mean(value)` over
// the x / y / z of an instruction cluster): float32 pairwise summation -- plain loop below 8 elements,
// eight interleaved partial sums up to 128 elements, recursive halving (to multiples of 8) above -- and
// a float32 division by the count.
static float np_pairwise_sum_f32(const float *a, size_t n) {
    if (n < 8) {
        float r = 0.f;
        for (size_t i = 0; i < n; i++) r += a[i];
        return r;
    }
    if (n <= 128) {
        float r[8];
        for (int k = 0; k < 8; k++) r[k] = a[k];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; k++) r[k] += a[i + k];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
}

// exact first/second moments of times over the instructions of a run -> mean and population std
static void combine_moments(const std::vector<int32_t> &set, const std::vector<int64_t> &T,
                            const int64_t *acc, int a_n, int a_s, int a_hi2, int a_hilo, int a_lo2,
                            long double &mean, long double &sigma) {
    __int128 S1 = 0, S2 = 0;
    int64_t n = 0;
    const int64_t T0 = T[set[0]];
    for (int32_t i : set) {
        const int64_t *a = acc + (int64_t)i * A_COUNT;
        const int64_t ni = a[a_n];
        if (!ni) continue;
        __int128 s1 = a[a_s];
        __int128 s2 = ((__int128)a[a_hi2] << 24) + ((__int128)a[a_hilo] << 13) + (__int128)a[a_lo2];
        const __int128 d = T[i] - T0;
        S1 += s1 + d * ni;
        S2 += s2 + 2 * d * s1 + d * d * ni;
        n += ni;
    }
    if (!n) { mean = sigma = 0; return; }
    const long double m = (long double)S1 / (long double)n;
    long double var = (long double)S2 / (long double)n - m * m;
    if (var < 0) var = 0;
    mean = (long double)T0 + m;
    sigma = sqrtl(var);
}

static void simulate_batch(Handle *H, Lane &L, Plan &P, int64_t batch_index, const BatchSpec &bs,
                           uint64_t seed, SimOut &so, Order &ord) {
    Frontend &F = *L.F;
    cudaStream_t s = L.stream;
    const wfs_params &p = H->cfg.p;
    const int n_ch = p.n_tpc_pmts;
    const int64_t j0 = P.cluster_start[bs.first_cluster], j1 = P.cluster_start[bs.last_cluster];
    const int64_t nprim = j1 - j0;
    // ---- host SoA of the primaries (signal-time order) ----
    std::vector<int32_t> h_type(nprim), h_amp(nprim), h_pat(nprim);
    std::vector<int64_t> h_time(nprim);
    std::vector<float> h_x(nprim), h_y(nprim), h_z(nprim);
    std::vector<uint64_t> h_gidx(nprim);
    std::vector<double> h_lce(nprim), h_scg(nprim), h_cy(nprim);
    std::vector<int32_t> h_recoil(nprim);
    F.has_gg = !P.gg_lo.empty();
    std::vector<int32_t> h_gglo(F.has_gg ? nprim : 0), h_gghi(F.has_gg ? nprim : 0);
    std::vector<double> h_ggfrac(F.has_gg ? nprim : 0);
    F.has_opt = !P.opt_n.empty();
    std::vector<int64_t> h_optfirst(F.has_opt ? nprim : 0);
    std::vector<int32_t> h_optn(F.has_opt ? nprim : 0);
    F.has_vd = !P.vd.empty(); F.has_dl = !P.dl.empty(); F.has_xy = !P.xo.empty();
    std::vector<double> h_vd(F.has_vd ? nprim : 0), h_dl(F.has_dl ? nprim : 0), h_xo(F.has_xy ? nprim : 0),
        h_yo(F.has_xy ? nprim : 0);
    F.has_hd = !P.hsr.empty();
    std::vector<double> h_hsr(F.has_hd ? nprim : 0), h_hsa(F.has_hd ? nprim : 0);
    F.has_lw = !P.lgap.empty();
    if (F.has_lw && !(F.lumw_dr > 0.0)) throw std::runtime_error("lum_gap / lum_e0 given but wfs_tables.lumw_* is not set");
    std::vector<double> h_lgap(F.has_lw ? nprim : 0), h_lgapmax(F.has_lw ? nprim : 0), h_le0(F.has_lw ? nprim : 0);
    std::unordered_map<int32_t, int32_t> rowmap;
    std::vector<int32_t> rows_used, dev_rows;
    for (int64_t j = 0; j < nprim; j++) {
        const int64_t gi = P.order[j0 + j];
        const HostInstr &h = P.instr[gi];
        h_type[j] = h.type; h_amp[j] = h.amp; h_time[j] = h.time;
        h_x[j] = h.x; h_y[j] = h.y; h_z[j] = h.z;
        h_gidx[j] = P.rng_id[gi];
        h_lce[j] = P.lce[gi]; h_scg[j] = P.scg[gi]; h_cy[j] = P.cy[gi];
        h_recoil[j] = h.recoil;
        if (F.has_opt) { h_optfirst[j] = P.opt_first[gi]; h_optn[j] = P.opt_n[gi]; }
        if (F.has_gg) { h_gglo[j] = P.gg_lo[gi]; h_gghi[j] = P.gg_hi[gi]; h_ggfrac[j] = P.gg_frac[gi]; }
        if (F.has_vd) h_vd[j] = P.vd[gi];
        if (F.has_dl) h_dl[j] = P.dl[gi];
        if (F.has_xy) { h_xo[j] = P.xo[gi]; h_yo[j] = P.yo[gi]; }
        if (F.has_lw) { h_lgap[j] = P.lgap[gi]; h_lgapmax[j] = P.lgapmax[gi]; h_le0[j] = P.le0[gi]; }
        if (F.has_hd) {
            h_hsr[j] = P.hsr[gi]; h_hsa[j] = P.hsa[gi];
            if (h.type != 1 && P.patrow[gi] >= 0)
                throw std::runtime_error("hdiff_sigma_* need the device-resident S2 pattern grid (pattern_row < 0)");
        }
        if (P.patrow[gi] < 0) {          // pattern evaluated on the device from the uploaded grid
            dev_rows.push_back((int32_t)j);
            continue;
        }
        auto it = rowmap.find(P.patrow[gi]);
        if (it == rowmap.end()) {
            it = rowmap.emplace(P.patrow[gi], (int32_t)rows_used.size()).first;
            rows_used.push_back(P.patrow[gi]);
        }
        h_pat[j] = it->second;
    }
    const int64_t n_host_rows = (int64_t)rows_used.size();
    for (size_t k = 0; k < dev_rows.size(); k++) {
        const int32_t j = dev_rows[k];
        const PatGrid &g = h_type[j] == 1 ? F.s1_pat : F.s2_pat;
        if (!g.v) throw std::runtime_error("pattern_row < 0 but no pattern grid was given for this signal type (wfs_tables)");
        h_pat[j] = (int32_t)(n_host_rows + (int64_t)k);
    }
    grow_instr(F, std::max<int64_t>(nprim, 1), 0, s);
    auto up = [&](DevBuf &b, const void *src, size_t bytes) {
        WFS_CUDA_CHECK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    };
    up(F.b_itype, h_type.data(), 4 * nprim); up(F.b_itime, h_time.data(), 8 * nprim);
    up(F.b_ix, h_x.data(), 4 * nprim); up(F.b_iy, h_y.data(), 4 * nprim); up(F.b_iz, h_z.data(), 4 * nprim);
    up(F.b_iamp, h_amp.data(), 4 * nprim); up(F.b_igidx, h_gidx.data(), 8 * nprim);
    up(F.b_ilce, h_lce.data(), 8 * nprim); up(F.b_iscg, h_scg.data(), 8 * nprim);
    up(F.b_icy, h_cy.data(), 8 * nprim); up(F.b_ipat, h_pat.data(), 4 * nprim);
    up(F.b_irecoil, h_recoil.data(), 4 * nprim);
    if (F.has_opt) { up(F.b_ioptfirst, h_optfirst.data(), 8 * nprim); up(F.b_ioptn, h_optn.data(), 4 * nprim); }
    if (F.has_gg) {
        up(F.b_igglo, h_gglo.data(), 4 * nprim); up(F.b_igghi, h_gghi.data(), 4 * nprim);
        up(F.b_iggfrac, h_ggfrac.data(), 8 * nprim);
    }
    if (F.has_vd) up(F.b_ivd, h_vd.data(), 8 * nprim);
    if (F.has_dl) up(F.b_idl, h_dl.data(), 8 * nprim);
    if (F.has_xy) { up(F.b_ixo, h_xo.data(), 8 * nprim); up(F.b_iyo, h_yo.data(), 8 * nprim); }
    if (F.has_hd) { up(F.b_ihsr, h_hsr.data(), 8 * nprim); up(F.b_ihsa, h_hsa.data(), 8 * nprim); }
    if (F.has_lw) {
        up(F.b_ilgap, h_lgap.data(), 8 * nprim); up(F.b_ilgapmax, h_lgapmax.data(), 8 * nprim);
        up(F.b_ile0, h_le0.data(), 8 * nprim);
    }
    // pattern rows of this batch -> CDF rows
    const int64_t nrows = n_host_rows + (int64_t)dev_rows.size();
    std::vector<float> h_rows((size_t)n_host_rows * n_ch);
    for (int64_t r = 0; r < n_host_rows; r++)
        memcpy(&h_rows[(size_t)r * n_ch], &P.pattern[(size_t)rows_used[r] * n_ch], sizeof(float) * n_ch);
    // few rows (dummy / constant maps): one thread per row is a long serial chain on the critical
    // path -- keep the CDF rows of the previous batch if the pattern rows are bit-identical
    uint64_t rows_hash = 1469598103934665603ull;
    if (nrows <= 16) {
        const uint8_t *bytes = reinterpret_cast<const uint8_t *>(h_rows.data());
        for (size_t k = 0; k < sizeof(float) * h_rows.size(); k++) rows_hash = (rows_hash ^ bytes[k]) * 1099511628211ull;
    }
    const bool cdf_cached = dev_rows.empty() && nrows > 0 && nrows <= 16 && F.cdf_rows == nrows && F.cdf_hash == rows_hash;
    if (!cdf_cached) {
        F.b_pattern.reserve(sizeof(float) * (size_t)nrows * n_ch);
        F.b_cdf.reserve(sizeof(double) * (size_t)nrows * n_ch);
        F.b_cdfok.reserve(sizeof(int32_t) * nrows);
        F.b_cdfguide.reserve(sizeof(uint16_t) * (size_t)nrows * (kCdfGuide + 1));
        if (n_host_rows) up(F.b_pattern, h_rows.data(), sizeof(float) * h_rows.size());
        if (!dev_rows.empty())
            FLAUNCH(k_pattern_eval, div_up(nprim * 32, 128), 128, (uint32_t)nprim, (int32_t)n_host_rows,
                    F.b_itype.as<int32_t>(), F.b_ix.as<float>(), F.b_iy.as<float>(), F.b_iz.as<float>(),
                    F.has_xy ? F.b_ixo.as<double>() : nullptr, F.has_xy ? F.b_iyo.as<double>() : nullptr,
                    F.b_ipat.as<int32_t>(), F.s1_pat, F.s2_pat, n_ch, F.b_pattern.as<float>());
        FLAUNCH(k_pattern_cdf, div_up(nrows, 64), 64, nrows, n_ch, F.b_pattern.as<float>(), H->cfg.gains,
                F.b_cdf.as<double>(), F.b_cdfok.as<int32_t>(), F.b_cdfguide.as<uint16_t>());
        F.cdf_rows = (dev_rows.empty() && nrows <= 16) ? nrows : -1;
        F.cdf_hash = rows_hash;
    }
    F.first_dev_row = n_host_rows;
    F.n_pattern_rows = nrows;
    // ---- pass A: primaries ----
    WFS_CUDA_CHECK(cudaEventRecord(L.ev_c, s));
    int64_t n_emit = 0, n_ph = 0, max_instr_photons = 0;
    generate(H, F, s, seed, 0, nprim, n_emit, n_ph, &max_instr_photons);
    // ---- secondaries: photo-ionisation electrons of the S2 calls (rawdata.py:193-197) ----
    int64_t ntot = nprim;
    std::vector<int32_t> h_parent;
    std::vector<int64_t> sec_time;
    std::vector<float> sec_x, sec_y, sec_z;
    std::vector<int32_t> sec_amp;
    std::vector<int32_t> sec_type;
    {
        // pass 0: how many secondaries does every S2 primary spawn (photo-ionisation: type 4,
        // rawdata.py:193-197; photo-electric / gate: type 6, rawdata.py:198-201)
        const bool do_pi = p.enable_electron_afterpulses && F.pi_coarse_len > 0 && n_ph > 0;
        const bool do_pe = p.enable_gate_afterpulses && p.photoelectric_p > 0.0 && n_ph > 0;
        uint32_t nsec_pi = 0, nsec_pe = 0;
        GenCtx g = make_ctx(F, seed);
        if (do_pi) {
            F.b_picount.reserve(4 * (size_t)(nprim + 1));
            F.b_pioff.reserve(4 * (size_t)(nprim + 1));
            FLAUNCH(k_photoionization, div_up(nprim * 32, 128), 128, g, p, (uint32_t)nprim, 0,
                    F.b_picount.as<uint32_t>(), nullptr, 0u, nullptr);
            F.prim.exclusive_scan_u32(F.b_picount.as<uint32_t>(), F.b_pioff.as<uint32_t>(), nprim, true);
            WFS_CUDA_CHECK(cudaMemcpyAsync(&nsec_pi, F.b_pioff.as<uint32_t>() + nprim, 4, cudaMemcpyDeviceToHost, s));
        }
        if (do_pe) {
            F.b_pecount.reserve(4 * (size_t)(nprim + 1));
            F.b_peoff.reserve(4 * (size_t)(nprim + 1));
            FLAUNCH(k_photoelectric, div_up(nprim * 32, 128), 128, g, p, (uint32_t)nprim, 0,
                    F.b_pecount.as<uint32_t>(), nullptr, 0u, nullptr);
            F.prim.exclusive_scan_u32(F.b_pecount.as<uint32_t>(), F.b_peoff.as<uint32_t>(), nprim, true);
            WFS_CUDA_CHECK(cudaMemcpyAsync(&nsec_pe, F.b_peoff.as<uint32_t>() + nprim, 4, cudaMemcpyDeviceToHost, s));
        }
        if (do_pi || do_pe) WFS_CUDA_CHECK(stream_sync(s));
        const int64_t nsec = (int64_t)nsec_pi + nsec_pe;
        if (nsec > 0) {
            ntot = nprim + nsec;
            grow_instr(F, ntot, nprim, s);
            struct Scoped : DevBuf { ~Scoped() { release(); } } d_parent;      // freed on every exit path
            d_parent.reserve(4 * (size_t)ntot);
            g = make_ctx(F, seed);
            if (nsec_pi)
                FLAUNCH(k_photoionization, div_up(nprim * 32, 128), 128, g, p, (uint32_t)nprim, 1, nullptr,
                        F.b_pioff.as<uint32_t>(), (uint32_t)nprim, d_parent.as<int32_t>());
            if (nsec_pe)
                FLAUNCH(k_photoelectric, div_up(nprim * 32, 128), 128, g, p, (uint32_t)nprim, 1, nullptr,
                        F.b_peoff.as<uint32_t>(), (uint32_t)(nprim + nsec_pi), d_parent.as<int32_t>());
            h_parent.resize(nsec); sec_time.resize(nsec); sec_x.resize(nsec); sec_y.resize(nsec);
            sec_z.resize(nsec); sec_amp.resize(nsec); sec_type.resize(nsec);
            auto down = [&](void *dst, const void *src, size_t bytes) {
                WFS_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
            };
            down(h_parent.data(), d_parent.as<int32_t>() + nprim, 4 * (size_t)nsec);
            down(sec_time.data(), F.b_itime.as<int64_t>() + nprim, 8 * (size_t)nsec);
            down(sec_x.data(), F.b_ix.as<float>() + nprim, 4 * (size_t)nsec);
            down(sec_y.data(), F.b_iy.as<float>() + nprim, 4 * (size_t)nsec);
            down(sec_z.data(), F.b_iz.as<float>() + nprim, 4 * (size_t)nsec);
            down(sec_amp.data(), F.b_iamp.as<int32_t>() + nprim, 4 * (size_t)nsec);
            down(sec_type.data(), F.b_itype.as<int32_t>() + nprim, 4 * (size_t)nsec);
            WFS_CUDA_CHECK(stream_sync(s));
            d_parent.release();
            generate(H, F, s, seed, nprim, ntot, n_emit, n_ph, &max_instr_photons);   // pass B
        }
    }
    // ---- PMT afterpulses of every photon (rawdata.py:176-178) ----
    int64_t n_ap = 0;
    F.b_apoff.reserve(4 * (size_t)(n_ph + 1));
    GenCtx g = make_ctx(F, seed);
    if (g.n_ap > 0 && n_ph > 0) {
        // ap_off = exclusive scan of ph_nap (widened to u32)
        FLAUNCH(k_widen_u8, div_up(n_ph, 256), 256, g.ph_nap, g.ap_off, (uint32_t)n_ph);
        F.prim.exclusive_scan_u32(g.ap_off, g.ap_off, n_ph, true);
        uint32_t tot;
        WFS_CUDA_CHECK(cudaMemcpyAsync(&tot, g.ap_off + n_ph, 4, cudaMemcpyDeviceToHost, s));
        WFS_CUDA_CHECK(stream_sync(s));
        n_ap = tot;
        if (n_ap > 0) {
            grow_photons(F, n_ph + n_ap, n_ph, s);
            g = make_ctx(F, seed);
            FLAUNCH(k_ap_fill, div_up(n_ph, 256), 256, g, p, H->cfg.gains, (uint32_t)n_emit, (uint32_t)n_ph,
                    (uint32_t)n_ph);
        }
    } else {
        WFS_CUDA_CHECK(cudaMemsetAsync(F.b_apoff.p, 0, 4 * (size_t)(n_ph + 1), s));
    }
    if (so.dump) {   // stage-level dump for the statistical parity tests
        PhotonDump &d = *so.dump;
        const int64_t base = d.n;
        if (d.stage == 0) {
            const int64_t m = n_ph + n_ap;
            std::vector<int64_t> t(m); std::vector<double> gn(m); std::vector<int32_t> ch(m), in_(m);
            std::vector<uint8_t> fl(m);
            WFS_CUDA_CHECK(cudaMemcpyAsync(t.data(), F.b_pht.p, 8 * m, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(gn.data(), F.b_phgain.p, 8 * m, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(ch.data(), F.b_phch.p, 4 * m, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(in_.data(), F.b_phinstr.p, 4 * m, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(fl.data(), F.b_phflags.p, m, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(stream_sync(s));
            for (int64_t q = 0; q < m; q++) {
                if (base + q < d.cap) {
                    uint8_t *r = d.out + (base + q) * 32;
                    const int32_t li = in_[q];
                    const bool sec = li >= nprim;
                    const int32_t root = sec ? h_parent[li - nprim] : li;
                    wr<int64_t>(r, t[q]); wr<double>(r + 8, gn[q]); wr<int32_t>(r + 16, ch[q]);
                    wr<int32_t>(r + 20, (int32_t)P.order[j0 + root]);
                    wr<int32_t>(r + 24, (int32_t)fl[q] | (sec ? 4 : 0));
                    wr<int32_t>(r + 28, sec ? (int32_t)(d.sec_base + li - nprim) : -1);
                }
            }
            d.n += m;
        } else if (d.stage == 4) {
            // secondary instructions in full: t = time, the 8 gain bytes = float32 x, y; channel = amp,
            // instruction = parent, flags = type, secondary = the bits of float32 z
            const int64_t nsec = ntot - nprim;
            for (int64_t q = 0; q < nsec; q++) {
                if (base + q < d.cap) {
                    uint8_t *r = d.out + (base + q) * 32;
                    wr<int64_t>(r, sec_time[q]);
                    wr<float>(r + 8, sec_x[q]); wr<float>(r + 12, sec_y[q]);
                    wr<int32_t>(r + 16, sec_amp[q]);
                    wr<int32_t>(r + 20, (int32_t)P.order[j0 + h_parent[q]]);
                    wr<int32_t>(r + 24, sec_type[q]);
                    wr<float>(r + 28, sec_z[q]);
                }
            }
            d.n += nsec;
        } else if (d.stage >= 2) {
            // secondary instructions (photo-ionisation type 4, photo-electric type 6):
            // t = time, gain = z (stage 2) or x^2 + y^2 (stage 3), channel = amp, flags = type
            const int64_t nsec = ntot - nprim;
            for (int64_t q = 0; q < nsec; q++) {
                if (base + q < d.cap) {
                    uint8_t *r = d.out + (base + q) * 32;
                    const double x = sec_x[q], y = sec_y[q];
                    wr<int64_t>(r, sec_time[q]);
                    wr<double>(r + 8, d.stage == 2 ? (double)sec_z[q] : x * x + y * y);
                    wr<int32_t>(r + 16, sec_amp[q]);
                    wr<int32_t>(r + 20, (int32_t)P.order[j0 + h_parent[q]]);
                    wr<int32_t>(r + 24, sec_type[q]);
                    wr<int32_t>(r + 28, (int32_t)(d.sec_base + q));
                }
            }
            d.n += nsec;
        } else {
            std::vector<int64_t> t(n_emit); std::vector<int32_t> in_(n_emit); std::vector<uint32_t> np_(n_emit);
            WFS_CUDA_CHECK(cudaMemcpyAsync(t.data(), F.b_et.p, 8 * n_emit, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(in_.data(), F.b_einstr.p, 4 * n_emit, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(cudaMemcpyAsync(np_.data(), F.b_enph.p, 4 * n_emit, cudaMemcpyDeviceToHost, s));
            WFS_CUDA_CHECK(stream_sync(s));
            for (int64_t q = 0; q < n_emit; q++) {
                if (base + q < d.cap) {
                    uint8_t *r = d.out + (base + q) * 32;
                    const int32_t li = in_[q];
                    const bool sec = li >= nprim;
                    const int32_t root = sec ? h_parent[li - nprim] : li;
                    wr<int64_t>(r, t[q]); wr<double>(r + 8, 0.0); wr<int32_t>(r + 16, (int32_t)np_[q]);
                    wr<int32_t>(r + 20, (int32_t)P.order[j0 + root]);
                    wr<int32_t>(r + 24, sec ? 4 : 0);
                    wr<int32_t>(r + 28, sec ? (int32_t)(d.sec_base + li - nprim) : -1);
                }
            }
            d.n += n_emit;
        }
        d.sec_base += ntot - nprim;
        return;
    }
    // ---- per-instruction truth accumulators ----
    F.b_acc.reserve(8 * (size_t)ntot * A_COUNT);
    g = make_ctx(F, seed);
    {
        // heavy instructions (S2s with 1e5..1e6 photons) are split over several CTAs
        const bool split = max_instr_photons > (int64_t)kTruthSlice;
        if (split) FLAUNCH(k_acc_init, div_up(ntot * A_COUNT, 256), 256, ntot * (int64_t)A_COUNT, F.b_acc.as<int64_t>());
        FLAUNCH(k_instr_truth, (unsigned)ntot, 128, g, H->cfg, (uint32_t)ntot, (uint32_t)n_ph,
                (uint32_t)n_ph, split ? 1 : 0, (const uint2 *)nullptr, (const uint32_t *)nullptr);
        if (split) {
            const int64_t cap_items = n_ph / kTruthSlice + 1;
            F.b_titems.reserve(sizeof(uint2) * (size_t)cap_items + 16);
            uint32_t *d_n = F.b_titems.as<uint32_t>();                     // [0]: item count; items from byte 16
            uint2 *d_items = reinterpret_cast<uint2 *>(F.b_titems.as<uint8_t>() + 16);
            WFS_CUDA_CHECK(cudaMemsetAsync(d_n, 0, 16, s));
            FLAUNCH(k_truth_items, div_up(ntot, 256), 256, g, (uint32_t)ntot, d_items, d_n);
            FLAUNCH(k_instr_truth, (unsigned)cap_items, 128, g, H->cfg, (uint32_t)ntot, (uint32_t)n_ph, (uint32_t)n_ph, 1,
                    (const uint2 *)d_items, (const uint32_t *)d_n);
        }
    }
    WFS_CUDA_CHECK(cudaEventRecord(L.ev_d, s));
    std::vector<int64_t> acc((size_t)ntot * A_COUNT);
    WFS_CUDA_CHECK(cudaMemcpyAsync(acc.data(), F.b_acc.p, 8 * acc.size(), cudaMemcpyDeviceToHost, s));
    std::vector<uint32_t> ph_start(2 * ((size_t)ntot + 1), 0u);     // [ntot + 1] photon starts, [ntot + 1] afterpulse starts
    if (n_ph > 0) {
        F.b_phstart.reserve(4 * ph_start.size());
        FLAUNCH(k_instr_ph_start, div_up(ntot + 1, 256), 256, (uint32_t)ntot, g.i_emitoff, g.e_phoff, g.ap_off,
                F.b_phstart.as<uint32_t>());
        WFS_CUDA_CHECK(cudaMemcpyAsync(ph_start.data(), F.b_phstart.p, 4 * ph_start.size(), cudaMemcpyDeviceToHost, s));
    }
    WFS_CUDA_CHECK(stream_sync(s));
    float ms_front = 0;
    cudaEventElapsedTime(&ms_front, L.ev_c, L.ev_d);
    const auto host_t0 = std::chrono::steady_clock::now();
    // ---- host scheduler ----
    SchedIn in;
    in.n_prim = nprim;
    in.rext = p.right_raw_extension;
    in.save_full_truth = p.save_full_truth != 0;
    in.v = p.drift_velocity_liquid;
    in.stime.resize(ntot); in.type.resize(ntot); in.parent.assign(ntot, -1); in.pend.resize(ntot);
    std::vector<int64_t> T(ntot);
    for (int64_t j = 0; j < nprim; j++) {
        in.stime[j] = P.stime[P.order[j0 + j]];
        in.type[j] = (int8_t)h_type[j];
        T[j] = h_time[j];
    }
    for (int64_t j = nprim; j < ntot; j++) {
        in.type[j] = (int8_t)sec_type[j - nprim];
        in.parent[j] = h_parent[j - nprim];
        T[j] = sec_time[j - nprim];
        in.stime[j] = signal_time(T[j], sec_z[j - nprim], 4, p.drift_velocity_liquid);
    }
    for (int64_t j = 0; j < ntot; j++) {
        const int64_t tmax = acc[(size_t)j * A_COUNT + A_PTMAX];
        if (tmax == LLONG_MIN) in.pend[j] = LLONG_MIN;
        else {
            int64_t q = tmax / p.dt; if (tmax % p.dt != 0 && tmax < 0) q--;
            in.pend[j] = (q + p.pulse_right_margin) * p.dt;
        }
    }
    for (int64_t c = bs.first_cluster; c < bs.last_cluster; c++)
        in.clusters.push_back({(int32_t)(P.cluster_start[c] - j0), (int32_t)(P.cluster_start[c + 1] - j0)});
    std::vector<Run> runs;
    int32_t ngroups = 0;
    schedule(in, runs, ngroups);
    const int64_t nruns = (int64_t)runs.size();
    const int64_t npc = 2 * nruns;
    std::vector<int32_t> instr_run((size_t)ntot, -1), pc_group((size_t)std::max<int64_t>(npc, 1)),
        pc_rank;
    for (int64_t r = 0; r < nruns; r++) {
        for (int32_t i : runs[r].instr) instr_run[i] = (int32_t)r;
        pc_group[2 * r] = pc_group[2 * r + 1] = runs[r].group;
    }
    int32_t max_rank = 0;
    pulse_call_ranks(pc_group.data(), npc, ngroups, pc_rank, max_rank);
    F.b_irun.reserve(4 * (size_t)ntot);
    F.b_pcgroup.reserve(4 * (size_t)std::max<int64_t>(npc, 1));
    F.b_pcrank.reserve(4 * (size_t)std::max<int64_t>(npc, 1));
    F.b_trig.reserve(4 * (size_t)std::max<int64_t>(2 * npc, 1));
    up(F.b_irun, instr_run.data(), 4 * (size_t)ntot);
    if (npc) {
        up(F.b_pcgroup, pc_group.data(), 4 * (size_t)npc);
        up(F.b_pcrank, pc_rank.data(), 4 * (size_t)npc);
        WFS_CUDA_CHECK(cudaMemsetAsync(F.b_trig.p, 0, 4 * (size_t)(2 * npc), s));
    }
    // Photons are laid out instruction by instruction in up to four runs: photons of the primaries, of
    // the secondaries, and behind them the PMT-afterpulse children of either (in parent order).  If in
    // every run the groups the scheduler formed follow each other (the common case), every group is the
    // concatenation of one range per run and the back end can order it in shared memory.
    std::vector<uint32_t> gstart;
    int64_t max_group_photons = 0;
    // records the batch is expected to make: a photon alone on its channel makes one record, photons that pile up
    // on a channel share records -- at most ~40 per (group, channel) window counted here; an underestimate only
    // costs the repeated back end below, an overestimate of 2 records per photon cost 23 GB per buffer for heavy S2s
    int64_t est_records = -1;
    const int n_ranges = 4;
    {
        bool contiguous = n_ph > 0 && ngroups > 0;
        const uint32_t *ap_start = ph_start.data() + (ntot + 1);
        if (contiguous) {
            gstart.assign((size_t)n_ranges * (ngroups + 1), UINT32_MAX);
            for (int cls = 0; cls < n_ranges && contiguous; cls++) {
                const int64_t i0 = (cls & 1) ? nprim : 0, i1 = (cls & 1) ? ntot : nprim;
                auto first_of = [&](int64_t i) -> uint32_t {
                    return cls < 2 ? ph_start[i] : (uint32_t)n_ph + ap_start[i];
                };
                uint32_t *gs = &gstart[(size_t)cls * (ngroups + 1)];
                int32_t last_group = -1;
                for (int64_t i = i0; i < i1 && contiguous; i++) {
                    if (first_of(i + 1) == first_of(i) || instr_run[i] < 0) continue;   // nothing / in no Pulse call
                    const int32_t gi = runs[instr_run[i]].group;
                    if (gi < last_group) contiguous = false;
                    else if (gi > last_group) { gs[gi] = first_of(i); last_group = gi; }
                }
                gs[ngroups] = first_of(i1);
                for (int32_t gi = ngroups - 1; gi >= 0; gi--)
                    if (gs[gi] == UINT32_MAX) gs[gi] = gs[gi + 1];                     // group without photons in this run
                gs[0] = first_of(i0);     // photons in front of the first group belong to no Pulse call: dropped as invalid
            }
        }
        if (contiguous) {
            est_records = 0;
            for (int32_t gi = 0; gi < ngroups; gi++) {
                int64_t cnt = 0;
                for (int cls = 0; cls < n_ranges; cls++)
                    cnt += gstart[(size_t)cls * (ngroups + 1) + gi + 1] - gstart[(size_t)cls * (ngroups + 1) + gi];
                max_group_photons = std::max(max_group_photons, cnt);
                est_records += std::min<int64_t>(cnt, (int64_t)40 * n_ch);
            }
            F.b_gstart.reserve(4 * gstart.size());
            up(F.b_gstart, gstart.data(), 4 * gstart.size());
        } else {
            gstart.clear();
        }
        if (getenv("WFS_DEBUG_SEG"))
            fprintf(stderr, "[wfs] batch %lld: n_ph %lld n_ap %lld groups %d contiguous %d max_group_photons %lld\n",
                    (long long)batch_index, (long long)n_ph, (long long)n_ap, ngroups, (int)contiguous,
                    (long long)max_group_photons);
    }
    // for the group-resident fused back end: first Pulse call of every group (the calls of a group are
    // consecutive in execution order) and a lower bound of its photon times (PMT afterpulses may precede
    // their parent by pmt_ap_t_modifier, afterpulse.py:219-223)
    int relpc_bits = 0;
    if (!gstart.empty()) {
        std::vector<int64_t> g_t0((size_t)ngroups, LLONG_MAX);
        std::vector<int32_t> g_run0((size_t)ngroups, 0), g_nruns((size_t)ngroups, 0);
        const int64_t ap_margin = p.enable_pmt_afterpulses ? (int64_t)std::ceil(std::max(p.pmt_ap_t_modifier, 0.0)) + 1 : 0;
        for (int64_t r = 0; r < nruns; r++) {
            const int32_t gi = runs[r].group;
            if (g_nruns[gi]++ == 0) g_run0[gi] = (int32_t)r;
            for (int32_t i : runs[r].instr)
                if (acc[(size_t)i * A_COUNT + A_NPHALL] > 0)
                    g_t0[gi] = std::min(g_t0[gi], acc[(size_t)i * A_COUNT + A_TMIN] - ap_margin);
        }
        int32_t max_runs = 1;
        for (int32_t gi = 0; gi < ngroups; gi++) max_runs = std::max(max_runs, g_nruns[gi]);
        while ((int64_t(1) << relpc_bits) < 2 * (int64_t)max_runs) relpc_bits++;
        F.b_gt0.reserve(8 * (size_t)ngroups);
        F.b_grun0.reserve(4 * (size_t)ngroups);
        up(F.b_gt0, g_t0.data(), 8 * (size_t)ngroups);
        up(F.b_grun0, g_run0.data(), 4 * (size_t)ngroups);
    }
    const double ms_host = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    {
        std::lock_guard<std::mutex> lk(ord.mu);
        if (ord.abort) throw LaneAborted();
    }
    // ---- back end ----
    PhotonBatch b;
    b.n = n_ph + n_ap;
    b.t = F.b_pht.as<int64_t>(); b.channel = F.b_phch.as<int32_t>(); b.gain = F.b_phgain.as<double>();
    b.pulse_call = F.b_phinstr.as<int32_t>();
    b.instr_run = F.b_irun.as<int32_t>();
    b.flags = F.b_phflags.as<uint8_t>();
    b.trig_dpe_out = F.b_trig.as<int32_t>();
    b.n_pulse_calls = npc;
    b.pc_group = F.b_pcgroup.as<int32_t>(); b.pc_rank = F.b_pcrank.as<int32_t>();
    b.max_rank = max_rank;
    b.n_groups = ngroups;
    b.seed = seed;
    if (!gstart.empty()) {
        b.group_start = F.b_gstart.as<uint32_t>();
        b.h_group_start = gstart.data();
        b.group_ranges = n_ranges;
        b.max_group_photons = max_group_photons;
        b.group_t0 = F.b_gt0.as<int64_t>();
        b.group_run0 = F.b_grun0.as<int32_t>();
        b.relpc_bits = relpc_bits;
    }
    const bool per_pmt = so.out && so.out->truth_pmt_counts && so.out->truth_pmt_areas && nruns > 0;
    if (per_pmt) {
        F.b_pmtcnt.reserve(sizeof(int32_t) * 4 * (size_t)n_ch * (size_t)nruns);
        F.b_pmtarea.reserve(sizeof(int64_t) * 2 * (size_t)n_ch * (size_t)nruns);
        WFS_CUDA_CHECK(cudaMemsetAsync(F.b_pmtcnt.p, 0, sizeof(int32_t) * 4 * (size_t)n_ch * (size_t)nruns, s));
        WFS_CUDA_CHECK(cudaMemsetAsync(F.b_pmtarea.p, 0, sizeof(int64_t) * 2 * (size_t)n_ch * (size_t)nruns, s));
        b.pmt_counts = F.b_pmtcnt.as<int32_t>();
        b.pmt_areas = F.b_pmtarea.as<int64_t>();
    }
    BackendResult res;
    wfs_outputs *out = so.out;
    F.b_groups.reserve(sizeof(wfs_group_info) * (size_t)std::max<int32_t>(ngroups, 1));
    // records are produced in a device buffer of the lane; with host outputs two buffers alternate:
    // batch k travels to the host on the copy stream while the lane's next batch runs.  Host
    // destinations get the compact transport form (transport.cuh) unless it is switched off.
    const int par = so.resident ? 0 : (int)(L.local_batches & 1);
    DevBuf &rb = par ? F.b_records2 : F.b_records;
    CompactStage &cs = F.cstage[par];
    if (!so.resident) {
        if (F.copy_pending[par]) {
            WFS_CUDA_CHECK(cudaEventSynchronize(F.ev_copy[par]));
            F.copy_pending[par] = false;
        }
        cs.job.wait();
    }
    const bool want_records = so.resident || (out && out->records);
    const bool compact = !so.resident && want_records && so.compact;
    CompactOut co;
    int64_t cap_here = 0;
    const double plain_fraction = compact && so.split ? H->split.fraction() : 0.0;
    auto reserve_records = [&](int64_t n_rec) {
        if (compact) {
            cs.reserve_device(n_rec);
            cap_here = cs.cap_records();
            co = cs.out();
            if (plain_fraction > 0.0) {
                rb.reserve((size_t)WFS_RECORD_BYTES * (size_t)n_rec);
                cap_here = std::min(cap_here, (int64_t)(rb.cap / WFS_RECORD_BYTES));
            }
        } else if (want_records) {
            rb.reserve((size_t)WFS_RECORD_BYTES * (size_t)n_rec);
            cap_here = (int64_t)(rb.cap / WFS_RECORD_BYTES);
        }
    };
    reserve_records(est_records >= 0 ? est_records + est_records / 2 + 65536 : std::max<int64_t>(2 * (n_ph + n_ap) + 65536, 1));
    uint8_t *d_rec = rb.as<uint8_t>();
    if (ngroups > 0) {
        L.B->run(b, d_rec, cap_here, F.b_groups.as<wfs_group_info>(), res, compact ? &co : nullptr, plain_fraction);
        if (res.error) {
            // reported through failure[] of run_plan: lanes never write the handle's error string
            throw std::runtime_error(res.error == WFS_E_PULSE_CACHE_TOO_LONG ? "Pulse cache too long"
                                                                             : "back end error (key bits)");
        }
        if (want_records && res.n_records > cap_here) {
            // the device buffer was too small: grow it and redo the back end
            reserve_records(res.n_records);
            d_rec = rb.as<uint8_t>();
            WFS_CUDA_CHECK(cudaMemsetAsync(F.b_trig.p, 0, 4 * (size_t)std::max<int64_t>(2 * npc, 1), s));
            if (per_pmt)   // the fused kernel adds the per-PMT areas photon by photon
                WFS_CUDA_CHECK(cudaMemsetAsync(F.b_pmtarea.p, 0, sizeof(int64_t) * 2 * (size_t)n_ch * (size_t)nruns, s));
            L.B->run(b, d_rec, cap_here, F.b_groups.as<wfs_group_info>(), res, compact ? &co : nullptr, plain_fraction);
        }
    }
    std::vector<wfs_group_info> h_groups((size_t)ngroups);
    if (ngroups)
        WFS_CUDA_CHECK(cudaMemcpyAsync(h_groups.data(), F.b_groups.p, sizeof(wfs_group_info) * ngroups,
                                       cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> trig((size_t)std::max<int64_t>(2 * npc, 1), 0);
    if (npc) WFS_CUDA_CHECK(cudaMemcpyAsync(trig.data(), F.b_trig.p, 4 * (size_t)(2 * npc), cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> pmt_cnt;
    std::vector<int64_t> pmt_area;
    if (per_pmt) {
        pmt_cnt.resize((size_t)4 * n_ch * nruns);
        pmt_area.resize((size_t)2 * n_ch * nruns);
        WFS_CUDA_CHECK(cudaMemcpyAsync(pmt_cnt.data(), F.b_pmtcnt.p, sizeof(int32_t) * pmt_cnt.size(), cudaMemcpyDeviceToHost, s));
        WFS_CUDA_CHECK(cudaMemcpyAsync(pmt_area.data(), F.b_pmtarea.p, sizeof(int64_t) * pmt_area.size(), cudaMemcpyDeviceToHost, s));
    }
    WFS_CUDA_CHECK(stream_sync(s));
    // truth rows of this batch (rawdata.py:313-375), one per Pulse call
    struct RunSum { int64_t sum[A_COUNT]; bool row; };
    std::vector<RunSum> rsum((size_t)nruns);
    int64_t n_pe = 0, n_truth_here = 0;
    for (int64_t r = 0; r < nruns; r++) {
        const Run &run = runs[r];
        int64_t *sum = rsum[r].sum;
        for (int a = 0; a < A_COUNT; a++) sum[a] = 0;
        sum[A_TMIN] = sum[A_ETMIN] = LLONG_MAX;
        sum[A_TMAX] = sum[A_ETMAX] = LLONG_MIN;
        for (int32_t i : run.instr) {
            const int64_t *a = &acc[(size_t)i * A_COUNT];
            for (int k = 0; k < A_COUNT; k++) {
                if (k == A_TMIN || k == A_ETMIN) sum[k] = std::min(sum[k], a[k]);
                else if (k == A_TMAX || k == A_ETMAX || k == A_PTMAX) sum[k] = std::max(sum[k], a[k]);
                else sum[k] += a[k];
            }
        }
        n_pe += sum[A_NPH] + sum[A_NDPE];
        rsum[r].row = !(sum[A_NPHALL] == 0 && run.type != 1 && run.type != 2);   // rawdata.py:336-337
        n_truth_here += rsum[r].row ? 1 : 0;
    }
    // ---- output ranges, in batch order ----
    wfs_counts *cn = so.counts;
    int64_t rec0 = 0, truth0 = 0, groups0 = 0;
    bool fits = false;
    {
        std::unique_lock<std::mutex> lk(ord.mu);
        ord.cv.wait(lk, [&] { return ord.abort || ord.out_done == batch_index; });
        if (ord.abort) throw LaneAborted();
        rec0 = so.n_rec; truth0 = so.n_truth; groups0 = so.n_groups;
        const int64_t user_room = so.resident ? (int64_t(1) << 40)
                                               : (out && out->records && !so.overflow ? out->cap_records - so.n_rec : 0);
        fits = res.n_records <= cap_here && res.n_records <= user_room;
        if (!so.resident && !fits) so.overflow = true;
        so.n_rec += res.n_records;
        so.n_truth += n_truth_here;
        so.n_groups += ngroups;
        so.n_batches = std::max(so.n_batches, batch_index + 1);
        for (int k = 0; k < 3; k++) cn->n_records[k] += fits ? res.n_rec_class[k] : 0;
        if (!so.resident && fits)
            cn->d2h_bytes += compact ? (int64_t)sizeof(CompactHdr) * (res.n_records - res.n_plain) + kBlockBytes * res.n_blocks +
                                           (int64_t)WFS_RECORD_BYTES * res.n_plain
                                     : (int64_t)WFS_RECORD_BYTES * res.n_records;
        cn->n_pe += n_pe;
        cn->n_photons += res.n_valid_photons;
        cn->n_pulses += res.n_pulses;
        cn->n_windows += res.n_windows;
        cn->n_intervals += res.n_intervals;
        cn->n_samples += res.n_samples;
        cn->n_pulse_calls += nruns;
        cn->n_instructions += ntot;
        cn->ms_digitize += res.ms_digitize;
        cn->ms_phase[0] += ms_front;
        cn->ms_phase[7] += ms_host;
        for (int k = 1; k <= 6; k++) cn->ms_phase[k] += res.ms_phase[k];
        cn->n_fused_batches += res.fused;
        cn->ms_phase[10] += res.segment_sorted_photons;
        cn->ms_phase[11] += res.segment_sorted_records;
        ord.out_done = batch_index + 1;
    }
    ord.cv.notify_all();
    if (!so.resident && fits && res.n_records > 0) {
        WFS_CUDA_CHECK(cudaEventRecord(F.ev_ready, s));
        WFS_CUDA_CHECK(cudaStreamWaitEvent(L.copy_stream, F.ev_ready, 0));
        // in pieces: the copy engine serves streams in FIFO order, and the other lane's 4-byte count
        // readbacks must not queue behind a gigabyte of records
        const size_t total_bytes = (size_t)res.n_records * WFS_RECORD_BYTES, piece = size_t(8) << 20;
        uint8_t *dst = out->records + (size_t)rec0 * WFS_RECORD_BYTES;
        if (compact) {
            // the compact streams first (the expansion starts as soon as they have arrived), the plain rows behind them
            const int64_t n_plain = res.n_plain, n_comp = res.n_records - n_plain;
            const bool both = n_plain > 0 && n_comp > 0;
            if (n_comp > 0)
                cs.ship(H->host_pool(), L.copy_stream, n_comp, res.n_blocks, dst + (size_t)n_plain * WFS_RECORD_BYTES,
                        H->record_fill(), (int16_t)p.dt, &H->tstats, both ? &H->split : nullptr);
            if (n_plain > 0) {
                const size_t plain_bytes = (size_t)n_plain * WFS_RECORD_BYTES;
                for (size_t o = 0; o < plain_bytes; o += piece)
                    WFS_CUDA_CHECK(cudaMemcpyAsync(dst + o, d_rec + o, std::min(piece, plain_bytes - o),
                                                   cudaMemcpyDeviceToHost, L.copy_stream));
                if (both) WFS_CUDA_CHECK(cudaLaunchHostFunc(L.copy_stream, ExpandJob::plain_done_callback, &cs.job));
                WFS_CUDA_CHECK(cudaEventRecord(F.ev_copy[par], L.copy_stream));
                F.copy_pending[par] = true;
                H->tstats.n_plain += n_plain;
            }
        } else {
            for (size_t o = 0; o < total_bytes; o += piece)
                WFS_CUDA_CHECK(cudaMemcpyAsync(dst + o, d_rec + o, std::min(piece, total_bytes - o),
                                               cudaMemcpyDeviceToHost, L.copy_stream));
            WFS_CUDA_CHECK(cudaEventRecord(F.ev_copy[par], L.copy_stream));
            F.copy_pending[par] = true;
        }
    }
    if (out && out->groups)
        for (int32_t gi = 0; gi < ngroups; gi++)
            if (groups0 + gi < out->cap_groups) out->groups[groups0 + gi] = h_groups[gi];
    if (out && out->batch_records && batch_index < out->cap_batches)
        for (int k = 0; k < 3; k++) out->batch_records[3 * batch_index + k] = fits ? res.n_rec_class[k] : 0;
    int64_t trow = truth0;
    for (int64_t r = 0; r < nruns; r++) {
        if (!rsum[r].row) continue;
        const Run &run = runs[r];
        if (out && out->truth && trow < out->cap_truth) {
            long double pm, psig, em, esig;
            combine_moments(run.instr, T, acc.data(), A_NPHALL, A_SREL, A_SHI2, A_SHILO, A_SLO2, pm, psig);
            combine_moments(run.instr, T, acc.data(), A_NE, A_ESREL, A_EHI2, A_EHILO, A_ELO2, em, esig);
            const int32_t i0 = run.instr[0];
            const int32_t root = i0 < nprim ? i0 : in.parent[i0];
            const HostInstr &h0 = P.instr[P.order[j0 + root]];
            float x, y, z; int32_t amp;
            auto fx = [&](int32_t i) { return i < nprim ? h_x[i] : sec_x[i - nprim]; };
            auto fy = [&](int32_t i) { return i < nprim ? h_y[i] : sec_y[i - nprim]; };
            auto fz = [&](int32_t i) { return i < nprim ? h_z[i] : sec_z[i - nprim]; };
            auto fa = [&](int32_t i) { return i < nprim ? h_amp[i] : sec_amp[i - nprim]; };
            if (run.instr.size() > 1) {
                std::vector<float> vx, vy, vz;
                int64_t sa = 0;
                for (int32_t i : run.instr) { vx.push_back(fx(i)); vy.push_back(fy(i)); vz.push_back(fz(i)); sa += fa(i); }
                const float nn = (float)run.instr.size();
                x = np_pairwise_sum_f32(vx.data(), vx.size()) / nn;
                y = np_pairwise_sum_f32(vy.data(), vy.size()) / nn;
                z = np_pairwise_sum_f32(vz.data(), vz.size()) / nn;
                amp = (int32_t)sa;
            } else {
                x = fx(i0); y = fy(i0); z = fz(i0); amp = fa(i0);
            }
            // mean observed position of the S2 call under a field-distortion model (rawdata.py:377-390):
            // the same model output the electrons were drifted with (wfs_instr_maps.x_obs / y_obs)
            float xme = std::numeric_limits<float>::quiet_NaN(), yme = xme;
            if (run.type == 2 && !P.xo.empty()) {
                double sx = 0, sy = 0;
                for (int32_t i : run.instr) { sx += P.xo[P.order[j0 + i]]; sy += P.yo[P.order[j0 + i]]; }
                xme = (float)(sx / (double)run.instr.size());
                yme = (float)(sy / (double)run.instr.size());
            }
            write_truth_row(out->truth + (size_t)trow * WFS_TRUTH_BYTES, h0, run.type, T[i0], x, y, z,
                            amp, rsum[r].sum, pm, psig, em, esig, trig[4 * r], trig[4 * r + 1], p, xme, yme);
            if (per_pmt) {
                memcpy(out->truth_pmt_counts + (size_t)trow * 4 * n_ch, &pmt_cnt[(size_t)r * 4 * n_ch],
                       sizeof(int32_t) * 4 * (size_t)n_ch);
                double *dst = out->truth_pmt_areas + (size_t)trow * 2 * n_ch;
                const int64_t *src = &pmt_area[(size_t)r * 2 * n_ch];
                for (int k = 0; k < 2 * n_ch; k++) dst[k] = (double)src[k] / kAreaScale;
            }
        }
        trow++;
    }
    L.local_batches++;
}

// Lane 0 is the handle's own stream / back end / front end; further lanes get their own stream and
// workspaces and share the device tables.
static void clone_tables(const Frontend &a, Frontend &b) {
    b.spe_ppf = a.spe_ppf; b.spe_row = a.spe_row; b.n_spe_rows = a.n_spe_rows; b.spe_len = a.spe_len;
    b.lum_cdf = a.lum_cdf; b.lum_t = a.lum_t; b.lum_len = a.lum_len; b.lum_guide = a.lum_guide;
    b.n_ap = a.n_ap;
    for (int e = 0; e < WFS_MAX_AP_ELEMENTS; e++) {
        b.ap_is_uniform[e] = a.ap_is_uniform[e];
        b.ap_delay_cdf[e] = a.ap_delay_cdf[e]; b.ap_delay_len[e] = a.ap_delay_len[e];
        b.ap_delay_bin[e] = a.ap_delay_bin[e];
        b.ap_amp_cdf[e] = a.ap_amp_cdf[e]; b.ap_amp_len[e] = a.ap_amp_len[e];
        b.ap_amp_rows[e] = a.ap_amp_rows[e]; b.ap_amp_bin[e] = a.ap_amp_bin[e];
    }
    b.pi_coarse_time = a.pi_coarse_time; b.pi_coarse_prob = a.pi_coarse_prob; b.pi_coarse_len = a.pi_coarse_len;
    b.h_pi_coarse_time = a.h_pi_coarse_time;
    b.s1_op_top = a.s1_op_top; b.s1_op_bottom = a.s1_op_bottom; b.s2_op_top = a.s2_op_top; b.s2_op_bottom = a.s2_op_bottom;
    b.s1_op_nz = a.s1_op_nz; b.s1_op_nu = a.s1_op_nu; b.s2_op_nu = a.s2_op_nu;
    b.s1_op_z0 = a.s1_op_z0; b.s1_op_z1 = a.s1_op_z1; b.s1_op_u0 = a.s1_op_u0; b.s1_op_u1 = a.s1_op_u1;
    b.s2_op_u0 = a.s2_op_u0; b.s2_op_u1 = a.s2_op_u1;
    b.gf_t = a.gf_t; b.gf_x = a.gf_x; b.gf_rows = a.gf_rows; b.gf_cols = a.gf_cols;
    b.gg_cdf = a.gg_cdf; b.gg_rows = a.gg_rows; b.gg_len = a.gg_len;
    b.lumw_alpha = a.lumw_alpha; b.lumw_ue = a.lumw_ue; b.lumw_pressure = a.lumw_pressure;
    b.lumw_ra = a.lumw_ra; b.lumw_rw = a.lumw_rw; b.lumw_dr = a.lumw_dr;
    b.s1_pat = a.s1_pat; b.s2_pat = a.s2_pat;
}

static void init_copy_events(Frontend &F) {
    if (F.ev_ready) return;
    WFS_CUDA_CHECK(cudaEventCreateWithFlags(&F.ev_ready, cudaEventDisableTiming));
    WFS_CUDA_CHECK(cudaEventCreateWithFlags(&F.ev_copy[0], cudaEventDisableTiming));
    WFS_CUDA_CHECK(cudaEventCreateWithFlags(&F.ev_copy[1], cudaEventDisableTiming));
}

static void ensure_lanes(Handle *H, int n) {
    while ((int)H->lanes.size() < n) {
        Lane *L = new Lane();
        L->id = (int)H->lanes.size();
        if (L->id == 0) {
            L->stream = H->stream; L->copy_stream = H->copy_stream;
            L->F = H->frontend; L->B = H->backend;
        } else {
            WFS_CUDA_CHECK(cudaStreamCreateWithFlags(&L->stream, cudaStreamNonBlocking));
            WFS_CUDA_CHECK(cudaStreamCreateWithFlags(&L->copy_stream, cudaStreamNonBlocking));
            L->F = new Frontend(H);
            clone_tables(*H->frontend, *L->F);
            L->F->prim.stream = L->stream;
            L->F->prim.lc = &H->launches;
            L->B = new Backend(&H->cfg, L->stream, &H->launches);
        }
        WFS_CUDA_CHECK(cudaEventCreate(&L->ev_c));
        WFS_CUDA_CHECK(cudaEventCreate(&L->ev_d));
        WFS_CUDA_CHECK(cudaEventCreateWithFlags(&L->ev_done, cudaEventDisableTiming));
        init_copy_events(*L->F);
        H->lanes.push_back(L);
    }
}

static void release_frontend_buffers(Frontend &F) {
    DevBuf *all[] = {&F.b_itype, &F.b_itime, &F.b_ix, &F.b_iy, &F.b_iz, &F.b_iamp, &F.b_igidx, &F.b_ilce,
                     &F.b_iscg, &F.b_icy, &F.b_ipat, &F.b_ivd, &F.b_idl, &F.b_ixo, &F.b_iyo, &F.b_irecoil, &F.b_ilrow, &F.b_ioptfirst, &F.b_ioptn,
                     &F.b_igglo, &F.b_igghi, &F.b_iggfrac, &F.b_iggmean, &F.b_ggpartial, &F.b_dmean, &F.b_dspread, &F.b_nemit, &F.b_emitoff,
                     &F.b_nhits, &F.b_acc, &F.b_titems, &F.b_cdf, &F.b_cdfok, &F.b_cdfguide, &F.b_pattern, &F.b_et, &F.b_einstr,
                     &F.b_enph, &F.b_ephoff, &F.b_pht, &F.b_phch, &F.b_phgain, &F.b_phinstr, &F.b_phflags,
                     &F.b_phnap, &F.b_apoff, &F.b_picount, &F.b_pioff, &F.b_pecount, &F.b_peoff, &F.b_irun, &F.b_pcgroup,
                     &F.b_pcrank, &F.b_trig, &F.b_records, &F.b_records2, &F.b_groups, &F.b_scal, &F.b_phstart,
                     &F.b_gstart, &F.b_gt0, &F.b_grun0, &F.b_pmtcnt, &F.b_pmtarea};
    for (DevBuf *b : all) b->release();
    for (CompactStage &cs : F.cstage) cs.release();
    F.cdf_rows = -1;
    F.prim.release();
    if (F.ev_ready) {
        cudaEventDestroy(F.ev_ready); cudaEventDestroy(F.ev_copy[0]); cudaEventDestroy(F.ev_copy[1]);
        F.ev_ready = nullptr;
    }
}

void Handle::lanes_release() {
    for (Lane *L : lanes) {
        cudaEventDestroy(L->ev_c); cudaEventDestroy(L->ev_d); cudaEventDestroy(L->ev_done);
        if (L->id != 0) {
            release_frontend_buffers(*L->F);
            delete L->F;
            delete L->B;
            cudaStreamDestroy(L->stream);
            cudaStreamDestroy(L->copy_stream);
        }
        delete L;
    }
    lanes.clear();
}

static int run_plan(Handle *H, Plan &P, uint64_t seed, wfs_outputs *out, wfs_counts *counts, bool resident,
                    PhotonDump *dump = nullptr) {
    if (!H->frontend) throw std::runtime_error("front-end tables missing (SPE table is required for wfs_simulate)");
    memset(counts, 0, sizeof(*counts));
    if (P.n_opt > 0) {      // the caller's photon lists, resident for the whole call (all lanes read them)
        H->d_opt_ch.reserve(sizeof(int32_t) * (size_t)P.n_opt);
        H->d_opt_t.reserve(sizeof(int64_t) * (size_t)P.n_opt);
        WFS_CUDA_CHECK(cudaMemcpy(H->d_opt_ch.p, P.opt_channels, sizeof(int32_t) * (size_t)P.n_opt, cudaMemcpyHostToDevice));
        WFS_CUDA_CHECK(cudaMemcpy(H->d_opt_t.p, P.opt_timings, sizeof(int64_t) * (size_t)P.n_opt, cudaMemcpyHostToDevice));
        H->opt_cutoff = P.opt_cutoff;
    }
    const int64_t launches0 = H->launches.n;
    H->tstats.ns_copy = 0;
    H->tstats.ns_expand = 0;
    H->tstats.n_plain = 0;
    SimOut so;
    so.out = out;
    so.counts = counts;
    so.resident = resident;
    so.dump = dump;
    so.compact = !resident && out && out->records && H->use_compact(out->records);
    so.split = so.compact && H->compact_mode == 1 && Handle::page_locked(out->records);
    if (so.split) H->split.init(H->host_pool()->size());
    const int64_t nb = (int64_t)P.batches.size();
    // lanes: host threads that each drive every n-th device batch on their own streams.  Four keep the GPU busy when
    // the rank has the cores for them next to the record expanders; three otherwise.  (Five: 44.2 against 45.9 ms per
    // 1e5 C1 events with the records staying on the device, but 390 against 379 ms end to end, and the C3 sample that
    // follows C1 and C2 in one process fell from 363 to 527 ms per step -- not taken.)
    const int n_lanes = dump ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(env_i64("WFS_LANES", default_lanes()), 8), nb));
    ensure_lanes(H, std::max(n_lanes, 1));
    cudaStream_t s = H->stream;
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_a, s));
    Order ord;
    std::exception_ptr failure[8] = {};
    auto lane_loop = [&](int li) {
        try {
            WFS_CUDA_CHECK(cudaSetDevice(H->device));
            for (int64_t k = li; k < nb; k += n_lanes) simulate_batch(H, *H->lanes[li], P, k, P.batches[k], seed, so, ord);
        } catch (const LaneAborted &) {
        } catch (...) {
            failure[li] = std::current_exception();
            {
                std::lock_guard<std::mutex> lk(ord.mu);
                ord.abort = true;
            }
            ord.cv.notify_all();
        }
    };
    {
        std::vector<std::thread> workers;
        for (int li = 1; li < n_lanes; li++) workers.emplace_back(lane_loop, li);
        lane_loop(0);
        for (auto &t : workers) t.join();
    }
    for (int li = 0; li < n_lanes; li++)
        if (failure[li]) {
            for (int lj = 0; lj < n_lanes; lj++) {   // drain before reporting
                stream_sync(H->lanes[lj]->stream);
                const bool copies_ok = stream_sync(H->lanes[lj]->copy_stream) == cudaSuccess;
                H->lanes[lj]->F->copy_pending[0] = H->lanes[lj]->F->copy_pending[1] = false;
                for (CompactStage &cs : H->lanes[lj]->F->cstage) {
                    if (copies_ok) cs.job.wait(); else cs.job.abandon();
                }
            }
            std::rethrow_exception(failure[li]);
        }
    for (int li = 1; li < n_lanes; li++) {   // the closing event on lane 0 covers every lane
        WFS_CUDA_CHECK(cudaEventRecord(H->lanes[li]->ev_done, H->lanes[li]->stream));
        WFS_CUDA_CHECK(cudaStreamWaitEvent(s, H->lanes[li]->ev_done, 0));
    }
    WFS_CUDA_CHECK(cudaEventRecord(H->ev_b, s));
    for (int li = 0; li < n_lanes; li++) {
        WFS_CUDA_CHECK(stream_sync(H->lanes[li]->stream));
        WFS_CUDA_CHECK(stream_sync(H->lanes[li]->copy_stream));
        H->lanes[li]->F->copy_pending[0] = H->lanes[li]->F->copy_pending[1] = false;
        for (CompactStage &cs : H->lanes[li]->F->cstage) cs.job.wait();   // expansion into the caller's array
    }
    float ms;
    WFS_CUDA_CHECK(cudaEventElapsedTime(&ms, H->ev_a, H->ev_b));
    counts->ms_total = ms;
    counts->ms_phase[8] = H->tstats.ns_copy.load() * 1e-6;
    counts->ms_phase[9] = H->tstats.ns_expand.load() * 1e-6;
    counts->n_plain_records = H->tstats.n_plain.load();
    counts->n_records_total = so.n_rec;
    counts->n_truth = so.n_truth;
    counts->n_groups = so.n_groups;
    counts->n_batches = so.n_batches;
    counts->need_records = so.n_rec;
    counts->need_truth = so.n_truth;
    counts->need_groups = so.n_groups;
    counts->need_batches = so.n_batches;
    counts->gpu_launches = H->launches.n - launches0;
    bool cap = so.overflow;
    if (out) {
        if (out->truth && so.n_truth > out->cap_truth) cap = true;
        if (out->groups && so.n_groups > out->cap_groups) cap = true;
        if (out->batch_records && so.n_batches > out->cap_batches) cap = true;
    }
    if (cap) {
        counts->n_records_total = 0;
        for (int k = 0; k < 3; k++) counts->n_records[k] = 0;
        return WFS_E_CAPACITY;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
template <typename T>
static T *upload_table(const T *host, size_t count, std::vector<void *> &owned) {
    if (!host || count == 0) return nullptr;
    T *d = nullptr;
    WFS_CUDA_CHECK(cudaMalloc((void **)&d, count * sizeof(T)));
    WFS_CUDA_CHECK(cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice));
    owned.push_back(d);
    return d;
}

void Handle::frontend_init(const wfs_tables &t) {
    frontend = nullptr;
    if (!t.spe_ppf || !t.spe_row) return;   // deterministic entry only
    Frontend *F = new Frontend(this);
    F->prim.stream = stream;
    F->prim.lc = &launches;
    const wfs_params &p = cfg.p;
    F->n_spe_rows = t.n_spe_rows; F->spe_len = t.spe_len;
    F->spe_ppf = upload_table(t.spe_ppf, (size_t)t.n_spe_rows * t.spe_len, owned);
    F->spe_row = upload_table(t.spe_row, (size_t)p.n_tpc_pmts, owned);
    if (t.lum_cdf && t.lum_t && t.lum_len > 1) {
        F->lum_cdf = upload_table(t.lum_cdf, (size_t)t.lum_len, owned);
        F->lum_t = upload_table(t.lum_t, (size_t)t.lum_len, owned);
        F->lum_len = t.lum_len;
        // guide[c] = first entry above c / kLumGuide: u in [c, c + 1) / kLumGuide is searched between guide[c] and guide[c + 1]
        std::vector<uint32_t> guide(kLumGuide + 1);
        for (int c = 0; c <= kLumGuide; c++)
            guide[c] = (uint32_t)(std::upper_bound(t.lum_cdf, t.lum_cdf + t.lum_len, (double)c / (double)kLumGuide) - t.lum_cdf);
        F->lum_guide = upload_table(guide.data(), guide.size(), owned);
    }
    F->n_ap = std::min<int32_t>(t.n_ap_elements, WFS_MAX_AP_ELEMENTS);
    for (int e = 0; e < WFS_MAX_AP_ELEMENTS; e++) {
        F->ap_is_uniform[e] = 0; F->ap_delay_cdf[e] = nullptr; F->ap_amp_cdf[e] = nullptr;
        F->ap_delay_len[e] = F->ap_amp_len[e] = F->ap_amp_rows[e] = 0;
        F->ap_delay_bin[e] = F->ap_amp_bin[e] = 0;
    }
    F->ap_max_delay_ns = 0.0;
    for (int e = 0; e < F->n_ap; e++) {
        if (t.ap_delay_cdf[e] && t.ap_delay_len[e] > 0) {
            // longest delay the element can draw: the last bin of the cdf, or the upper edge of a 'Uniform' row
            // ([t_low_bin, t_high_bin, P], afterpulse.py:192,214-215)
            double reach = (double)t.ap_delay_len[e];
            if (t.ap_is_uniform[e] && t.ap_delay_len[e] >= 2) {
                reach = 0.0;
                for (int ch = 0; ch < p.n_tpc_pmts; ch++)
                    reach = std::max(reach, t.ap_delay_cdf[e][(size_t)ch * t.ap_delay_len[e] + 1]);
            }
            F->ap_max_delay_ns = std::max(F->ap_max_delay_ns, reach * t.ap_delay_bin[e]);
        }
        F->ap_is_uniform[e] = t.ap_is_uniform[e];
        F->ap_delay_len[e] = t.ap_delay_len[e];
        F->ap_delay_bin[e] = t.ap_delay_bin[e];
        F->ap_delay_cdf[e] = upload_table(t.ap_delay_cdf[e], (size_t)p.n_tpc_pmts * t.ap_delay_len[e], owned);
        F->ap_amp_len[e] = t.ap_amp_len[e];
        F->ap_amp_rows[e] = t.ap_amp_rows[e];
        F->ap_amp_bin[e] = t.ap_amp_bin[e];
        if (t.ap_amp_cdf[e])
            F->ap_amp_cdf[e] = upload_table(t.ap_amp_cdf[e], (size_t)std::max(1, t.ap_amp_rows[e]) * t.ap_amp_len[e], owned);
    }
    if (t.pi_coarse_time && t.pi_coarse_prob && t.pi_coarse_len > 0) {
        F->pi_coarse_len = t.pi_coarse_len;
        F->pi_coarse_time = upload_table(t.pi_coarse_time, (size_t)t.pi_coarse_len, owned);
        F->pi_coarse_prob = upload_table(t.pi_coarse_prob, (size_t)t.pi_coarse_len, owned);
        F->h_pi_coarse_time.assign(t.pi_coarse_time, t.pi_coarse_time + t.pi_coarse_len);
    }
    if (p.s1_model_optical) {
        if (!t.s1_op_top || !t.s1_op_bottom || t.s1_op_nz < 2 || t.s1_op_nu < 2)
            throw std::runtime_error("s1_model_type contains optical_propagation but no s1_optical_propagation_spline grid was given");
        F->s1_op_top = upload_table(t.s1_op_top, (size_t)t.s1_op_nz * t.s1_op_nu, owned);
        F->s1_op_bottom = upload_table(t.s1_op_bottom, (size_t)t.s1_op_nz * t.s1_op_nu, owned);
        F->s1_op_nz = t.s1_op_nz; F->s1_op_nu = t.s1_op_nu;
        F->s1_op_z0 = t.s1_op_z0; F->s1_op_z1 = t.s1_op_z1; F->s1_op_u0 = t.s1_op_u0; F->s1_op_u1 = t.s1_op_u1;
    }
    if (p.s2_time_model == 2) {
        if (!t.s2_op_top || !t.s2_op_bottom || t.s2_op_nu < 2)
            throw std::runtime_error("s2_time_model is optical_propagation but no s2_optical_propagation_spline grid was given");
        F->s2_op_top = upload_table(t.s2_op_top, (size_t)t.s2_op_nu, owned);
        F->s2_op_bottom = upload_table(t.s2_op_bottom, (size_t)t.s2_op_nu, owned);
        F->s2_op_nu = t.s2_op_nu; F->s2_op_u0 = t.s2_op_u0; F->s2_op_u1 = t.s2_op_u1;
    }
    if (t.s1_pat_grid) {
        if (t.s1_pat_npmt != p.n_tpc_pmts || t.s1_pat_n[0] < 2 || t.s1_pat_n[1] < 2 || t.s1_pat_n[2] < 2)
            throw std::runtime_error("s1 pattern grid: needs >= 2 points per axis and n_tpc_pmts columns");
        F->s1_pat.nd = 3; F->s1_pat.npmt = t.s1_pat_npmt;
        size_t cells = 1;
        for (int d = 0; d < 3; d++) { F->s1_pat.n[d] = t.s1_pat_n[d]; F->s1_pat.lo[d] = t.s1_pat_lo[d]; F->s1_pat.hi[d] = t.s1_pat_hi[d]; cells *= (size_t)t.s1_pat_n[d]; }
        F->s1_pat.v = upload_table(t.s1_pat_grid, cells * (size_t)t.s1_pat_npmt, owned);
    }
    if (t.s2_pat_grid) {
        if (t.s2_pat_npmt < 1 || t.s2_pat_npmt > p.n_tpc_pmts || t.s2_pat_n[0] < 2 || t.s2_pat_n[1] < 2)
            throw std::runtime_error("s2 pattern grid: needs >= 2 points per axis and at most n_tpc_pmts columns");
        F->s2_pat.nd = 2; F->s2_pat.npmt = t.s2_pat_npmt;
        size_t cells = 1;
        for (int d = 0; d < 2; d++) { F->s2_pat.n[d] = t.s2_pat_n[d]; F->s2_pat.lo[d] = t.s2_pat_lo[d]; F->s2_pat.hi[d] = t.s2_pat_hi[d]; cells *= (size_t)t.s2_pat_n[d]; }
        F->s2_pat.v = upload_table(t.s2_pat_grid, cells * (size_t)t.s2_pat_npmt, owned);
    }
    if (p.s2_luminescence_model == 1) {
        if (!t.gf_t || !t.gf_x || t.gf_rows < 1 || t.gf_cols < 1)
            throw std::runtime_error("s2_luminescence model not found");   // s2.py:391
        F->gf_t = upload_table(t.gf_t, (size_t)t.gf_rows * t.gf_cols, owned);
        F->gf_x = upload_table(t.gf_x, (size_t)t.gf_rows, owned);
        F->gf_rows = t.gf_rows; F->gf_cols = t.gf_cols;
    } else if (p.s2_luminescence_model == 2) {
        if (!t.gg_cdf || t.gg_rows < 1 || t.gg_len < 3)
            throw std::runtime_error("s2_luminescence_gg model not found");   // s2.py:471
        F->gg_cdf = upload_table(t.gg_cdf, (size_t)t.gg_rows * t.gg_len, owned);
        F->gg_rows = t.gg_rows; F->gg_len = t.gg_len;
    } else if (p.s2_luminescence_model == 0 && F->lum_len <= 0) {
        // per-position gas gaps (enable_gas_gap_warping): the instructions carry lum_gap / lum_e0
        if (!(t.lumw_dr > 0.0 && t.lumw_alpha > 0.0 && t.lumw_ra > t.lumw_rw && t.lumw_rw > 0.0))
            throw std::runtime_error("s2_luminescence_model 'simple': neither the constant-gap table (lum_cdf) nor the field scalars (lumw_*) were given");
    }
    F->lumw_alpha = t.lumw_alpha; F->lumw_ue = t.lumw_ue; F->lumw_pressure = t.lumw_pressure;
    F->lumw_ra = t.lumw_ra; F->lumw_rw = t.lumw_rw; F->lumw_dr = t.lumw_dr;
    frontend = F;
}

void Handle::frontend_release() {
    lanes_release();
    if (!frontend) return;
    release_frontend_buffers(*frontend);
    delete reinterpret_cast<Plan *>(staged_plan);
    staged_plan = nullptr;
    delete frontend;
    frontend = nullptr;
}

}  // namespace wfs

using namespace wfs;

#define API_TRY(h)                                   \
    Handle *H = reinterpret_cast<Handle *>(h);       \
    if (!H) return WFS_E_ARG;                        \
    try {                                            \
        WFS_CUDA_CHECK(cudaSetDevice(H->device));

#define API_CATCH                                                          \
    } catch (const std::exception &e) {                                    \
        H->last_error = e.what();                                          \
        cudaGetLastError();                                                \
        if (H->last_error == "Pulse cache too long") return WFS_E_PULSE_CACHE_TOO_LONG; \
        return H->last_error.find("CUDA") != std::string::npos ? WFS_E_CUDA : WFS_E_ARG; \
    }

extern "C" {

int64_t wfs_quiet_gap(void *handle) {
    Handle *H = reinterpret_cast<Handle *>(handle);
    return H ? (int64_t)std::ceil(quiet_gap_ns(H)) : -1;
}

int wfs_schedule(int64_t right_raw_extension, double drift_velocity_liquid, int save_full_truth,
                 int64_t n_prim, const int64_t *time, const float *z, const int8_t *type,
                 int64_t n_sec, const int64_t *sec_time, const float *sec_z, const int8_t *sec_type,
                 const int32_t *sec_parent, const int64_t *pulse_end, int32_t *run_of, int32_t *run_type,
                 int32_t *run_group, int64_t cap_runs, int64_t *n_runs, int64_t *n_groups) {
    try {
        if (n_prim < 0 || n_sec < 0 || !n_runs || !n_groups) return WFS_E_ARG;
        const int64_t n_tot = n_prim + n_sec;
        // primaries in signal-time order, clustered at gaps > rext: what make_plan hands to a device batch
        std::vector<int64_t> st((size_t)n_prim), order((size_t)n_prim), where((size_t)n_prim);
        for (int64_t i = 0; i < n_prim; i++) {
            st[i] = signal_time(time[i], z[i], type[i], drift_velocity_liquid);
            order[i] = i;
        }
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return st[a] < st[b]; });
        for (int64_t j = 0; j < n_prim; j++) where[order[j]] = j;
        SchedIn in;
        in.n_prim = n_prim;
        in.rext = right_raw_extension;
        in.save_full_truth = save_full_truth != 0;
        in.v = drift_velocity_liquid;
        in.stime.resize(n_tot); in.type.resize(n_tot); in.parent.assign(n_tot, -1); in.pend.resize(n_tot);
        for (int64_t j = 0; j < n_prim; j++) {
            in.stime[j] = st[order[j]];
            in.type[j] = type[order[j]];
            in.pend[j] = pulse_end[order[j]];
            if (j == 0 || in.stime[j] - in.stime[j - 1] > in.rext) in.clusters.push_back({(int32_t)j, (int32_t)j});
            in.clusters.back().second = (int32_t)(j + 1);
        }
        for (int64_t k = 0; k < n_sec; k++) {
            if (sec_parent[k] < 0 || sec_parent[k] >= n_prim) return WFS_E_ARG;
            in.stime[n_prim + k] = signal_time(sec_time[k], sec_z[k], sec_type[k], drift_velocity_liquid);
            in.type[n_prim + k] = sec_type[k];
            in.parent[n_prim + k] = (int32_t)where[sec_parent[k]];
            in.pend[n_prim + k] = pulse_end[n_prim + k];
        }
        std::vector<Run> runs;
        int32_t ngroups = 0, nreal = 0;
        schedule(in, runs, ngroups, &nreal);
        *n_runs = (int64_t)runs.size();
        *n_groups = nreal;
        if ((int64_t)runs.size() > cap_runs) return WFS_E_CAPACITY;
        for (int64_t i = 0; i < n_tot; i++) run_of[i] = -1;
        for (size_t r = 0; r < runs.size(); r++) {
            run_type[r] = runs[r].type;
            run_group[r] = runs[r].group;
            for (int32_t i : runs[r].instr) run_of[i < n_prim ? order[i] : i] = (int32_t)r;
        }
        return 0;
    } catch (const std::exception &) {
        return WFS_E_ARG;
    }
}

int wfs_simulate(void *handle, const uint8_t *instructions, int64_t n_instructions,
                 const wfs_instr_maps *maps, uint64_t seed, wfs_outputs *out, wfs_counts *counts) {
    API_TRY(handle)
    if (!counts) throw std::runtime_error("counts is required");
    if (n_instructions < 0 || (n_instructions > 0 && !instructions)) throw std::runtime_error("bad instructions");
    Plan P;
    make_plan(H, instructions, n_instructions, maps, P);
    return run_plan(H, P, seed, out, counts, false);
    API_CATCH
}

int wfs_stage_instructions(void *handle, const uint8_t *instructions, int64_t n_instructions,
                           const wfs_instr_maps *maps) {
    API_TRY(handle)
    if (!H->frontend) throw std::runtime_error("front-end tables missing");
    Plan *P = new Plan();
    make_plan(H, instructions, n_instructions, maps, *P);
    delete reinterpret_cast<Plan *>(H->staged_plan);
    H->staged_plan = P;
    return 0;
    API_CATCH
}

int wfs_run_staged(void *handle, uint64_t seed, wfs_outputs *out, wfs_counts *counts) {
    API_TRY(handle)
    if (!H->staged_plan) throw std::runtime_error("nothing staged");
    if (!counts) throw std::runtime_error("counts is required");
    return run_plan(H, *reinterpret_cast<Plan *>(H->staged_plan), seed, out, counts, true);
    API_CATCH
}

int wfs_sample_stage(void *handle, int stage, const uint8_t *instructions, int64_t n_instructions,
                     const wfs_instr_maps *maps, uint64_t seed, void *out, int64_t cap, int64_t *n_out) {
    API_TRY(handle)
    if (!n_out) throw std::runtime_error("n_out is required");
    Plan P;
    make_plan(H, instructions, n_instructions, maps, P);
    PhotonDump d;
    d.stage = stage;
    d.out = reinterpret_cast<uint8_t *>(out);
    d.cap = out ? cap : 0;
    wfs_counts counts;
    run_plan(H, P, seed, nullptr, &counts, false, &d);
    *n_out = d.n;
    return d.n > d.cap ? WFS_E_CAPACITY : 0;
    API_CATCH
}

}  // extern "C"
