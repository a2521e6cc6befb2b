// Sampling front end (Philox-driven S1/S2/afterpulse photon generation) + host scheduler.
#pragma once
#include "handle.cuh"

namespace wfs {

// Per-instruction truth accumulators (all int64, exact -> deterministic under atomics).
enum Acc {
    A_NPH = 0, A_NDPE, A_NTRIG, A_AREA, A_AREA_TRIG,             // pulse.py:259-266 (totals)
    A_NPH_B, A_NDPE_B, A_NTRIG_B, A_AREA_B, A_AREA_TRIG_B,       // ... bottom array
    A_NPHALL, A_TMIN, A_TMAX, A_SREL, A_SHI2, A_SHILO, A_SLO2,   // photon times, rawdata.py:325-332
    A_NE, A_ETMIN, A_ETMAX, A_ESREL, A_EHI2, A_EHILO, A_ELO2,    // electron times
    A_PTMAX,                                                     // last photon on a live PMT (incl. its PMT afterpulses)
    A_NAP,
    A_COUNT
};

constexpr double kAreaScale = 4294967296.0;   // fixed-point scale of the area accumulators

// Instruction as parsed from the 70-byte rows (host side)
struct HostInstr {
    int32_t event_number;
    int8_t type;
    int64_t time;
    float x, y, z;
    int32_t amp;
    int8_t recoil;
    float e_dep, tot_e;
    int32_t g4id, vol_id;
    double local_field;
    int32_t n_excitons;
    float x_pri, y_pri, z_pri;
};

// One Pulse call of the reference scheduler (rawdata.py:102-145)
struct Run {
    int32_t type;                    // 1, 2, 4, 6
    int32_t group;
    std::vector<int32_t> instr;      // batch-local instruction indices
};

struct DeviceInstr {   // SoA views, batch-local
    int32_t *type;
    int64_t *time;
    float *x, *y, *z;
    int32_t *amp;
    uint64_t *gidx;      // RNG identity of the instruction (independent of batching)
    double *lce, *scgain, *cyextra;
    int32_t *patrow;
    // derived
    double *dmean, *dspread;
    uint32_t *nemit;     // emitters (S1: 0/1, S2-like: electrons)
    uint32_t *emit_off;  // exclusive scan, [n+1]
    int64_t *nhits;      // S1 detected photons
    int64_t *acc;        // [A_COUNT][cap]
};

// Pattern map on a regular grid, resident on the device (wfs_tables.s1_pat_* / s2_pat_*)
struct PatGrid {
    const double *v = nullptr;     // [n0][n1]([n2])[npmt]
    int32_t nd = 0, npmt = 0;
    int32_t n[3] = {1, 1, 1};
    double lo[3] = {0, 0, 0}, hi[3] = {1, 1, 1};
};

struct Frontend {
    Handle *H;
    PatGrid s1_pat, s2_pat;
    // device tables
    double *spe_ppf = nullptr;
    int32_t *spe_row = nullptr;
    int32_t n_spe_rows = 0, spe_len = 0;
    double *lum_cdf = nullptr, *lum_t = nullptr;
    uint32_t *lum_guide = nullptr;      // guide of the search in lum_cdf (interp_table)
    int32_t lum_len = 0;
    int32_t n_ap = 0;
    int32_t ap_is_uniform[WFS_MAX_AP_ELEMENTS];
    double *ap_delay_cdf[WFS_MAX_AP_ELEMENTS];
    int32_t ap_delay_len[WFS_MAX_AP_ELEMENTS];
    double ap_delay_bin[WFS_MAX_AP_ELEMENTS];
    double ap_max_delay_ns = 0.0;      // longest PMT-afterpulse delay the tables can draw (quiet_gap_ns)
    double *ap_amp_cdf[WFS_MAX_AP_ELEMENTS];
    int32_t ap_amp_len[WFS_MAX_AP_ELEMENTS], ap_amp_rows[WFS_MAX_AP_ELEMENTS];
    double ap_amp_bin[WFS_MAX_AP_ELEMENTS];
    double *pi_coarse_time = nullptr, *pi_coarse_prob = nullptr;
    int32_t pi_coarse_len = 0;
    std::vector<double> h_pi_coarse_time;
    // optical propagation grids and the garfield luminescence table
    double *s1_op_top = nullptr, *s1_op_bottom = nullptr, *s2_op_top = nullptr, *s2_op_bottom = nullptr;
    int32_t s1_op_nz = 0, s1_op_nu = 0, s2_op_nu = 0;
    double s1_op_z0 = 0, s1_op_z1 = 1, s1_op_u0 = 0, s1_op_u1 = 1, s2_op_u0 = 0, s2_op_u1 = 1;
    int32_t *gf_t = nullptr;
    double *gf_x = nullptr;
    int32_t gf_rows = 0, gf_cols = 0;
    double *gg_cdf = nullptr;       // 'garfield_gas_gap' luminescence table
    int32_t gg_rows = 0, gg_len = 0;
    // 'simple' luminescence with per-position gas gaps (wfs_tables.lumw_*); dr == 0: not available
    double lumw_alpha = 0, lumw_ue = 0, lumw_pressure = 0, lumw_ra = 0, lumw_rw = 0, lumw_dr = 0;
    Primitives prim;
    // workspaces
    DevBuf b_itype, b_itime, b_ix, b_iy, b_iz, b_iamp, b_igidx, b_ilce, b_iscg, b_icy, b_ipat,
        b_ivd, b_idl, b_ixo, b_iyo, b_irecoil, b_ilrow, b_ioptfirst, b_ioptn, b_igglo, b_igghi, b_iggfrac, b_iggmean, b_ihsr, b_ihsa, b_ilgap, b_ilgapmax, b_ile0, b_ilavgt,
        b_ggpartial,
        b_dmean, b_dspread, b_nemit, b_emitoff, b_nhits, b_acc, b_titems, b_cdf, b_cdfok, b_cdfguide, b_pattern,
        b_et, b_einstr, b_enph, b_ephoff, b_pht, b_phch, b_phgain, b_phinstr, b_phflags, b_phnap,
        b_apoff, b_picount, b_pioff, b_pecount, b_peoff, b_irun, b_pcgroup, b_pcrank, b_trig, b_records, b_records2,
        b_groups, b_scal, b_phstart, b_gstart, b_gt0, b_grun0, b_pmtcnt, b_pmtarea;
    CompactStage cstage[2];     // compact record transport: batch k ships while batch k+1 runs
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_ready = nullptr;
    bool copy_pending[2] = {false, false};
    bool has_gg = false;        // the current call carries the per-instruction gas-gap rows
    bool has_opt = false;       // the current call supplies photons (wfs_instr_maps.opt_*)
    int64_t cdf_rows = -1;      // pattern CDF rows resident in b_cdf (few-row case only) and their hash
    uint64_t cdf_hash = 0;
    bool has_lw = false;                 // per-instruction gas gaps given (wfs_instr_maps.lum_gap / lum_e0)
    bool has_hd = false;                 // transverse-diffusion sigmas given (wfs_instr_maps.hdiff_sigma_*)
    int64_t first_dev_row = 0, n_pattern_rows = 0;   // pattern rows of the current batch: [0, first_dev_row) from the host
    bool has_vd = false, has_dl = false, has_xy = false;   // optional per-instruction arrays of the current batch
    int64_t *h_pin = nullptr;   // pinned scratch for small readbacks
    // staged instructions (device-resident measurement mode)
    std::vector<uint8_t> staged_instr;
    std::vector<double> staged_lce, staged_scg, staged_cy;
    std::vector<float> staged_pattern;
    std::vector<int32_t> staged_patrow;
    int64_t staged_n = 0, staged_pattern_rows = 0;
    bool staged_has[5] = {false, false, false, false, false};

    explicit Frontend(Handle *h) : H(h) {}
};

}  // namespace wfs
