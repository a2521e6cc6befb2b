/*
 * wfsim_b200 -- C-ABI of the B200-native WFSim hot path (wfsim_instructions -> strax raw_records).
 *
 * The reference (XENONnT/WFSim v1.2.2) is pure Python and has no FFI; the entry points below are
 * what a ctypes binding inside the unchanged `RawRecordsFromFaxNT` plugin binds instead of the
 * reference's Python call chain.  Each entry point cites the reference interface it replaces
 * (paths relative to the reference checkout).  Plain pointers and sizes only; the caller owns
 * every host buffer; the library owns device memory.  See INTEGRATION.md for the plugin-side stub.
 *
 * Return convention (all int-returning functions):
 *    0   success
 *   >0   WFS_E_CAPACITY: an output buffer is too small; `wfs_counts` holds the needed sizes,
 *        re-allocate and call again (results are reproducible: counter-based Philox RNG)
 *   <0   error; text via wfs_last_error()
 */
#ifndef WFSIM_B200_H
#define WFSIM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WFS_ABI_VERSION 5
#define WFS_E_CAPACITY 1
#define WFS_E_CUDA (-1)
#define WFS_E_ARG (-2)
#define WFS_E_PULSE_CACHE_TOO_LONG (-3) /* "Pulse cache too long", wfsim/core/rawdata.py:219 */
#define WFS_E_KEYBITS (-4)

#define WFS_RECORD_BYTES 244      /* strax.raw_record_dtype(110) */
#define WFS_SAMPLES_PER_RECORD 110
#define WFS_INSTRUCTION_BYTES 70  /* wfsim/strax_interface.py:25-42 */
#define WFS_TRUTH_BYTES 218       /* instruction_dtype + truth_extra_dtype, :49-73 */
#define WFS_MAX_AP_ELEMENTS 8

/* Scalar configuration: the fax_config keys the path reads (wfsim/core/*.py `self.config[...]`),
 * flattened on the host by wfsim_b200/params.py.  Units as in the reference (ns, cm, V). */
typedef struct wfs_params {
    int32_t abi_version;
    /* detector / channel layout: strax_interface.py:593-595, rawdata.py:241-254 */
    int32_t detector_nt;            /* config['detector'] == 'XENONnT' */
    int32_t n_tpc_pmts;
    int32_t n_top_pmts;
    int32_t he_first, he_last;      /* channel_map['he'] */
    int32_t he_mult;                /* int(high_energy_deamplification_factor) (sic) */
    int32_t n_rows;                 /* 801: rows of the reference's _raw_data */
    /* digitizer: pulse.py:121-127, rawdata.py:211-222,263-272 */
    int32_t dt;                     /* sample_duration */
    int32_t template_length;        /* samples_before_pulse_center + samples_after_pulse_center */
    int32_t pulse_left_margin;      /* samples_to_store_before + samples_before_pulse_center */
    int32_t pulse_right_margin;     /* samples_to_store_after + samples_after_pulse_center */
    int32_t trigger_window;
    int32_t baseline;               /* digitizer_reference_baseline */
    int32_t enable_noise;
    int32_t zle_threshold;          /* for truth trigger counters, pulse.py:240-243 */
    double current_2_adc;           /* pulse.py:33-35 */
    int64_t right_raw_extension;    /* strax_interface.py:531 */
    /* ---- sampling stages ---- */
    int32_t s1_model_simple;        /* 'simple' in s1_model_type, s1.py:191 */
    int32_t s1_model_optical;       /* 'optical_propagation' in s1_model_type, s1.py:186 (needs spline table) */
    int32_t s2_luminescence_model;  /* 0 simple, 1 garfield, 2 garfield_gas_gap, s2.py:518-536 */
    int32_t s2_time_model;          /* 0 's2_time_spread around zero', 1 zero_delay, 2 optical_propagation, s2.py:542-552 */
    int32_t enable_pmt_afterpulses; /* rawdata.py:176 */
    int32_t enable_electron_afterpulses; /* rawdata.py:194 */
    int32_t enable_gate_afterpulses;     /* rawdata.py:198 */
    int32_t save_full_truth;        /* rawdata.py:42 */
    double p_double_pe_emision;     /* (sic) pulse.py:76 */
    double pmt_transit_time_mean, pmt_transit_time_spread; /* pulse.py:54-55 */
    double s1_detection_efficiency; /* s1.py:131 */
    double s1_decay_time, s1_decay_spread; /* s1.py:193-194 */
    double singlet_fraction_gas, singlet_lifetime_gas, triplet_lifetime_gas; /* pulse.py:321-341 */
    double drift_velocity_liquid, drift_time_gate, diffusion_constant_longitudinal; /* s2.py:139-179 */
    double electron_lifetime_liquid, electron_extraction_yield, electron_trapping_time; /* s2.py:212-286 */
    double s2_secondary_sc_gain, s2_gain_spread, s2_time_spread; /* s2.py:197,309,550 */
    double tpc_radius, tpc_length;
    double pmt_ap_modifier, pmt_ap_t_modifier; /* afterpulse.py:193-223 */
    double photoionization_modifier;           /* afterpulse.py:39 */
    double photoelectric_modifier, photoelectric_p, photoelectric_t_center, photoelectric_t_spread; /* afterpulse.py:108-115 */
    double ele_ap_n;                /* uniform_to_ele_ap.n, afterpulse.py:37 */
    double s2_aft_sigma, s2_aft_skewness; /* s2.py:630-631 (applied on the host, per-instruction pattern rows) */
    /* S1 'custom' model (s1.py:201-217, 263-337): timing by recoil type (NestId, s1.py:19-30) */
    int32_t s1_model_custom;        /* 'custom' in s1_model_type */
    int32_t gf_avgt;                /* int(np.average(s2_luminescence['t'])), s2.py:408 */
    double singlet_lifetime_liquid, triplet_lifetime_liquid;   /* pulse.py:331-333 */
    double s1_ER_alpha_singlet_fraction, s1_NR_singlet_fraction, led_pulse_length; /* s1.py:271,279,337 */
    /* S2 'garfield' luminescence (s2.py:381-409) */
    double anode_xaxis_angle, anode_pitch;      /* defaults pi/4, 0.5 */
    double s2_garfield_confine_position;        /* <= 0: distance to the nearest anode wire from xy */
} wfs_params;

/* Tables (host pointers, copied to the device at wfs_create).  A NULL pointer means "absent". */
typedef struct wfs_tables {
    const double *templates;        /* [dt][template_length], pulse.py:146-187 */
    const double *gains;            /* [n_tpc_pmts], strax_interface.py:584-587; 0 == turned off */
    const int32_t *zle_thresholds;  /* [n_rows] baseline - (special|zle)_threshold - 1, rawdata.py:290-294 */
    const double *noise;            /* [noise_len][noise_nch] resource.noise_data, rawdata.py:264-268 */
    int64_t noise_len;
    int32_t noise_nch;
    /* SPE inverse CDF: pulse.py:189-227.  spe_row[ch] selects the row of spe_ppf used for ch */
    const double *spe_ppf;          /* [n_spe_rows][spe_len] */
    const int32_t *spe_row;         /* [n_tpc_pmts] */
    int32_t n_spe_rows, spe_len;    /* spe_len == 2001 */
    /* S2 luminescence 'simple' model: inverse CDF of emission time tabulated once for the
     * constant gas gap (s2.py:317-378): t = interp(U, lum_cdf, lum_t) */
    const double *lum_cdf, *lum_t;
    int32_t lum_len;
    /* PMT afterpulse elements (resource.uniform_to_pmt_ap, afterpulse.py:172-249) */
    int32_t n_ap_elements;
    int32_t ap_is_uniform[WFS_MAX_AP_ELEMENTS];
    const double *ap_delay_cdf[WFS_MAX_AP_ELEMENTS];   /* [n_tpc_pmts][ap_delay_len] */
    int32_t ap_delay_len[WFS_MAX_AP_ELEMENTS];
    double ap_delay_bin[WFS_MAX_AP_ELEMENTS];
    const double *ap_amp_cdf[WFS_MAX_AP_ELEMENTS];     /* [n_tpc_pmts or 1][ap_amp_len] */
    int32_t ap_amp_len[WFS_MAX_AP_ELEMENTS];
    int32_t ap_amp_rows[WFS_MAX_AP_ELEMENTS];
    double ap_amp_bin[WFS_MAX_AP_ELEMENTS];
    /* photo-ionisation electrons (uniform_to_ele_ap, afterpulse.py:33-80): the coarse delay grid of
     * _reduce_instruction_timing and the probability that one delay drawn from the delay
     * histogram falls into each coarse bin (np.digitize convention) */
    const double *pi_coarse_time;   /* [pi_coarse_len] */
    const double *pi_coarse_prob;   /* [pi_coarse_len] */
    int32_t pi_coarse_len;
    /* optical propagation delays (resource.s1/s2_optical_propagation_spline, regular grids with
     * linear interpolation / extrapolation as scipy's RegularGridInterpolator(fill_value=None)):
     * S1: delay(z, U) per array (s1.py:241-260); S2: delay(U) per array (s2.py:486-501) */
    const double *s1_op_top, *s1_op_bottom;     /* [s1_op_nz][s1_op_nu] */
    int32_t s1_op_nz, s1_op_nu;
    double s1_op_z0, s1_op_z1, s1_op_u0, s1_op_u1;
    const double *s2_op_top, *s2_op_bottom;     /* [s2_op_nu] */
    int32_t s2_op_nu;
    double s2_op_u0, s2_op_u1;
    /* S2 'garfield' luminescence (resource.s2_luminescence, s2.py:381-409) */
    const int32_t *gf_t;            /* [gf_rows][gf_cols] emission times */
    const double *gf_x;             /* [gf_rows] distance to the anode wire of each row */
    int32_t gf_rows, gf_cols;
    /* S2 'garfield_gas_gap' luminescence (resource.s2_luminescence_gg['timing_inv_cdf'], s2.py:411-483):
     * inverse CDFs of the excitation time, one row per tabulated gas gap */
    const double *gg_cdf;           /* [gg_rows][gg_len] */
    int32_t gg_rows, gg_len;
    /* PMT pattern maps on regular grids (resource.s1_pattern_map over (x, y, z), s1.py:148;
     * resource.s2_pattern_map over the observed (x, y), s2.py:637-645), evaluated on the device for
     * every instruction whose wfs_instr_maps.pattern_row is negative: multilinear interpolation with
     * linear extrapolation outside the grid, the arithmetic of scipy's RegularGridInterpolator
     * (fill_value=None) that straxen.InterpolatingMap wraps (load_resource.py:399,433), result
     * rounded to float32 like the host-evaluated rows.  An S2 map with fewer than n_tpc_pmts columns
     * (top array only) is padded with ones (s2.py:642-644).  NULL -> host rows only. */
    const double *s1_pat_grid;      /* [n0][n1][n2][s1_pat_npmt] */
    int32_t s1_pat_n[3], s1_pat_npmt;
    double s1_pat_lo[3], s1_pat_hi[3];
    const double *s2_pat_grid;      /* [n0][n1][s2_pat_npmt] */
    int32_t s2_pat_n[2], s2_pat_npmt, s2_pat_pad;
    double s2_pat_lo[2], s2_pat_hi[2];
    /* 'simple' S2 luminescence with enable_gas_gap_warping (s2.py:317-378 with resource.gas_gap_length):
     * scalars of the field model between liquid surface and anode wire -- alpha =
     * gas_drift_velocity_slope / number density, uE = kV/cm, pressure in bar, anode_field_domination_distance,
     * anode_wire_radius, radial step (0.0001 cm).  Used with wfs_instr_maps.lum_gap / lum_e0;
     * lumw_dr == 0 -> not available. */
    double lumw_alpha, lumw_ue, lumw_pressure, lumw_ra, lumw_rw, lumw_dr;
} wfs_tables;

/* Per-instruction map values evaluated on the host with the reference's own map objects
 * (straxen.InterpolatingMap is third party; DummyMap for config[0]).  All arrays [n_instr]
 * unless noted; NULL -> the constant stated. */
typedef struct wfs_instr_maps {
    const double *s1_lce;           /* s1_lce_correction_map(xyz), s1.py:125; NULL -> 1 */
    const double *s2_sc_gain;       /* get_s2_light_yield(positions) incl. /(1+p_dpe), s2.py:182-209 */
    const double *s2_cy_extra;      /* p_surv (and map-driven extraction yield) factor, s2.py:227-252; NULL -> 1 */
    const float *pattern;           /* [n_pattern_rows][n_tpc_pmts] un-normalised per-PMT pattern */
    const int32_t *pattern_row;     /* [n_instr] row of `pattern` for the instruction; NULL -> row 0;
                                     * negative -> evaluate the device-resident pattern grid (wfs_tables) */
    int64_t n_pattern_rows;
    double s2_sc_gain_default;      /* used when s2_sc_gain is NULL */
    const uint64_t *rng_id;         /* [n_instr] Philox identity of each instruction; NULL -> its index.
                                     * Shards of one instruction set pass the global indices so that the
                                     * result does not depend on the sharding. */
    /* field dependencies (s2.py:139-179): per-instruction drift velocity [cm/ns] and longitudinal
     * diffusion constant [cm^2/ns]; NULL -> the config constants */
    const double *drift_velocity;
    const double *diffusion_long;
    /* observed S2 position after the field-distortion model (s2.py:80-87); used by the garfield
     * luminescence model for the distance to the anode wires; NULL -> the instruction's x, y */
    const double *x_obs, *y_obs;
    /* Externally supplied photons (RawDataOptical.sim_primary, rawdata.py:478-495: G4 optical output,
     * neutron-veto style inputs): a type-1 instruction with opt_last[i] > opt_first[i] takes its photons
     * from the lists below instead of sampling them -- channel opt_channels[k], arrival time
     * instruction time + opt_timings[k], k in [opt_first[i], opt_last[i]); photons with a timing < 0
     * or >= opt_time_cutoff (config nveto_time_max_cutoff) are dropped.  Transit-time spread, double
     * photo-electrons and SPE gains are sampled for them as for any photon (Pulse.__call__).  All NULL /
     * 0 -> none. */
    const int64_t *opt_first, *opt_last;    /* [n_instr] (the `_first`, `_last` columns) */
    const int32_t *opt_channels;            /* [n_opt] */
    const int64_t *opt_timings;             /* [n_opt] ns relative to the instruction time */
    int64_t n_opt;
    int64_t opt_time_cutoff;
    /* 'garfield_gas_gap' luminescence (s2.py:460-483), per S2-like instruction from
     * resource.garfield_gas_gap_map at the observed position: the two rows of gg_cdf the gas gap lies
     * between (np.digitize - 1 and the row above, clipped; python's negative index included) and
     * (gas gap - gas_gap[row]) / row spacing.  Required when s2_luminescence_model == 2. */
    const int32_t *gg_lo_row, *gg_hi_row;
    const double *gg_frac;
    /* Transverse diffusion of the S2 hit pattern (S2.s2_pattern_map_diffuse, s2.py:560-613, with
     * enable_field_dependencies.diffusion_transverse_map): per S2-like instruction the sigma [cm] of the
     * radial and azimuthal displacement of one electron at the liquid surface,
     * sqrt(2 * D_radial|azimuthal * drift_time_mean).  The pattern of the instruction is then the average
     * of the device-resident S2 pattern grid over its electrons' displaced positions inside tpc_radius
     * (requires pattern_row < 0 for these instructions).  Both NULL -> pattern at the observed position. */
    const double *hdiff_sigma_r, *hdiff_sigma_a;
    /* 'simple' luminescence with enable_gas_gap_warping: per S2-like instruction the local gas gap dG
     * [cm] (resource.gas_gap_length(xy), s2.py:361-362) and the field scale E0 [V/cm] derived from it
     * (s2.py:365-370).  Required when s2_luminescence_model == 0 and no constant-gap table was given. */
    const double *lum_gap, *lum_e0;
} wfs_instr_maps;

typedef struct wfs_counts {
    int64_t n_records[3];           /* raw_records, raw_records_he, raw_records_aqmon */
    int64_t n_records_total;
    int64_t n_truth;
    int64_t n_photons;              /* photons superposed (after dead-PMT removal) */
    int64_t n_pe;                   /* sum of truth n_pe (photons + DPE), pulse.py:262 */
    int64_t n_pulses;               /* (pulse call, channel) pulses */
    int64_t n_windows;              /* (digitisation group, channel) windows incl. HE rows */
    int64_t n_intervals;            /* ZLE intervals */
    int64_t n_samples;              /* samples digitised (window samples) */
    int64_t n_groups;
    int64_t n_pulse_calls;
    int64_t n_instructions;
    int64_t n_batches;
    int64_t gpu_launches;           /* kernels launched by this call */
    int64_t need_records;           /* capacities needed when WFS_E_CAPACITY */
    int64_t need_truth;
    int64_t need_groups;
    int64_t need_batches;
    int64_t d2h_bytes;              /* record bytes that crossed PCIe (compact transport or 244 B/record) */
    double ms_total;                /* CUDA-event time of the device work of this call */
    double ms_digitize;             /* of which: the digitize (superpose+ADC+noise+clip) kernel */
    double ms_h2d, ms_d2h;
    /* device time per phase, summed over batches (CUDA events on the library stream):
     * 0 sampling front end, 1 photon keys + sort, 2 pulses/windows, 3 digitize, 4 ZLE,
     * 5 record keys + sort, 6 record pack (or plain -> compact transport form), 7 host scheduler + truth (wall clock); compact transport
     * (wall clock): 8 batch shipped -> its D2H copies complete, 9 copies complete -> expanded;
     * 10 / 11: number of batches whose photons / records were ordered per group in shared memory */
    double ms_phase[12];
    int64_t n_fused_batches;        /* device batches that went through the group-resident fused kernel (one CTA per
                                     * digitisation group, photons -> records; then ms_phase[3] is that kernel and
                                     * phases 1, 2, 4, 5 are zero; 3 = k_group_analyse, 6 = k_group_records
                                     * [+ plain -> compact transport form]) */
    int64_t n_plain_records;        /* records that crossed PCIe as plain 244-byte rows (split transport into a
                                     * page-locked destination; the rest travelled compact and was expanded on the host) */
} wfs_counts;

/* Digitisation-group bookkeeping returned to the host-side chunker
 * (replaces RawData.left/right as read by ChunkRawRecords, strax_interface.py:394-403). */
typedef struct wfs_group_info {
    int64_t left, right;            /* rawdata.py:215-222 (left made even) */
    int64_t n_intervals;            /* ZLE intervals yielded by the group */
} wfs_group_info;

/* Lifetime.  `device` is the CUDA ordinal this handle is bound to (one handle per GPU). */
int wfs_create(const wfs_params *params, const wfs_tables *tables, int device, void **handle);
void wfs_destroy(void *handle);
const char *wfs_last_error(void *handle);   /* handle may be NULL for create errors */
int wfs_abi_version(void);
void wfs_struct_sizes(int64_t *out6);       /* sizeof of the six structs of this header, for binding checks */
int wfs_device_count(void);

/* Smallest signal-time gap [ns] at which the caller may cut a run into independent calls (plugin
 * pieces, GPU shards): right_raw_extension (rawdata.py:63) plus the longest delay of the secondary
 * instructions an S2 can spawn (afterpulse.py:63-80, 105-115).  The library cuts its own device batches
 * at the same gaps.  Replaces nothing in the reference (which simulates a run in one generator). */
int64_t wfs_quiet_gap(void *handle);

/* Pinned host memory for output buffers (so device->host copies are plain DMA). */
void *wfs_host_alloc(int64_t bytes);
void wfs_host_free(void *p);
/* Page-locks / releases memory the caller owns (the record arenas of ChunkRawRecords,
 * strax_interface.py:387-392: one buffer reused for the whole run).  Into a page-locked destination
 * wfs_simulate sends part of every batch as plain 244-byte rows by DMA (no host core involved) and the rest
 * in the compact form, the share following which of the two finished first for the previous batches; into
 * ordinary memory everything travels compact.  The records are the same either way.  Returns 0 or WFS_E_CUDA. */
int wfs_host_register(void *p, int64_t bytes);
int wfs_host_unregister(void *p);

/* Host half of the compact record transport (csrc/transport.cuh): when the destination of the records
 * is host memory they cross PCIe as 24-byte headers + the 8-byte (4-sample) blocks that differ from the
 * fill pattern (baseline below `length`, 0 behind it) and are expanded into 244-byte raw_records
 * (strax_interface.py:425-436) by host threads.  wfs_simulate / wfs_simulate_photons do this
 * internally; this entry expands a compact batch the caller holds (and lets the expander be tested
 * without a GPU).  hdr: n_records x {i64 time, i32 pulse_length, i16 channel, i16 record_i,
 * u32 first block, u32 block mask} (length = min(pulse_length - 110 record_i, 110)); blocks: 8 bytes
 * each.  Returns 0. */
int wfs_expand_compact(const void *hdr, const void *blocks, int64_t n_records, uint8_t *records,
                       int fill, int dt, int n_threads);

/* Deterministic entry: photons in, records out.
 * Replaces Pulse.__call__ with preset gains (pulse.py:82-144) + Pulse.add_current (:276-318) +
 * RawData.digitize_pulse_cache (rawdata.py:204-272) + RawData.ZLE (:274-311) + the record packing
 * and (time, channel) sort of ChunkRawRecords (strax_interface.py:391-436,446-453).
 *   pulse_call[i]  id of the Pulse call photon i belongs to (0..n_pulse_calls-1)
 *   group_of[p]    digitisation group of pulse call p (0..n_groups-1)
 *   ix_rand[g]     noise start offset of group g (rawdata.py:407-417); NULL -> a Philox draw keyed by
 *                  (seed, first sample of the group's digitisation window): unique per group, and the
 *                  same however a run is cut into calls, device batches or GPU shards
 * Output: `records` holds [raw_records | raw_records_he | raw_records_aqmon], each segment sorted
 * by (time, channel); segment lengths in counts->n_records[].  `groups` (optional, [n_groups]).
 * Host pointers unless `on_device` is non-zero (then all photon arrays and `records` are device
 * pointers and no host<->device copy happens inside the call). */
int wfs_simulate_photons(void *handle, int64_t n_photons, const int64_t *t_ns,
                         const int32_t *channel, const double *gain, const int32_t *pulse_call,
                         int64_t n_pulse_calls, const int32_t *group_of, int64_t n_groups,
                         const int64_t *ix_rand, uint64_t seed, int on_device,
                         uint8_t *records, int64_t cap_records, wfs_group_info *groups,
                         wfs_counts *counts);

/* Output buffers of the full path (caller-allocated host memory; NULL/0 = not wanted). */
typedef struct wfs_outputs {
    uint8_t *records;               /* 244-byte rows; per batch [raw_records | _he | _aqmon], batches in time order */
    int64_t cap_records;
    uint8_t *truth;                 /* 218-byte rows, one per Pulse call, in execution order */
    int64_t cap_truth;
    wfs_group_info *groups;         /* one per digitisation group, in time order */
    int64_t cap_groups;
    int64_t *batch_records;         /* [cap_batches][3]: records per data type of each batch */
    int64_t cap_batches;
    /* per_pmt_truth (strax_interface.py:77-116, pulse.py:257-269): per truth row and PMT the four
     * counters n_photon, n_pe, n_photon_trigger, n_pe_trigger and the two areas raw_area,
     * raw_area_trigger of Pulse.add_truth.  NULL -> not produced. */
    int32_t *truth_pmt_counts;      /* [cap_truth][4][n_tpc_pmts] */
    double *truth_pmt_areas;        /* [cap_truth][2][n_tpc_pmts] */
} wfs_outputs;

/* Full path: instructions in, records + truth out.
 * Replaces ChunkRawRecords.__call__ -> RawData.__call__ -> S1/S2/afterpulse Pulse calls ->
 * digitize_pulse_cache -> ZLE -> record packing (strax_interface.py:368-497, rawdata.py:38-375).
 *   instructions   packed 70-byte rows (instruction_dtype); any order
 *   maps           per-instruction map values (see wfs_instr_maps)
 * When a capacity is too small the call still runs every batch (to learn the sizes), returns
 * WFS_E_CAPACITY and leaves the needed sizes in counts->need_*; a second call with larger
 * buffers reproduces the same data (Philox). */
int wfs_simulate(void *handle, const uint8_t *instructions, int64_t n_instructions,
                 const wfs_instr_maps *maps, uint64_t seed, wfs_outputs *out, wfs_counts *counts);

/* The scheduling decisions of the full path on their own (host code, no GPU): which instructions form
 * a Pulse call and which Pulse calls are digitised together.
 * Replaces the bookkeeping of RawData.__call__ (wfsim/core/rawdata.py:61-63 signal times and primary
 * clusters, :87-93 re-clustering with pending secondaries, :96-98 when the pulse cache is pushed out,
 * :102-150 Pulse calls per type incl. save_full_truth on / off and stop_at_this_group) given what the
 * sampling stages produced: the end of the last pulse of every instruction and the secondary (type 4 / 6)
 * instructions every S2 spawned.  wfs_simulate runs the same function on the device's numbers.
 *   time, z, type      [n_prim] primary instructions, any order
 *   sec_*              [n_sec] secondary instructions and the primary (index) that spawned each
 *   pulse_end          [n_prim + n_sec] max(right) * dt over the pulses of the instruction incl. its PMT
 *                      afterpulses (rawdata.py:186-190); INT64_MIN = the instruction made no pulse
 * Output: run_of[n_prim + n_sec] = Pulse call of each instruction in execution order (-1: none),
 * run_type / run_group [cap_runs] = type and digitisation group of each Pulse call; *n_groups = number
 * of groups that hold pulses (trailing calls that made no pulse carry the next, unused index). */
int wfs_schedule(int64_t right_raw_extension, double drift_velocity_liquid, int save_full_truth,
                 int64_t n_prim, const int64_t *time, const float *z, const int8_t *type,
                 int64_t n_sec, const int64_t *sec_time, const float *sec_z, const int8_t *sec_type,
                 const int32_t *sec_parent, const int64_t *pulse_end, int32_t *run_of, int32_t *run_type,
                 int32_t *run_group, int64_t cap_runs, int64_t *n_runs, int64_t *n_groups);

/* Device-resident variant for throughput measurement: instructions/maps are parsed, planned and
 * uploaded once by wfs_stage_instructions; wfs_run_staged then runs the whole path with every
 * buffer in HBM and leaves the records on the device (counts are returned; `out` may carry
 * truth/groups buffers, its records pointer is ignored). */
int wfs_stage_instructions(void *handle, const uint8_t *instructions, int64_t n_instructions,
                           const wfs_instr_maps *maps);
int wfs_run_staged(void *handle, uint64_t seed, wfs_outputs *out, wfs_counts *counts);

/* Stage-level dump used by the statistical parity tests: runs only the sampling front end (the
 * same Philox streams the full path uses) and returns 32-byte rows
 *   stage 0 (photons):  int64 t_ns; double gain; int32 channel; int32 instruction (index into the
 *                       given array; the parent S2 for secondaries); int32 flags (1 = double-pe,
 *                       2 = PMT afterpulse, 4 = photo-ionisation secondary); int32 secondary id
 *   stage 1 (emitters): int64 t_ns; double 0; int32 n_photons; int32 instruction; int32 flags; int32 id
 *   stage 2 / 3 (secondary instructions, afterpulse.py:24-139): int64 time; double z (2) or x^2 + y^2 (3);
 *                       int32 amp; int32 parent instruction; int32 type (4 / 6); int32 id
 *   stage 4 (secondary instructions in full): int64 time; float32 x, y; int32 amp; int32 parent
 *                       instruction; int32 type; float32 z -- row k is the secondary with id k
 * Replaces nothing in the reference; it exposes S1/S2.photon_timings/photon_channels (s1.py:138-238,
 * s2.py:258-315,504-682) and Pulse.__call__'s TTS/DPE/SPE draws (pulse.py:53-103) for testing. */
int wfs_sample_stage(void *handle, int stage, const uint8_t *instructions, int64_t n_instructions,
                     const wfs_instr_maps *maps, uint64_t seed, void *out, int64_t cap, int64_t *n_out);

#ifdef __cplusplus
}
#endif
#endif
