"""ctypes binding of libwfsim_b200.so (C-ABI declared in include/wfsim_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module
raises, and every compute entry point fails when no CUDA device is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'csrc', 'libwfsim_b200.so')

ABI_VERSION = 5
E_CAPACITY = 1
E_CUDA = -1
E_ARG = -2
E_PULSE_CACHE_TOO_LONG = -3
E_KEYBITS = -4
MAX_AP_ELEMENTS = 8
RECORD_BYTES = 244

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class Params(C.Structure):
    _fields_ = (
        [(n, i32) for n in (
            'abi_version', 'detector_nt', 'n_tpc_pmts', 'n_top_pmts', 'he_first', 'he_last',
            'he_mult', 'n_rows', 'dt', 'template_length', 'pulse_left_margin',
            'pulse_right_margin', 'trigger_window', 'baseline', 'enable_noise', 'zle_threshold')]
        + [('current_2_adc', f64), ('right_raw_extension', i64)]
        + [(n, i32) for n in (
            's1_model_simple', 's1_model_optical', 's2_luminescence_model', 's2_time_model',
            'enable_pmt_afterpulses', 'enable_electron_afterpulses', 'enable_gate_afterpulses',
            'save_full_truth')]
        + [(n, f64) for n in (
            'p_double_pe_emision', 'pmt_transit_time_mean', 'pmt_transit_time_spread',
            's1_detection_efficiency', 's1_decay_time', 's1_decay_spread',
            'singlet_fraction_gas', 'singlet_lifetime_gas', 'triplet_lifetime_gas',
            'drift_velocity_liquid', 'drift_time_gate', 'diffusion_constant_longitudinal',
            'electron_lifetime_liquid', 'electron_extraction_yield', 'electron_trapping_time',
            's2_secondary_sc_gain', 's2_gain_spread', 's2_time_spread',
            'tpc_radius', 'tpc_length', 'pmt_ap_modifier', 'pmt_ap_t_modifier',
            'photoionization_modifier', 'photoelectric_modifier', 'photoelectric_p',
            'photoelectric_t_center', 'photoelectric_t_spread', 'ele_ap_n',
            's2_aft_sigma', 's2_aft_skewness')]
        + [('s1_model_custom', i32), ('gf_avgt', i32)]
        + [(n, f64) for n in (
            'singlet_lifetime_liquid', 'triplet_lifetime_liquid', 's1_ER_alpha_singlet_fraction',
            's1_NR_singlet_fraction', 'led_pulse_length', 'anode_xaxis_angle', 'anode_pitch',
            's2_garfield_confine_position')])


class Tables(C.Structure):
    _fields_ = [
        ('templates', vp), ('gains', vp), ('zle_thresholds', vp), ('noise', vp),
        ('noise_len', i64), ('noise_nch', i32),
        ('spe_ppf', vp), ('spe_row', vp), ('n_spe_rows', i32), ('spe_len', i32),
        ('lum_cdf', vp), ('lum_t', vp), ('lum_len', i32),
        ('n_ap_elements', i32), ('ap_is_uniform', i32 * MAX_AP_ELEMENTS),
        ('ap_delay_cdf', vp * MAX_AP_ELEMENTS), ('ap_delay_len', i32 * MAX_AP_ELEMENTS),
        ('ap_delay_bin', f64 * MAX_AP_ELEMENTS),
        ('ap_amp_cdf', vp * MAX_AP_ELEMENTS), ('ap_amp_len', i32 * MAX_AP_ELEMENTS),
        ('ap_amp_rows', i32 * MAX_AP_ELEMENTS), ('ap_amp_bin', f64 * MAX_AP_ELEMENTS),
        ('pi_coarse_time', vp), ('pi_coarse_prob', vp), ('pi_coarse_len', i32),
        ('s1_op_top', vp), ('s1_op_bottom', vp), ('s1_op_nz', i32), ('s1_op_nu', i32),
        ('s1_op_z0', f64), ('s1_op_z1', f64), ('s1_op_u0', f64), ('s1_op_u1', f64),
        ('s2_op_top', vp), ('s2_op_bottom', vp), ('s2_op_nu', i32),
        ('s2_op_u0', f64), ('s2_op_u1', f64),
        ('gf_t', vp), ('gf_x', vp), ('gf_rows', i32), ('gf_cols', i32),
        ('gg_cdf', vp), ('gg_rows', i32), ('gg_len', i32),
        ('s1_pat_grid', vp), ('s1_pat_n', i32 * 3), ('s1_pat_npmt', i32), ('s1_pat_lo', f64 * 3), ('s1_pat_hi', f64 * 3),
        ('s2_pat_grid', vp), ('s2_pat_n', i32 * 2), ('s2_pat_npmt', i32), ('s2_pat_pad', i32),
        ('s2_pat_lo', f64 * 2), ('s2_pat_hi', f64 * 2),
        ('lumw_alpha', f64), ('lumw_ue', f64), ('lumw_pressure', f64), ('lumw_ra', f64), ('lumw_rw', f64),
        ('lumw_dr', f64)]


class InstrMaps(C.Structure):
    _fields_ = [('s1_lce', vp), ('s2_sc_gain', vp), ('s2_cy_extra', vp), ('pattern', vp),
                ('pattern_row', vp), ('n_pattern_rows', i64), ('s2_sc_gain_default', f64),
                ('rng_id', vp), ('drift_velocity', vp), ('diffusion_long', vp), ('x_obs', vp), ('y_obs', vp),
                ('opt_first', vp), ('opt_last', vp), ('opt_channels', vp), ('opt_timings', vp),
                ('n_opt', i64), ('opt_time_cutoff', i64), ('gg_lo_row', vp), ('gg_hi_row', vp), ('gg_frac', vp),
                ('hdiff_sigma_r', vp), ('hdiff_sigma_a', vp), ('lum_gap', vp), ('lum_e0', vp)]


class Outputs(C.Structure):
    _fields_ = [('records', vp), ('cap_records', i64), ('truth', vp), ('cap_truth', i64),
                ('groups', vp), ('cap_groups', i64), ('batch_records', vp), ('cap_batches', i64),
                ('truth_pmt_counts', vp), ('truth_pmt_areas', vp)]


class Counts(C.Structure):
    _fields_ = ([('n_records', i64 * 3)]
                + [(n, i64) for n in (
                    'n_records_total', 'n_truth', 'n_photons', 'n_pe', 'n_pulses', 'n_windows',
                    'n_intervals', 'n_samples', 'n_groups', 'n_pulse_calls', 'n_instructions',
                    'n_batches', 'gpu_launches', 'need_records', 'need_truth', 'need_groups',
                    'need_batches', 'd2h_bytes')]
                + [(n, f64) for n in ('ms_total', 'ms_digitize', 'ms_h2d', 'ms_d2h')]
                + [('ms_phase', f64 * 12), ('n_fused_batches', i64), ('n_plain_records', i64)])

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n not in ('n_records', 'ms_phase')}
        d['n_records'] = list(self.n_records)
        d['ms_phase'] = list(self.ms_phase)
        return d


class GroupInfo(C.Structure):
    _fields_ = [('left', i64), ('right', i64), ('n_intervals', i64)]


EXPORTS = ['wfs_create', 'wfs_destroy', 'wfs_last_error', 'wfs_abi_version', 'wfs_struct_sizes',
           'wfs_device_count', 'wfs_host_alloc', 'wfs_host_free', 'wfs_host_register', 'wfs_host_unregister',
           'wfs_simulate_photons',
           'wfs_simulate', 'wfs_stage_instructions', 'wfs_run_staged', 'wfs_sample_stage',
           'wfs_expand_compact', 'wfs_quiet_gap', 'wfs_schedule']

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load():
    """Load the CUDA extension; raises LibraryMissing (never falls back to a CPU path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise LibraryMissing(
            f'{LIB_PATH} not built: run `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C wfsim_b200/csrc`.  wfsim_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise LibraryMissing(f'{LIB_PATH} does not export {name}')
    lib.wfs_last_error.restype = C.c_char_p
    lib.wfs_last_error.argtypes = [vp]
    lib.wfs_host_alloc.restype = vp
    lib.wfs_host_alloc.argtypes = [i64]
    lib.wfs_host_free.argtypes = [vp]
    lib.wfs_host_register.argtypes = [vp, i64]
    lib.wfs_host_unregister.argtypes = [vp]
    lib.wfs_destroy.argtypes = [vp]
    lib.wfs_schedule.argtypes = [i64, f64, C.c_int, i64, vp, vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64,
                                 C.POINTER(i64), C.POINTER(i64)]
    lib.wfs_quiet_gap.restype = i64
    lib.wfs_quiet_gap.argtypes = [vp]
    lib.wfs_expand_compact.argtypes = [vp, vp, i64, vp, C.c_int, C.c_int, C.c_int]
    lib.wfs_create.argtypes = [C.POINTER(Params), C.POINTER(Tables), C.c_int, C.POINTER(vp)]
    lib.wfs_simulate_photons.argtypes = [
        vp, i64, vp, vp, vp, vp, i64, vp, i64, vp, C.c_uint64, C.c_int, vp, i64, vp,
        C.POINTER(Counts)]
    lib.wfs_simulate.argtypes = [
        vp, vp, i64, C.POINTER(InstrMaps), C.c_uint64, C.POINTER(Outputs), C.POINTER(Counts)]
    lib.wfs_stage_instructions.argtypes = [vp, vp, i64, C.POINTER(InstrMaps)]
    lib.wfs_run_staged.argtypes = [vp, C.c_uint64, C.POINTER(Outputs), C.POINTER(Counts)]
    lib.wfs_sample_stage.argtypes = [vp, C.c_int, vp, i64, C.POINTER(InstrMaps), C.c_uint64,
                                     vp, i64, C.POINTER(i64)]
    sizes = (i64 * 6)()
    lib.wfs_struct_sizes(sizes)
    want = [C.sizeof(Params), C.sizeof(Tables), C.sizeof(InstrMaps), C.sizeof(Counts),
            C.sizeof(GroupInfo), C.sizeof(Outputs)]
    if list(sizes) != want:
        raise LibraryMissing(f'struct layout mismatch between lib.py {want} and the .so {list(sizes)}')
    if lib.wfs_abi_version() != ABI_VERSION:
        raise LibraryMissing('ABI version mismatch')
    _lib = lib
    return lib
