"""Flatten the fax_config dict + resource tables into the C-ABI structs (include/wfsim_b200.h)."""
import ctypes as C

import numpy as np

from . import lib as wlib
from . import tables as wtab
from .config import current_2_adc

S2_LUM_MODELS = {'simple': 0, 'garfield': 1, 'garfield_gas_gap': 2}


def _s2_time_model(name):
    if 'optical_propagation' in name:
        return 2
    if 'zero_delay' in name:
        return 1
    if 's2_time_spread around zero' in name:
        return 0
    raise KeyError(f"{name} is not in any of the valid s2 time models")


class HostTables:
    """Keeps numpy arrays alive for the lifetime of the Tables struct."""

    def __init__(self):
        self.struct = wlib.Tables()
        self._keep = []

    def set(self, field, arr, dtype):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        self._keep.append(arr)
        setattr(self.struct, field, arr.ctypes.data)
        return arr

    def set_elem(self, field, k, arr, dtype):
        arr = np.ascontiguousarray(arr, dtype=dtype)
        self._keep.append(arr)
        getattr(self.struct, field)[k] = arr.ctypes.data
        return arr


def build_params(cfg):
    p = wlib.Params()
    p.abi_version = wlib.ABI_VERSION
    nt = cfg.get('detector', 'XENONnT') == 'XENONnT'
    p.detector_nt = int(nt)
    p.n_tpc_pmts = len(cfg['gains'])
    p.n_top_pmts = int(cfg.get('n_top_pmts', 253))
    he = cfg.get('channel_map', {}).get('he', (500, 752))
    p.he_first, p.he_last = int(he[0]), int(he[-1])
    p.he_mult = int(cfg.get('high_energy_deamplification_factor', 0)) if nt else 0
    p.n_rows = wtab.N_ROWS
    p.dt = int(cfg.get('sample_duration', 10))
    sb = cfg.get('samples_before_pulse_center', 2)
    sa = cfg.get('samples_after_pulse_center', 20)
    p.template_length = int(sb + sa)
    p.pulse_left_margin = int(cfg['samples_to_store_before']) + int(sb)
    p.pulse_right_margin = int(cfg['samples_to_store_after']) + int(sa)
    p.trigger_window = int(cfg['trigger_window'])
    p.baseline = int(cfg['digitizer_reference_baseline'])
    p.enable_noise = int(bool(cfg.get('enable_noise', True)))
    p.zle_threshold = int(cfg['zle_threshold'])
    p.current_2_adc = current_2_adc(cfg)
    p.right_raw_extension = int(cfg.get('right_raw_extension', 100000))
    s1m = cfg.get('s1_model_type', 'simple')
    p.s1_model_simple = int('simple' in s1m)
    p.s1_model_optical = int('optical_propagation' in s1m)
    p.s2_luminescence_model = S2_LUM_MODELS[cfg.get('s2_luminescence_model', 'simple')]
    p.s2_time_model = _s2_time_model(cfg.get('s2_time_model', 's2_time_spread around zero'))
    p.enable_pmt_afterpulses = int(bool(cfg.get('enable_pmt_afterpulses', True)))
    p.enable_electron_afterpulses = int(bool(cfg.get('enable_electron_afterpulses', True)))
    p.enable_gate_afterpulses = int(bool(cfg.get('enable_gate_afterpulses', False)))
    p.save_full_truth = int(bool(cfg.get('save_full_truth', True)))
    for name in ('p_double_pe_emision', 'pmt_transit_time_mean', 'pmt_transit_time_spread',
                 's1_detection_efficiency', 's1_decay_time', 's1_decay_spread',
                 'singlet_fraction_gas', 'singlet_lifetime_gas', 'triplet_lifetime_gas',
                 'drift_velocity_liquid', 'drift_time_gate', 'diffusion_constant_longitudinal',
                 'electron_lifetime_liquid', 'electron_extraction_yield', 'electron_trapping_time',
                 's2_secondary_sc_gain', 's2_time_spread', 'tpc_radius', 'tpc_length'):
        setattr(p, name, float(cfg.get(name, 0.0)))
    p.s2_gain_spread = float(cfg.get('s2_gain_spread', 0))
    p.pmt_ap_modifier = float(cfg.get('pmt_ap_modifier', 1))
    p.pmt_ap_t_modifier = float(cfg.get('pmt_ap_t_modifier', 0))
    p.photoionization_modifier = float(cfg.get('photoionization_modifier', 1))
    p.photoelectric_modifier = float(cfg.get('photoelectric_modifier', 1))
    p.photoelectric_p = float(cfg.get('photoelectric_p', 0))
    p.photoelectric_t_center = float(cfg.get('photoelectric_t_center', 0))
    p.photoelectric_t_spread = float(cfg.get('photoelectric_t_spread', 0))
    p.s2_aft_sigma = float(cfg.get('s2_aft_sigma', 0.0))
    p.s2_aft_skewness = float(cfg.get('s2_aft_skewness', 0.0))
    if 'nest' in s1m:
        # s1.py:222-234 calls nestpy.GetPhotonTimes (third-party, not in this image)
        raise NotImplementedError("s1_model_type 'nest' needs nestpy, which the device path does not bind")
    p.s1_model_custom = int('custom' in s1m)
    for name in ('singlet_lifetime_liquid', 'triplet_lifetime_liquid', 's1_ER_alpha_singlet_fraction',
                 's1_NR_singlet_fraction', 'led_pulse_length'):
        setattr(p, name, float(cfg.get(name, 0.0)))
    p.anode_xaxis_angle = float(cfg.get('anode_xaxis_angle', np.pi / 4))
    p.anode_pitch = float(cfg.get('anode_pitch', 0.5))
    cp = cfg.get('s2_garfield_confine_position', 0.0)
    p.s2_garfield_confine_position = float(cp) if isinstance(cp, float) and cp > 0.0 else 0.0
    return p


def set_resource_scalars(p, cfg, resource):
    """Scalars that depend on resource tables."""
    set_ele_ap_n(p, resource)
    lum = getattr(resource, 's2_luminescence', None) if not isinstance(resource, dict) else resource.get('s2_luminescence')
    if cfg.get('s2_luminescence_model', 'simple') == 'garfield' and lum is not None:
        p.gf_avgt = int(np.average(lum['t']).astype(int))      # s2.py:408


def set_ele_ap_n(p, resource):
    ele = getattr(resource, 'uniform_to_ele_ap', None) if not isinstance(resource, dict) else resource.get('uniform_to_ele_ap')
    if ele is not None:
        p.ele_ap_n = float(ele.n)


def build_tables(cfg, resource=None):
    """`resource` is an object/dict with the optional attributes the reference's Resource
    carries (noise_data, photon_area_distribution or spe_ppf/spe_row, uniform_to_pmt_ap,
    uniform_to_ele_ap); absent pieces leave the matching stage disabled."""
    def get(name, default=None):
        if resource is None:
            return default
        if isinstance(resource, dict):
            return resource.get(name, default)
        return getattr(resource, name, default)

    t = HostTables()
    n_ch = len(cfg['gains'])
    t.set('templates', wtab.pmt_current_templates(cfg), np.float64)
    t.set('gains', cfg['gains'], np.float64)
    t.set('zle_thresholds', wtab.zle_thresholds(cfg), np.int32)
    noise = get('noise_data')
    if cfg.get('enable_noise', True) and noise is not None:
        noise = t.set('noise', noise, np.float64)
        t.struct.noise_len, t.struct.noise_nch = noise.shape
    spe_ppf, spe_row = get('spe_ppf'), get('spe_row')
    if spe_ppf is None and get('photon_area_distribution') is not None:
        spe_ppf, spe_row = wtab.spe_table_from_dataframe(get('photon_area_distribution'), n_ch)
    if spe_ppf is not None:
        spe_ppf = t.set('spe_ppf', spe_ppf, np.float64)
        t.set('spe_row', spe_row, np.int32)
        t.struct.n_spe_rows, t.struct.spe_len = spe_ppf.shape
    # S2 luminescence ('simple' model, constant gas gap)
    if cfg.get('s2_luminescence_model', 'simple') == 'simple' and not cfg.get('enable_gas_gap_warping', False) \
            and 'elr_gas_gap_length' in cfg:
        cdf, tt = wtab.luminescence_table(cfg)
        t.set('lum_cdf', cdf, np.float64)
        t.set('lum_t', tt, np.float64)
        t.struct.lum_len = len(cdf)
    elif cfg.get('s2_luminescence_model', 'simple') == 'simple' and cfg.get('enable_gas_gap_warping', False):
        # per-position gas gaps (resource.gas_gap_length): the device evaluates the field model itself
        sc = wtab.luminescence_field_scalars(cfg)
        for k, v in sc.items():
            setattr(t.struct, 'lumw_' + k, float(v))
    # PMT afterpulse elements (afterpulse.py:155-159, 181-186)
    ap = get('uniform_to_pmt_ap')
    if cfg.get('enable_pmt_afterpulses', True) and ap:
        names = list(ap.keys())
        if len(names) > wlib.MAX_AP_ELEMENTS:
            raise ValueError('too many PMT afterpulse elements')
        t.struct.n_ap_elements = len(names)
        for k, name in enumerate(names):
            el = ap[name]
            dc = np.asarray(el['delaytime_cdf'], dtype=np.float64)
            if dc.ndim != 2 or dc.shape[0] < n_ch:
                raise ValueError(f'afterpulse element {name}: delaytime_cdf must be [n_pmt, n]')
            dc = t.set_elem('ap_delay_cdf', k, dc[:n_ch], np.float64)
            t.struct.ap_delay_len[k] = dc.shape[1]
            t.struct.ap_delay_bin[k] = float(el['delaytime_bin_size'])
            t.struct.ap_is_uniform[k] = int('Uniform' in name)
            ac = np.atleast_1d(np.asarray(el.get('amplitude_cdf', [0.0, 1.0]), dtype=np.float64))
            if ac.ndim == 1:
                ac = ac[None, :]
            else:
                ac = ac[:n_ch]
            ac = t.set_elem('ap_amp_cdf', k, ac, np.float64)
            t.struct.ap_amp_rows[k], t.struct.ap_amp_len[k] = ac.shape
            t.struct.ap_amp_bin[k] = float(el.get('amplitude_bin_size', 1.0))
    # optical propagation grids (s1.py:241-260, s2.py:486-501)
    if 'optical_propagation' in cfg.get('s1_model_type', 'simple'):
        m = get('s1_optical_propagation_spline')
        if m is None or not hasattr(m, 'grid'):
            raise ValueError("s1_model_type 'optical_propagation' needs resource.s1_optical_propagation_spline "
                             "as a wfsim_b200.resource.GridMap with maps 'top' and 'bottom' over (z, U)")
        axes, top = m.grid('top')
        _, bot = m.grid('bottom')
        t.set('s1_op_top', top, np.float64)
        t.set('s1_op_bottom', bot, np.float64)
        t.struct.s1_op_nz, t.struct.s1_op_nu = axes[0][2], axes[1][2]
        t.struct.s1_op_z0, t.struct.s1_op_z1 = axes[0][0], axes[0][1]
        t.struct.s1_op_u0, t.struct.s1_op_u1 = axes[1][0], axes[1][1]
    if 'optical_propagation' in cfg.get('s2_time_model', ''):
        m = get('s2_optical_propagation_spline')
        if m is None or not hasattr(m, 'grid'):
            raise ValueError("s2_time_model 'optical_propagation' needs resource.s2_optical_propagation_spline "
                             "as a wfsim_b200.resource.GridMap with maps 'top' and 'bottom' over (U,)")
        axes, top = m.grid('top')
        _, bot = m.grid('bottom')
        t.set('s2_op_top', top, np.float64)
        t.set('s2_op_bottom', bot, np.float64)
        t.struct.s2_op_nu = axes[0][2]
        t.struct.s2_op_u0, t.struct.s2_op_u1 = axes[0][0], axes[0][1]
    # garfield luminescence table (s2.py:381-409)
    if cfg.get('s2_luminescence_model', 'simple') == 'garfield':
        lum = get('s2_luminescence')
        assert lum is not None, 's2_luminescence model not found'
        tt = np.asarray(lum['t'])
        assert len(tt.shape) == 2, 'Timing data is expected to have D2'
        tt = t.set('gf_t', tt, np.int32)
        t.set('gf_x', lum['x'], np.float64)
        t.struct.gf_rows, t.struct.gf_cols = tt.shape
    # garfield luminescence per gas gap (s2.py:411-483)
    if cfg.get('s2_luminescence_model', 'simple') == 'garfield_gas_gap':
        gg = get('s2_luminescence_gg')
        assert gg is not None, 's2_luminescence_gg model not found'          # s2.py:471
        cdf = t.set('gg_cdf', np.asarray(gg['timing_inv_cdf']), np.float64)
        assert cdf.ndim == 2 and cdf.shape[1] >= 3 and cdf.shape[0] == len(gg['gas_gap'])
        t.struct.gg_rows, t.struct.gg_len = cdf.shape
    # pattern maps on regular grids: evaluated on the device (resource.GridMap with map 'map')
    for key, nd in (('s1', 3), ('s2', 2)):
        m = get(key + '_pattern_map')
        if m is None or not hasattr(m, 'grid') or 'map' not in getattr(m, 'maps', {}):
            continue
        axes, vals = m.grid('map')
        if len(axes) != nd or vals.ndim != nd + 1 or vals.shape[-1] > n_ch or min(a[2] for a in axes) < 2:
            continue            # unusual layout: the host evaluates this map
        if key == 's1' and vals.shape[-1] != n_ch:
            continue
        t.set(key + '_pat_grid', vals, np.float64)
        setattr(t.struct, key + '_pat_npmt', vals.shape[-1])
        for d, (lo, hi, n) in enumerate(axes):
            getattr(t.struct, key + '_pat_n')[d] = n
            getattr(t.struct, key + '_pat_lo')[d] = lo
            getattr(t.struct, key + '_pat_hi')[d] = hi
    # photo-ionisation electrons (afterpulse.py:33-80)
    ele = get('uniform_to_ele_ap')
    if cfg.get('enable_electron_afterpulses', True) and ele is not None:
        coarse = wtab.pi_coarse_grid(cfg, ele.bin_centers)
        prob = wtab.pi_coarse_probabilities(coarse, ele.histogram, ele.bin_edges)
        t.set('pi_coarse_time', coarse, np.float64)
        t.set('pi_coarse_prob', prob, np.float64)
        t.struct.pi_coarse_len = len(coarse)
    return t
