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
    return p


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
    return t
