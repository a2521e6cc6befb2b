"""fax_config handling for the B200 hot path (schema unchanged from the reference).

* `load_fax_config`  -- tolerant JSON reader (the shipped fax configs carry `//` and `#`
  comments and trailing commas; straxen reads them with commentjson, which is third party).
* `plugin_config`    -- the derived keys `SimulatorPlugin.set_config` adds before the simulator
  sees the dict (wfsim/strax_interface.py:566-608): `gains`, `channel_map['sum_signal']`,
  `channels_bottom`, `field_distortion_model`, plus the strax.Option defaults
  (wfsim/strax_interface.py:506-535).
"""
import json
import re

import numpy as np

N_TPC_PMTS = 494
N_TOP_PMTS = 253

# straxen.contexts.xnt_common_config['channel_map'] (third party; published XENONnT layout)
XNT_CHANNEL_MAP = dict(
    tpc=(0, 493), he=(500, 752), aqmon=(790, 807), aqmon_nv=(808, 815),
    tpc_blank=(999, 999), mv=(1000, 1083), aux_mv=(1084, 1087), mv_blank=(1999, 1999),
    nveto=(2000, 2119), nveto_blank=(2999, 2999))

PLUGIN_OPTION_DEFAULTS = dict(
    detector='XENONnT', event_rate=1000, chunk_size=100, n_chunk=10, per_pmt_truth=False,
    fax_file=None, fax_config_override=None, fax_config_override_from_cmt=None,
    right_raw_extension=100000, seed=False)


def _strip_comments(text):
    out = []
    for line in text.splitlines():
        res = []
        in_str = False
        i = 0
        while i < len(line):
            c = line[i]
            if c == '"' and (i == 0 or line[i - 1] != '\\'):
                in_str = not in_str
            if not in_str and (c == '#' or line.startswith('//', i)):
                break
            res.append(c)
            i += 1
        out.append(''.join(res))
    return '\n'.join(out)


def loads_tolerant(text):
    text = _strip_comments(text)
    text = re.sub(r',(\s*[\]\}])', r'\1', text)
    return json.loads(text)


def load_fax_config(path):
    with open(path) as f:
        return loads_tolerant(f.read())


def plugin_config(fax_config, overrides=None, to_pe=None, **options):
    """Build the dict the simulator consumes, as `SimulatorPlugin.set_config` would."""
    cfg = dict(PLUGIN_OPTION_DEFAULTS)
    cfg.update(options)
    cfg.update(fax_config)
    if overrides:
        cfg.update(overrides)
    if 'field_distortion_on' in cfg and 'field_distortion_model' not in cfg:
        cfg['field_distortion_model'] = 'inverse_fdc' if cfg['field_distortion_on'] else 'none'
    cfg.setdefault('n_tpc_pmts', N_TPC_PMTS)
    cfg.setdefault('n_top_pmts', N_TOP_PMTS)
    if to_pe is None:
        to_pe = np.full(cfg['n_tpc_pmts'], 0.008)
    to_pe = np.asarray(to_pe, dtype=np.float64)
    adc_2_current = (cfg['digitizer_voltage_range'] / 2 ** cfg['digitizer_bits']
                     / cfg['pmt_circuit_load_resistor'])
    gains = np.zeros_like(to_pe)
    np.divide(adc_2_current, to_pe, out=gains, where=to_pe != 0)
    cfg['gains'] = gains
    cmap = dict(cfg.get('channel_map') or XNT_CHANNEL_MAP)
    cmap['sum_signal'] = 800
    cfg['channel_map'] = cmap
    cfg['channels_bottom'] = np.arange(cfg['n_top_pmts'], cfg['n_tpc_pmts'])
    return cfg


def current_2_adc(cfg):
    """wfsim/core/pulse.py:33-35."""
    return (cfg['pmt_circuit_load_resistor'] * cfg['external_amplification']
            / (cfg['digitizer_voltage_range'] / 2 ** cfg['digitizer_bits']))
