#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE UNMODIFIED REFERENCE (WFSim v1.2.2).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.json, *.npz

The reference's own tests pin no numbers, so these vectors are the pin:
inputs + the outputs the reference's code produced for them.  Fixtures:

  c0_config.json      merged fax_config for BASELINE config[0]: files/XENONnT_wfsim_config.json
                      + the overrides of tests/test_load_resource.py:22-44 (plugin-derived keys
                      are re-derived by wfsim_b200.config.plugin_config, to_pe = 0.008).
  c0_tables.npz       tables the reference derives from that config: `_pmt_current_templates`
                      (pulse.py:146-187), SPE inverse-CDF table (pulse.py:189-223; SPE csv widened
                      to 494 identical columns as tests/test_wfsim.py:83-88 does).
  det_*.npz           deterministic leg: photons (pulse-call id, channel, t_ns, gain) + group ids
                      -> reference Pulse.__call__/add_current -> digitize_pulse_cache -> ZLE ->
                      reference ChunkRawRecords record packing.
  stoch_*.npz         samples drawn from the reference's stochastic stage functions, for the
                      two-sample KS / chi-square tests of the Philox kernels.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader as RL  # noqa: E402
from wfsim_b200 import config as wcfg  # noqa: E402

# tests/test_load_resource.py:22-44 (values restated; 'gains' is overwritten by the plugin)
TEST_LOAD_NT_OVERRIDES = {
    "detector": "XENONnT",
    "s2_luminescence_model": "simple",
    "enable_gas_gap_warping": False,
    "enable_pmt_afterpulses": False,
    "enable_electron_afterpulses": False,
    "enable_noise": False,
    "field_distortion_on": False,
    "enable_field_dependencies": {
        "survival_probability_map": True, "drift_speed_map": False,
        "diffusion_longitudinal_map": False, "diffusion_transverse_map": False},
    "photon_area_distribution": "XENONnT_spe_distributions_single_channel.csv",
    "s1_pattern_map": ["constant dummy", 14e-5, [494]],
    "s1_lce_correction_map": ["constant dummy", 1, []],
    "s2_pattern_map": ["constant dummy", 30e-5, [494]],
    "s2_correction_map": ["constant dummy", 1, []],
    "field_dependencies_map": ["constant dummy", 1, []],
    "se_gain_map": ["constant dummy", 1, []],
}


def widened_spe_hook(path, fmt):
    """Serve the single-channel SPE csv widened to 494 columns (tests/test_wfsim.py:83-88)."""
    import pandas as pd
    if fmt == 'csv' and 'spe_distributions' in path:
        df = pd.read_csv(os.path.join(RL.REFERENCE_ROOT, 'files',
                                      'XENONnT_spe_distributions_single_channel.csv'))
        cols = {str(i): df['0'] for i in range(1, 494)}
        df = pd.concat([df, pd.DataFrame(cols)], axis=1)
        return df
    if fmt == 'json.gz' and 'synthetic_ap' in path:
        from tests.golden.synth_tables import pmt_ap_tables
        return pmt_ap_tables()
    if fmt == 'dill' and 'synthetic_ele_ap' in path:
        from tests.golden.synth_tables import EleApHist
        return EleApHist()
    return None


def c0_config(**extra):
    fax = wcfg.load_fax_config(os.path.join(RL.REFERENCE_ROOT, 'files', 'XENONnT_wfsim_config.json'))
    ov = dict(TEST_LOAD_NT_OVERRIDES)
    ov.update(extra)
    return wcfg.plugin_config(fax, overrides=ov, to_pe=np.full(494, 0.008)), fax, ov


def jsonable(o):
    if isinstance(o, dict):
        return {k: jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, np.ndarray):
        return o.tolist()
    if isinstance(o, np.generic):
        return o.item()
    return o


# ----------------------------------------------------------------------------------------
# deterministic leg
# ----------------------------------------------------------------------------------------
from tests.golden.synth import synth_photons  # noqa: E402


def run_reference_deterministic(ref, cfg, pcall, ch, t, g, group_of, noise=None, noise_seed=None):
    """Feed preset photons through the reference's Pulse/RawData/ChunkRawRecords code."""
    import numba
    n_groups = int(group_of.max()) + 1
    dt = cfg['sample_duration']

    @numba.njit
    def _seed(s):
        np.random.seed(s)

    @numba.njit
    def _draw(high):
        return np.random.randint(0, high)

    class PresetRawData(ref.RawData):
        """RawData whose generator replays preset photons group by group, using the
        reference's own Pulse.__call__, digitize_pulse_cache and ZLE."""

        def __init__(self, config):
            self.config = config
            self.resource = ref.load_resource.load_config(config)
            self.pulse = ref.Pulse(config)
            self.ix_rand = []

        def __call__(self, instructions=None, truth_buffer=None, **kw):
            self.source_finished = False
            self._pulses_cache = []
            for grp in range(n_groups):
                for pc in np.where(group_of == grp)[0]:
                    m = pcall == pc
                    order = np.argsort(ch[m], kind='stable')
                    self.pulse._photon_timings = t[m][order].copy()
                    self.pulse._photon_channels = ch[m][order].astype(np.int64)
                    self.pulse._photon_gains = g[m][order].copy()
                    self.pulse()
                    self._pulses_cache += self.pulse._pulses
                if noise is not None:
                    lo = min(p['left'] for p in self._pulses_cache)
                    hi = max(p['right'] for p in self._pulses_cache)
                    span = hi - lo + 2 * cfg['trigger_window']
                    high = len(noise) - span - 1
                    if high < 0:
                        high = len(noise) - 1
                    _seed(noise_seed + grp)
                    self.ix_rand.append(int(_draw(high)) if high > 0 else 0)
                    _seed(noise_seed + grp)
                self.digitize_pulse_cache()
                yield from self.ZLE()
            self.source_finished = True

    if noise is not None:
        res = ref.load_resource.load_config(cfg)
        res.noise_data = noise
    crr = ref.ChunkRawRecords(cfg, rawdata_generator=PresetRawData)
    crr.record_buffer = np.zeros(400000, dtype=crr.record_buffer.dtype)
    fake_instr = np.zeros(1, dtype=ref.strax_interface.instruction_dtype)
    fake_instr['time'] = t.min()
    cfg_cs = dict(cfg)
    outs = list(crr(fake_instr))
    rr = np.concatenate([o['raw_records'] for o in outs])
    rr_he = np.concatenate([o['raw_records_he'] for o in outs])
    rr_aq = np.concatenate([o['raw_records_aqmon'] for o in outs])
    return rr, rr_he, rr_aq, np.asarray(crr.rawdata.ix_rand, np.int64)


def make_det_case(ref, name, seed, n_groups, cfg_extra=None, with_noise=False, big=False):
    rng = np.random.default_rng(seed)
    cfg, _, _ = c0_config(**(cfg_extra or {}))
    cfg['chunk_size'] = 10000  # one chunk: chunk cutting is tested separately
    # a few dead PMTs (gain 0 -> pulse skipped, pulse.py:89-90)
    gains = cfg['gains'].copy()
    gains[[3, 100, 300]] = 0
    cfg['gains'] = gains
    pcall, ch, t, g, group_of = synth_photons(cfg, rng, n_groups, big=big)
    noise = None
    if with_noise:
        cfg['enable_noise'] = True
        noise = rng.integers(-12, 13, (2000, 494)) / 2.0   # half-integers: exercises trunc; short: exercises wrap
    rr, rr_he, rr_aq, ix_rand = run_reference_deterministic(
        ref, cfg, pcall, ch, t, g, group_of, noise=noise, noise_seed=seed * 7 + 1)
    print(f'{name}: {len(t)} photons, {group_of.max() + 1} groups, {len(group_of)} pulse calls -> '
          f'{len(rr)} records, {len(rr_he)} he, {len(rr_aq)} aqmon')
    out = dict(pcall=pcall, channel=ch, t=t, gain=g, group_of=group_of,
               gains=np.asarray(cfg['gains']), ix_rand=ix_rand,
               cfg_extra=json.dumps(jsonable(cfg_extra or {})),
               rr=rr.view(np.uint8), rr_he=rr_he.view(np.uint8))
    if noise is not None:
        out['noise_x2'] = np.round(noise * 2).astype(np.int8)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)


def main():
    ref = RL.load_reference(resource_hook=widened_spe_hook)
    cfg, fax, ov = c0_config()
    merged = dict(fax)
    merged.update(ov)
    with open(os.path.join(HERE, 'c0_config.json'), 'w') as f:
        json.dump(jsonable(merged), f, indent=0, sort_keys=True)
    pulse = ref.Pulse(dict(cfg))
    templates = np.asarray(pulse._pmt_current_templates)
    spe = pulse._Pulse__uniform_to_pe_arr
    uniq, inverse = np.unique(spe, axis=0, return_inverse=True)
    print('templates', templates.shape, 'spe', spe.shape, 'unique rows', uniq.shape)
    np.savez_compressed(os.path.join(HERE, 'c0_tables.npz'),
                        templates=templates, spe_unique=uniq, spe_row=inverse.astype(np.int32),
                        current_max=pulse.current_max, current_2_adc=pulse.current_2_adc)

    which = sys.argv[1:] or ['det']
    if 'det' in which:
        make_det_case(ref, 'det_basic', seed=11, n_groups=5)
        make_det_case(ref, 'det_he_thr', seed=12, n_groups=3,
                      cfg_extra={'high_energy_deamplification_factor': 2.0,
                                 'special_thresholds': {'7': 40, '255': 5},
                                 'zle_threshold': 10})
        make_det_case(ref, 'det_noise', seed=13, n_groups=3, with_noise=True)
    if 'ap' in which:
        from tests.golden import make_golden_ap
        make_golden_ap.main(ref, c0_config)
    if 'chunks' in which:
        from tests.golden import make_golden_chunks
        make_golden_chunks.main(ref, c0_config)
    if 'sched' in which:
        from tests.golden import make_golden_sched
        make_golden_sched.main(ref, c0_config)
    if 'stoch' in which:
        from tests.golden import make_golden_stoch
        make_golden_stoch.main(ref, c0_config)
    if 'models' in which:
        from tests.golden import make_golden_models
        make_golden_models.main(ref, c0_config)
    if 'gg' in which:
        from tests.golden import make_golden_gg
        make_golden_gg.main(ref, c0_config)
    if 'lumw' in which:
        from tests.golden import make_golden_lumw
        make_golden_lumw.main(ref, c0_config)
    if 'diffuse' in which:
        from tests.golden import make_golden_diffuse
        make_golden_diffuse.main(ref, c0_config)


if __name__ == '__main__':
    main()
