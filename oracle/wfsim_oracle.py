"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the WFSim hot path (numpy + oracle_core.c).

This is a restatement of the reference's algorithm (WFSim v1.2.2; citations are paths relative
to /root/reference), written to be the *checker* for the CUDA path:

* only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
  legs may import it;
* the product package `wfsim_b200` never imports it and has no CPU fallback.

Parity pin: `tests/test_oracle_golden.py` checks this module against the golden vectors the
unmodified reference produced (`tests/golden/make_golden.py`) -- bit-exact for the
deterministic stages -- and, when /root/reference is present, against the reference live.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RECORD_LENGTH = 110


class DigiCfg(ctypes.Structure):
    _fields_ = [('current_2_adc', ctypes.c_double),
                ('trigger_window', ctypes.c_int64),
                ('baseline', ctypes.c_int64),
                ('he_first', ctypes.c_int64),
                ('he_mult', ctypes.c_int64),
                ('n_top', ctypes.c_int64),
                ('n_rows', ctypes.c_int64),
                ('enable_noise', ctypes.c_int64),
                ('noise_len', ctypes.c_int64),
                ('noise_nch', ctypes.c_int64),
                ('ix_rand', ctypes.c_int64)]


def build():
    subprocess.check_call(['make', '-s', '-C', HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(HERE, '_build', 'liboracle.so')
        if not os.path.isfile(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_find_intervals.restype = ctypes.c_int64
        _LIB.orc_pack_records.restype = ctypes.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def raw_record_dtype(samples_per_record=RECORD_LENGTH):
    return np.dtype([('time', np.int64), ('length', np.int32), ('dt', np.int16),
                     ('channel', np.int16), ('pulse_length', np.int32), ('record_i', np.int16),
                     ('baseline', np.int16), ('data', np.int16, samples_per_record)])


# ------------------------------------------------------------------------------------------
# tables
# ------------------------------------------------------------------------------------------
def pmt_current_templates(cfg):
    """wfsim/core/pulse.py:146-187: one template per ns remainder, each summing to 1/dt."""
    ts = np.asarray(cfg['pe_pulse_ts'], dtype=np.float64)
    cdf = np.cumsum(np.asarray(cfg['pe_pulse_ys'], dtype=np.float64))
    dt = cfg.get('sample_duration', 10)
    before = cfg.get('samples_before_pulse_center', 2)
    after = cfg.get('samples_after_pulse_center', 20)
    assert cfg.get('pmt_pulse_time_rounding', 1.0) == 1
    edges = np.linspace(-before * dt, after * dt, 1 + before + after)
    rows = []
    for r in np.arange(0, dt, 1.0):
        # interp1d(bounds_error=False, fill_value=(0, 1)) == np.interp with left=0, right=1
        c = np.interp(edges - r, ts, cdf, left=0.0, right=1.0)
        cur = np.diff(c) / dt
        cur *= (1 / dt) / np.sum(cur)
        rows.append(cur)
    return np.array(rows)


def spe_ppf_table(charge, pdf_columns):
    """wfsim/core/pulse.py:189-223: inverse CDF per column on a 2001-point grid using
    interp1d(kind='next'); `pdf_columns` is what `spe_shapes.columns[1:]` iterates over (note the
    reference's off-by-one: with an unnamed index column the first entry is the charge axis)."""
    grid = np.linspace(0, 1, 2001)
    rows = []
    for pdf in pdf_columns:
        pdf = np.asarray(pdf, dtype=np.float64)
        if pdf.sum() > 0:
            bins = np.asarray(charge, dtype=np.float64)
            cdf = np.cumsum(pdf) / np.sum(pdf)
        else:
            cdf = np.linspace(0, 1, 10)
            bins = np.zeros_like(cdf)
        # 'next' interpolation: value at the first knot >= x; outside -> fill values
        order = np.argsort(cdf, kind='stable')
        xs, ys = cdf[order], bins[order]
        idx = np.searchsorted(xs, grid, side='left')
        out = np.where(idx < len(xs), ys[np.minimum(idx, len(xs) - 1)], bins[-1])
        out = np.where(grid < xs[0], bins[0], out)
        out = np.where(grid > xs[-1], bins[-1], out)
        rows.append(out)
    return np.stack(rows)


def current_2_adc(cfg):
    """wfsim/core/pulse.py:33-35"""
    return (cfg['pmt_circuit_load_resistor'] * cfg['external_amplification']
            / (cfg['digitizer_voltage_range'] / 2 ** cfg['digitizer_bits']))


def zle_thresholds(cfg, n_rows=801):
    """wfsim/core/rawdata.py:290-294"""
    thr = np.full(n_rows, cfg['digitizer_reference_baseline'] - cfg['zle_threshold'] - 1, np.int64)
    for k, v in cfg.get('special_thresholds', {}).items():
        thr[int(k)] = cfg['digitizer_reference_baseline'] - v - 1
    return thr


# ------------------------------------------------------------------------------------------
# deterministic back end
# ------------------------------------------------------------------------------------------
class Pulses:
    """Pulses of one or more Pulse calls: arrays (channel, left, right, offset) + currents."""

    def __init__(self, ch=None, left=None, right=None, off=None, cur=None):
        self.ch = np.zeros(0, np.int32) if ch is None else ch
        self.left = np.zeros(0, np.int64) if left is None else left
        self.right = np.zeros(0, np.int64) if right is None else right
        self.off = np.zeros(0, np.int64) if off is None else off
        self.cur = np.zeros(0, np.float64) if cur is None else cur

    def __len__(self):
        return len(self.ch)

    @staticmethod
    def concat(parts):
        parts = [p for p in parts if len(p)]
        if not parts:
            return Pulses()
        shift = np.cumsum([0] + [len(p.cur) for p in parts[:-1]])
        return Pulses(np.concatenate([p.ch for p in parts]), np.concatenate([p.left for p in parts]),
                      np.concatenate([p.right for p in parts]),
                      np.concatenate([p.off + s for p, s in zip(parts, shift)]),
                      np.concatenate([p.cur for p in parts]))


def pulse_call(cfg, templates, t, ch, gain):
    """wfsim/core/pulse.py:82-144 with the photon gains already decided: per channel pulse
    extents and float64 current.  Photons of ONE Pulse call, any order."""
    dt = cfg.get('sample_duration', 10)
    gains = np.ascontiguousarray(cfg['gains'], np.float64)
    before = int(cfg['samples_to_store_before']) + cfg.get('samples_before_pulse_center', 2)
    after = int(cfg['samples_to_store_after']) + cfg.get('samples_after_pulse_center', 20)
    L = lib()
    t = np.asarray(t, np.int64)
    order = np.lexsort((t, ch))           # by channel, then time (stable)
    t = np.ascontiguousarray(t[order])
    ch = np.ascontiguousarray(np.asarray(ch)[order], np.int32)
    gain = np.ascontiguousarray(np.asarray(gain)[order], np.float64)
    tm = np.ascontiguousarray(templates, dtype=np.float64)
    out_n = np.zeros(2, np.int64)
    args = (ctypes.c_int64(len(t)), _p(t), _p(ch), _p(gain), _p(gains), ctypes.c_int64(dt),
            ctypes.c_int64(before), ctypes.c_int64(after), _p(tm), ctypes.c_int(tm.shape[1]))
    L.orc_pulse_call(*args, None, None, None, None, None, _p(out_n))
    n_p, n_c = int(out_n[0]), int(out_n[1])
    P = Pulses(np.zeros(n_p, np.int32), np.zeros(n_p, np.int64), np.zeros(n_p, np.int64),
               np.zeros(n_p, np.int64), np.zeros(n_c, np.float64))
    if n_p:
        L.orc_pulse_call(*args, _p(P.ch), _p(P.left), _p(P.right), _p(P.off), _p(P.cur), _p(out_n))
    return P


def digitize_zle(cfg, pulses, noise=None, ix_rand=0):
    """rawdata.py:204-311 for one pulse cache -> (itv_ch, itv_left, itv_right, itv_off, samples,
    (group_left, group_right))."""
    L = lib()
    if isinstance(pulses, list):
        pulses = Pulses.concat(pulses)
    n = len(pulses)
    p_ch, p_left, p_right, p_off = (np.ascontiguousarray(a) for a in
                                     (pulses.ch, pulses.left, pulses.right, pulses.off))
    cur = np.ascontiguousarray(pulses.cur, np.float64)
    c = DigiCfg()
    c.current_2_adc = current_2_adc(cfg)
    c.trigger_window = cfg['trigger_window']
    c.baseline = cfg['digitizer_reference_baseline']
    xnt = cfg['detector'] == 'XENONnT'
    c.he_first = cfg['channel_map']['he'][0] if xnt else -1
    c.he_mult = int(cfg['high_energy_deamplification_factor']) if xnt else 0
    c.n_top = cfg['n_top_pmts']
    c.n_rows = 801
    c.enable_noise = int(bool(cfg.get('enable_noise', True)) and noise is not None)
    if noise is not None:
        noise = np.ascontiguousarray(noise, np.float64)
        c.noise_len, c.noise_nch = noise.shape
    c.ix_rand = int(ix_rand)
    thr = zle_thresholds(cfg)
    cap_itv, cap_s = 4096, 1 << 20
    while True:
        itv_ch = np.zeros(cap_itv, np.int32)
        itv_l = np.zeros(cap_itv, np.int64)
        itv_r = np.zeros(cap_itv, np.int64)
        itv_o = np.zeros(cap_itv, np.int64)
        samples = np.zeros(cap_s, np.int16)
        n_itv = ctypes.c_int64()
        n_s = ctypes.c_int64()
        glr = np.zeros(2, np.int64)
        rc = L.orc_digitize_zle(
            ctypes.c_int64(n), _p(p_ch), _p(p_left), _p(p_right), _p(p_off), _p(cur),
            ctypes.byref(c), _p(thr), _p(noise) if noise is not None else None,
            ctypes.c_int64(cap_itv), _p(itv_ch), _p(itv_l), _p(itv_r), _p(itv_o),
            ctypes.c_int64(cap_s), _p(samples), ctypes.byref(n_itv), ctypes.byref(n_s), _p(glr))
        if rc == 0:
            break
        cap_itv = max(cap_itv, n_itv.value)
        cap_s = max(cap_s, n_s.value)
    k = n_itv.value
    return itv_ch[:k], itv_l[:k], itv_r[:k], itv_o[:k], samples[:n_s.value], (int(glr[0]), int(glr[1]))


def pack_records(cfg, itv_ch, itv_l, itv_r, itv_o, samples):
    """strax_interface.py:391-436"""
    L = lib()
    dt = cfg['sample_duration']
    pl = itv_r - itv_l + 1
    n_rec = int(np.sum(np.where(pl > 0, (pl + RECORD_LENGTH - 1) // RECORD_LENGTH, 0)))
    rec = np.zeros(n_rec, raw_record_dtype())
    got = L.orc_pack_records(ctypes.c_int64(len(itv_ch)), _p(np.ascontiguousarray(itv_ch)),
                             _p(np.ascontiguousarray(itv_l)), _p(np.ascontiguousarray(itv_r)),
                             _p(np.ascontiguousarray(itv_o)), _p(np.ascontiguousarray(samples)),
                             ctypes.c_int64(dt), ctypes.c_int64(RECORD_LENGTH), _p(rec),
                             ctypes.c_int64(n_rec))
    assert got == n_rec
    return rec


def sort_by_time(x):
    """strax.sort_by_time (third party): stable sort on (time, channel)."""
    if len(x) == 0:
        return x
    key = (x['time'] - x['time'].min()) * (int(x['channel'].max()) + 1) + x['channel']
    return x[np.argsort(key, kind='mergesort')]


def simulate_photons(cfg, pcall, ch, t, gain, group_of, noise=None, ix_rand=None,
                     templates=None):
    """Deterministic leg end to end: photons -> per-pulse-call currents -> per-group digitise +
    ZLE -> records sorted by (time, channel), split tpc / he as strax_interface.py:489-494."""
    if templates is None:
        templates = pmt_current_templates(cfg)
    n_groups = int(group_of.max()) + 1 if len(group_of) else 0
    recs = []
    groups_lr = []
    for grp in range(n_groups):
        cache = Pulses.concat([pulse_call(cfg, templates, t[pcall == pc], ch[pcall == pc],
                                          gain[pcall == pc])
                               for pc in np.flatnonzero(group_of == grp)])
        if not len(cache):
            groups_lr.append(None)
            continue
        ir = 0 if ix_rand is None else ix_rand[grp]
        out = digitize_zle(cfg, cache, noise=noise, ix_rand=ir)
        groups_lr.append(out[5])
        recs.append(pack_records(cfg, *out[:5]))
    rec = np.concatenate(recs) if recs else np.zeros(0, raw_record_dtype())
    rec = sort_by_time(rec)
    he0, he1 = cfg['channel_map']['he'][0], cfg['channel_map']['he'][-1]
    return dict(raw_records=rec[rec['channel'] < he0],
                raw_records_he=rec[(rec['channel'] >= he0) & (rec['channel'] <= he1)],
                raw_records_aqmon=rec[rec['channel'] == 800],
                groups_lr=groups_lr)
