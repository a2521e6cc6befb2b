"""Host-side table construction for the device kernels (runs once per config at setup()).

These follow the reference's definitions so that the same fax_config produces the same tables:
* PMT current templates        wfsim/core/pulse.py:146-187
* SPE inverse-CDF table        wfsim/core/pulse.py:189-227
* ZLE thresholds per row       wfsim/core/rawdata.py:290-294
* S2 'simple' luminescence     wfsim/core/s2.py:317-378 (tabulated once for the constant gas gap)
* photo-ionisation tables      wfsim/core/afterpulse.py:33-80
"""
import numpy as np

N_ROWS = 801          # rows of the reference's dense _raw_data (rawdata.py:224)
BOLTZMANN = 8.617343e-5   # eV/K   (wfsim/units.py pax unit system: eV = 1, K = 1)
# pax unit system (wfsim/units.py): base units ns=1, cm=1(?), V=1, e=1 ...
# only the combination used by the luminescence model is needed, see luminescence_table().


def pmt_current_templates(cfg):
    ts = np.asarray(cfg['pe_pulse_ts'], dtype=np.float64)
    cdf = np.cumsum(np.asarray(cfg['pe_pulse_ys'], dtype=np.float64))
    dt = cfg.get('sample_duration', 10)
    before = cfg.get('samples_before_pulse_center', 2)
    after = cfg.get('samples_after_pulse_center', 20)
    if cfg.get('pmt_pulse_time_rounding', 1.0) != 1:
        raise AssertionError('pmt_pulse_time_rounding must be 1')
    edges = np.linspace(-before * dt, after * dt, 1 + before + after)
    shifts = np.arange(0, dt, 1.0)
    # CDF evaluated at (edge - shift); outside the tabulated range the CDF is 0 / 1
    c = np.interp(edges[None, :] - shifts[:, None], ts, cdf, left=0.0, right=1.0)
    cur = np.diff(c, axis=1) / dt
    out = np.empty_like(cur)
    for i in range(cur.shape[0]):          # row-wise to keep the reference's operation order
        row = cur[i].copy()
        row *= (1 / dt) / np.sum(row)
        out[i] = row
    return out


def spe_ppf_rows(charge, pdfs):
    """Inverse CDF on the 2001-point uniform grid, 'next'-neighbour interpolation.
    `pdfs`: iterable of per-row pdf arrays.  Returns (unique_rows [k, 2001], row_index [n])."""
    grid = np.linspace(0, 1, 2001)
    cache, rows, index = {}, [], []
    charge = np.asarray(charge, dtype=np.float64)
    for pdf in pdfs:
        pdf = np.asarray(pdf, dtype=np.float64)
        key = pdf.tobytes()
        if key not in cache:
            if pdf.sum() > 0:
                bins, cdf = charge, np.cumsum(pdf) / np.sum(pdf)
            else:
                cdf = np.linspace(0, 1, 10)
                bins = np.zeros_like(cdf)
            order = np.argsort(cdf, kind='stable')
            xs, ys = cdf[order], bins[order]
            idx = np.searchsorted(xs, grid, side='left')
            val = ys[np.minimum(idx, len(xs) - 1)]
            val = np.where(idx >= len(xs), bins[-1], val)
            val = np.where(grid < xs[0], bins[0], val)
            cache[key] = len(rows)
            rows.append(val)
        index.append(cache[key])
    return np.stack(rows), np.asarray(index, dtype=np.int32)


def spe_table_from_dataframe(df, n_channels):
    """Row r of the reference table is built from column position r+1 of the csv (the
    reference iterates `columns[1:]`; with an unnamed index column the first row is the charge
    axis itself -- reproduced, pulse.py:201)."""
    cols = list(df.columns[1:])
    charge = df['charge'].values
    uniq, index = spe_ppf_rows(charge, [df[c].values for c in cols])
    if len(index) < n_channels:
        raise ValueError(f'SPE table has {len(index)} rows, need {n_channels}')
    return uniq, index[:n_channels].copy()


def zle_thresholds(cfg, n_rows=N_ROWS):
    thr = np.full(n_rows, cfg['digitizer_reference_baseline'] - cfg['zle_threshold'] - 1,
                  dtype=np.int32)
    for k, v in (cfg.get('special_thresholds') or {}).items():
        if 0 <= int(k) < n_rows:
            thr[int(k)] = cfg['digitizer_reference_baseline'] - v - 1
    return thr
