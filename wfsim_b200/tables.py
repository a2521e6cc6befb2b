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


# pax unit system constants used by the S2 luminescence model (wfsim/units.py: cm, ns, eV, V)
_ELECTRON_CHARGE_SI = 1.602176565e-19
_BOLTZMANN = 1.3806488e-23 / _ELECTRON_CHARGE_SI      # eV / K
_KV_PER_CM = 1000.0
_BAR = 1e5 / _ELECTRON_CHARGE_SI / 100.0 / 100.0 ** 2


def luminescence_table(cfg, gas_gap=None):
    """Inverse-CDF table of the 'simple' S2 luminescence model for one gas gap.

    The reference rebuilds these arrays for every instruction (s2.py:317-341); without gas-gap
    warping the gap is a constant so one table serves all: emission = interp(U, cdf, t)."""
    dG = float(cfg['elr_gas_gap_length'] if gas_gap is None else gas_gap)
    number_density_gas = cfg['pressure'] / (_BOLTZMANN * cfg['temperature'])
    alpha = cfg['gas_drift_velocity_slope'] / number_density_gas
    pressure = cfg['pressure'] / _BAR
    rA = cfg['anode_field_domination_distance']
    rW = cfg['anode_wire_radius']
    dL = cfg['gate_to_anode_distance'] - dG
    VG = cfg['anode_voltage'] / (1 + dL / dG / cfg['lxe_dielectric_constant'])
    E0 = VG / ((dG - rA) / rA + np.log(rA / rW))
    dr = 0.0001
    r = np.arange(dG, rW, -dr)
    rr = np.clip(1 / r, 1 / rA, 1 / rW)
    dt = dr / (alpha * E0 * rr)
    dy = E0 * rr / _KV_PER_CM - 0.8 * pressure
    avgt = np.sum(np.cumsum(dt) * dy) / np.sum(dy)
    j = int(np.argmax(r <= dG))
    t = np.cumsum(dt[j:]) - avgt
    y = np.cumsum(dy[j:])
    return y / y[-1], t


def luminescence_field_scalars(cfg):
    """Scalars of the 'simple' luminescence field model (s2.py:355-373) that do not depend on the gas gap."""
    number_density_gas = cfg['pressure'] / (_BOLTZMANN * cfg['temperature'])
    return dict(alpha=cfg['gas_drift_velocity_slope'] / number_density_gas, ue=_KV_PER_CM,
                pressure=cfg['pressure'] / _BAR, ra=cfg['anode_field_domination_distance'],
                rw=cfg['anode_wire_radius'], dr=0.0001)


def luminescence_field_scale(cfg, gas_gap):
    """E0 [V/cm] of s2.py:365-370 for an array of gas gaps [cm]."""
    dG = np.asarray(gas_gap, dtype=np.float64)
    rA, rW = cfg['anode_field_domination_distance'], cfg['anode_wire_radius']
    dL = cfg['gate_to_anode_distance'] - dG
    VG = cfg['anode_voltage'] / (1 + dL / dG / cfg['lxe_dielectric_constant'])
    return VG / ((dG - rA) / rA + np.log(rA / rW))


def template_maxima(templates):
    """current_max of pulse.py:32."""
    return np.max(np.asarray(templates), axis=1)


def pi_coarse_grid(cfg, bin_centers):
    """Coarse delay grid of PhotoIonization_Electron._reduce_instruction_timing
    (afterpulse.py:63-75)."""
    bin_centers = np.asarray(bin_centers, dtype=np.float64)
    spread = np.sqrt(2 * cfg['diffusion_constant_longitudinal'] * bin_centers)
    spread /= cfg['drift_velocity_liquid']
    coarse, cur = [], 100.0
    while cur < bin_centers[-1]:
        coarse.append(cur)
        cur += spread[np.argmin(np.abs(cur - bin_centers))]
    return np.array(coarse)


def pi_coarse_probabilities(coarse, histogram, bin_edges):
    """Probability that one delay drawn from the (piecewise-uniform) delay histogram
    (multihist.Hist1d.get_random: pick a bin by weight, uniform inside it) lands in each coarse
    bin of np.digitize(delay, coarse) -- i.e. index i collects coarse[i-1] <= delay < coarse[i],
    index 0 collects delay < coarse[0]; delays >= coarse[-1] are dropped (afterpulse.py:77)."""
    histogram = np.asarray(histogram, dtype=np.float64)
    bin_edges = np.asarray(bin_edges, dtype=np.float64)
    w = histogram / histogram.sum()
    cum = np.concatenate([[0.0], np.cumsum(w)])

    def cdf(x):
        x = np.clip(x, bin_edges[0], bin_edges[-1])
        k = np.clip(np.searchsorted(bin_edges, x, side='right') - 1, 0, len(w) - 1)
        frac = (x - bin_edges[k]) / (bin_edges[k + 1] - bin_edges[k])
        return cum[k] + w[k] * frac
    edges = np.concatenate([[-np.inf], coarse])
    c = cdf(np.where(np.isinf(edges), bin_edges[0], edges))
    c[0] = 0.0
    return np.diff(c)
