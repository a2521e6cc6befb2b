"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the *stochastic* front end, the event scheduler, the
truth rows and the chunker of the WFSim hot path (numpy; inner loops in oracle_core.c).

Restates (WFSim v1.2.2, paths relative to /root/reference):
  S1           wfsim/core/s1.py:60-238
  S2           wfsim/core/s2.py:73-136,139-315,343-378,504-557,616-682
  PMT stage    wfsim/core/pulse.py:39-144,225-271,321-341
  afterpulses  wfsim/core/afterpulse.py:14-88,143-249
  scheduler    wfsim/core/rawdata.py:38-202,313-375
  chunker      wfsim/strax_interface.py:368-497
Random numbers come from one numpy Generator, so -- exactly like the reference -- results are
comparable to the CUDA path (Philox) only statistically.  The deterministic back end
(oracle/wfsim_oracle.py) is shared and bit-exact.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module.
"""
import numpy as np

from . import wfsim_oracle as det

INSTR_FIELDS = ('event_number', 'type', 'time', 'x', 'y', 'z', 'amp', 'recoil', 'e_dep', 'tot_e',
                'g4id', 'vol_id', 'local_field', 'n_excitons', 'x_pri', 'y_pri', 'z_pri')


class ConstMap:
    """["constant dummy", value, shape] maps (load_resource.py:438-457)."""

    def __init__(self, const, shape=()):
        self.const, self.shape = const, tuple(shape)

    def __call__(self, x, **kw):
        return np.ones([len(x)] + list(self.shape)) * self.const


def maps_from_config(cfg):
    """Constant maps named in the config; anything else must be supplied by the caller."""
    out = {}
    for k in ('s1_pattern_map', 's2_pattern_map', 's1_lce_correction_map', 's2_correction_map',
              'se_gain_map', 'field_dependencies_map'):
        v = cfg.get(k)
        if isinstance(v, (list, tuple)) and v and v[0] == 'constant dummy':
            out[k] = ConstMap(v[1], v[2])
    return out


def luminescence_table(cfg):
    """s2.py:317-378 for the constant gas gap (no warping): (cdf, t) with emission time
    = interp(U, cdf, t).  Unit constants: wfsim/units.py (cm, ns, eV, V, e)."""
    e_si = 1.602176565e-19
    kb = 1.3806488e-23 / e_si
    bar = 1e5 / e_si / 100.0 / 100.0 ** 2
    dG = cfg['elr_gas_gap_length']
    n_gas = cfg['pressure'] / (kb * cfg['temperature'])
    alpha = cfg['gas_drift_velocity_slope'] / n_gas
    pressure = cfg['pressure'] / bar
    rA, rW = cfg['anode_field_domination_distance'], cfg['anode_wire_radius']
    dL = cfg['gate_to_anode_distance'] - dG
    VG = cfg['anode_voltage'] / (1 + dL / dG / cfg['lxe_dielectric_constant'])
    E0 = VG / ((dG - rA) / rA + np.log(rA / rW))
    dr = 0.0001
    r = np.arange(dG, rW, -dr)
    rr = np.clip(1 / r, 1 / rA, 1 / rW)
    dt = dr / (alpha * E0 * rr)
    dy = E0 * rr / 1000.0 - 0.8 * pressure
    avgt = np.sum(np.cumsum(dt) * dy) / np.sum(dy)
    j = int(np.argmax(r <= dG))
    t = np.cumsum(dt[j:]) - avgt
    y = np.cumsum(dy[j:])
    return y / y[-1], t


def _trunc(a):
    return np.asarray(a).astype(np.int64)      # float -> int64 truncates toward zero


def _choice_rows(rng, pattern, counts):
    """Categorical channel draw per row with `counts[i]` samples from pattern[i] (already
    normalised): same distribution as np.random.choice(p=...) per instruction."""
    out = []
    for p, n in zip(pattern, counts):
        if n == 0:
            continue
        if np.isnan(p).any():
            out.append(np.full(n, -1, np.int64))
            continue
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        out.append(np.searchsorted(cdf, rng.random(n), side='right').astype(np.int64))
    return np.concatenate(out) if out else np.zeros(0, np.int64)


class OracleSimulator:
    def __init__(self, cfg, spe_table, maps=None, noise=None, pmt_ap=None, ele_ap=None, seed=0):
        """spe_table: [n_ch, 2001] inverse-CDF rows by channel (pulse.py:189-227)."""
        self.cfg = cfg
        self.rng = np.random.default_rng(seed)
        self.spe = np.asarray(spe_table)
        self.maps = maps_from_config(cfg)
        if maps:
            self.maps.update(maps)
        self.noise = noise
        self.pmt_ap = pmt_ap
        self.ele_ap = ele_ap
        self.templates = det.pmt_current_templates(cfg)
        self.current_max = self.templates.max(axis=1)
        self.c2a = det.current_2_adc(cfg)
        self.gains = np.asarray(cfg['gains'], np.float64)
        self.dead = self.gains == 0
        self.n_ch = len(self.gains)
        self.dt = cfg.get('sample_duration', 10)
        if cfg.get('s2_luminescence_model', 'simple') == 'simple':
            self.lum_cdf, self.lum_t = luminescence_table(cfg)
        self.n_photons_made = 0

    # ---- S1 ---------------------------------------------------------------------------------
    def s1_hits(self, instr):
        xyz = np.stack([instr['x'], instr['y'], instr['z']], axis=1).astype(np.float64)
        ly = np.asarray(self.maps['s1_lce_correction_map'](xyz), np.float64)
        if ly.ndim != 1:
            ly = np.squeeze(ly, axis=-1)
        ly = ly / (1 + self.cfg['p_double_pe_emision']) * self.cfg['s1_detection_efficiency']
        return self.rng.binomial(instr['amp'], ly)

    def s1_photons(self, instr):
        cfg, rng = self.cfg, self.rng
        hits = self.s1_hits(instr)
        xyz = np.stack([instr['x'], instr['y'], instr['z']], axis=1).astype(np.float64)
        pat = np.array(self.maps['s1_pattern_map'](xyz), np.float64)
        pat[:, self.dead] = 0
        pat = pat / pat.sum(axis=1, keepdims=True)
        ch = _choice_rows(rng, pat, hits)
        t = np.repeat(instr['time'].astype(np.int64), hits)
        n = len(t)
        if n and 'simple' in cfg['s1_model_type']:
            t = t + _trunc(rng.exponential(cfg['s1_decay_time'], n))
            t = t + _trunc(rng.normal(0, cfg['s1_decay_spread'], n))
        return t, ch

    # ---- S2 ---------------------------------------------------------------------------------
    def drift_params(self, z):
        cfg = self.cfg
        mean = np.clip(-z / cfg['drift_velocity_liquid'] + cfg['drift_time_gate'], 0, np.inf)
        spread = np.sqrt(2 * cfg['diffusion_constant_longitudinal'] * mean) / cfg['drift_velocity_liquid']
        return mean, spread

    def s2_sc_gain(self, xy):
        cfg = self.cfg
        if cfg.get('se_gain_from_map', False):
            g = np.asarray(self.maps['se_gain_map'](xy), np.float64)
        else:
            g = np.asarray(self.maps['s2_correction_map'](xy), np.float64) * cfg['s2_secondary_sc_gain']
        if g.ndim != 1:
            g = np.squeeze(g, axis=-1)
        g = g / (1 + cfg['p_double_pe_emision'])
        g[np.isnan(g)] = 0
        return g

    def s2_electrons(self, instr):
        cfg, rng = self.cfg, self.rng
        z = instr['z'].astype(np.float64)
        xy = np.stack([instr['x'], instr['y']], axis=1).astype(np.float64)
        mean, spread = self.drift_params(z)
        cy = cfg['electron_extraction_yield'] * np.exp(-mean / cfg['electron_lifetime_liquid'])
        if cfg['enable_field_dependencies']['survival_probability_map']:
            r = np.sqrt(xy[:, 0] ** 2 + xy[:, 1] ** 2)
            ps = np.asarray(self.maps['field_dependencies_map'](
                np.array([r, z]).T, map_name='survival_probability_map'), np.float64).reshape(-1)
            cy = cy * np.clip(ps, 0, 1)
        n_e = rng.binomial(instr['amp'], np.clip(cy, 0, 1))
        tot = int(n_e.sum())
        te = rng.exponential(cfg['electron_trapping_time'], tot) + \
            rng.normal(np.repeat(mean, n_e), np.repeat(spread, n_e))
        te = np.repeat(instr['time'].astype(np.int64), n_e) + _trunc(te)
        g = np.repeat(self.s2_sc_gain(xy), n_e)
        nph = rng.poisson(g)
        sp = cfg.get('s2_gain_spread', 0)
        if sp:
            nph = nph + _trunc(rng.normal(0, sp, tot))
        nph[nph < 0] = 0
        return n_e, te, nph, xy

    def s2_photons(self, instr):
        cfg, rng = self.cfg, self.rng
        n_e, te, nph, xy = self.s2_electrons(instr)
        per_instr = np.add.reduceat(nph, np.concatenate([[0], np.cumsum(n_e)[:-1]])) if len(nph) else \
            np.zeros(len(instr), np.int64)
        per_instr = np.where(n_e > 0, per_instr, 0)
        pat = np.array(self.maps['s2_pattern_map'](xy), np.float64)
        if pat.shape[1] < self.n_ch:
            pat = np.pad(pat, [[0, 0], [0, self.n_ch - pat.shape[1]]], 'constant', constant_values=1)
        pat[:, self.dead] = 0
        s = pat.sum(axis=1, keepdims=True)
        pat = np.divide(pat, s, out=np.zeros_like(pat), where=s != 0)
        ch = _choice_rows(rng, pat, per_instr)
        n = int(nph.sum())
        if cfg['s2_luminescence_model'] != 'simple':
            raise NotImplementedError('oracle: only the simple luminescence model')
        t = _trunc(np.interp(rng.random(n), self.lum_cdf, self.lum_t))
        delay = np.where(rng.random(n) < cfg['singlet_fraction_gas'],
                         cfg['singlet_lifetime_gas'], cfg['triplet_lifetime_gas'])
        t = t + _trunc(rng.exponential(1, n) * delay)
        model = cfg['s2_time_model']
        if 'optical_propagation' in model:
            raise NotImplementedError('oracle: s2 optical propagation needs the spline map')
        elif 'zero_delay' in model:
            pass
        elif 's2_time_spread around zero' in model:
            t = t + _trunc(rng.normal(0, cfg['s2_time_spread'], n))
        else:
            raise KeyError(model)
        self.last_te_per_photon = np.repeat(te, nph)
        t = t + self.last_te_per_photon
        return t, ch, te

    # ---- PMT stage --------------------------------------------------------------------------
    def pmt_stage(self, t, ch):
        """TTS, DPE flags, SPE gains (pulse.py:53-56,76-79,95-103)."""
        cfg, rng = self.cfg, self.rng
        n = len(t)
        t = t + _trunc(rng.normal(cfg['pmt_transit_time_mean'], cfg['pmt_transit_time_spread'] / 2.35482, n))
        dpe = rng.random(n) < cfg['p_double_pe_emision']
        chs = np.clip(ch, 0, self.n_ch - 1)
        gain = self.gains[chs] * self.spe[chs, _trunc(rng.random(n) * 2000) + 1]
        extra = self.gains[chs] * self.spe[chs, _trunc(rng.random(n) * 2000) + 1]
        gain = gain + np.where(dpe, extra, 0.0)
        return t, dpe, gain

    def truth_counters(self, t, ch, gain, dpe, keep_order=False):
        """pulse.py:229-271 incl. the `[:n_double_pe]` quirk (per channel, first photons in the
        channel slice; the slice order here is time order, or -- keep_order -- the order given, for
        inputs that already are channel slices as Pulse.__call__ cut them)."""
        cfg = self.cfg
        out = dict.fromkeys(['n_photon', 'n_pe', 'n_photon_trigger', 'n_pe_trigger', 'raw_area',
                             'raw_area_trigger'], 0)
        out.update({k + '_bottom': 0 for k in list(out)})
        live = (ch >= 0) & ~self.dead[np.clip(ch, 0, self.n_ch - 1)]
        if not live.any():
            return out
        t, ch, gain, dpe = t[live], ch[live], gain[live], dpe[live]
        order = np.argsort(ch, kind='stable') if keep_order else np.lexsort((t, ch))
        t, ch, gain, dpe = t[order], ch[order], gain[order], dpe[order]
        thr = np.full(self.n_ch, cfg['zle_threshold'] - 0.5)
        for k, v in (cfg.get('special_thresholds') or {}).items():
            if int(k) < self.n_ch:
                thr[int(k)] = v - 0.5
        above = gain * self.current_max[(t % self.dt).astype(int)] * self.c2a > thr[ch]
        area = gain / self.gains[ch]
        bounds = np.flatnonzero(np.diff(ch)) + 1
        starts = np.concatenate([[0], bounds])
        stops = np.concatenate([bounds, [len(ch)]])
        n_top = cfg['n_top_pmts']
        for a, b in zip(starts, stops):
            ndpe = int(dpe[a:b].sum())
            vals = dict(n_photon=b - a, n_photon_trigger=int(above[a:b].sum()),
                        n_pe=b - a + ndpe,
                        n_pe_trigger=int(above[a:b].sum()) + int(above[a:a + ndpe].sum()),
                        raw_area=area[a:b].sum(), raw_area_trigger=area[a:b][above[a:b]].sum())
            for k, v in vals.items():
                out[k] += v
                if ch[a] >= n_top:
                    out[k + '_bottom'] += v
        return out

    # ---- afterpulses --------------------------------------------------------------------------
    def pmt_afterpulses(self, t, ch, dpe):
        """afterpulse.py:172-249.  Returns (t, ch, gain) of the afterpulse photons."""
        cfg, rng = self.cfg, self.rng
        ts, chs, amps = [], [], []
        live = ch >= 0
        t, ch, dpe = t[live], ch[live], dpe[live]
        for name, el in self.pmt_ap.items():
            dcdf = np.asarray(el['delaytime_cdf'])
            acdf = np.asarray(el['amplitude_cdf'])
            rU0 = (1 - rng.random(len(t))) / cfg['pmt_ap_modifier']
            rU0[dpe] /= 2
            sel = np.flatnonzero(rU0 <= dcdf[ch, -1])
            if not len(sel):
                continue
            sch = ch[sel]
            rU1 = 1 - rng.random(len(sel))
            if 'Uniform' in name:
                delay = rng.uniform(dcdf[sch, 0], dcdf[sch, 1]) * el['delaytime_bin_size']
                amp = np.ones_like(delay)
            else:
                delay = np.argmin(np.abs(dcdf[sch] - rU0[sel][:, None]), axis=-1) * el['delaytime_bin_size'] \
                    - cfg['pmt_ap_t_modifier']
                rows = acdf[sch] if acdf.ndim == 2 else acdf[None, :]
                amp = np.argmin(np.abs(rows - rU1[:, None]), axis=-1) * el['amplitude_bin_size']
            ts.append(_trunc(t[sel] + delay))
            chs.append(sch)
            amps.append(np.atleast_1d(amp))
        if not ts:
            return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)
        ts, chs, amps = np.concatenate(ts), np.concatenate(chs), np.concatenate(amps)
        return ts, chs, self.gains[chs] * amps

    def pi_coarse_grid(self):
        cfg, h = self.cfg, self.ele_ap
        spread = np.sqrt(2 * cfg['diffusion_constant_longitudinal'] * h.bin_centers) / cfg['drift_velocity_liquid']
        coarse, cur = [], 100
        while cur < h.bin_centers[-1]:
            coarse.append(cur)
            cur += spread[np.argmin(np.abs(cur - h.bin_centers))]
        return np.array(coarse)

    def photoionization(self, photon_t, parent_row):
        """afterpulse.py:29-88: type-4 instruction rows spawned by one S2 Pulse call."""
        cfg, rng, h = self.cfg, self.rng, self.ele_ap
        if len(photon_t) == 0:
            return parent_row[:0].copy()
        n_e = rng.poisson(h.n * len(photon_t) * cfg['photoionization_modifier'])
        delay = h.get_random(n_e, rng)
        coarse = self.pi_coarse_grid()
        idx = np.digitize(delay[delay < coarse[-1]], coarse)
        idxs, cnt = np.unique(idx, return_counts=True)
        n = len(idxs)
        rows = np.repeat(parent_row[:1], n)
        rows['type'] = 4
        rows['time'] = photon_t[rng.integers(0, len(photon_t), n)] - cfg['drift_time_gate']
        r = np.sqrt(rng.uniform(0, cfg['tpc_radius'] ** 2, n))
        ang = rng.uniform(-np.pi, np.pi, n)
        rows['x'], rows['y'] = r * np.cos(ang), r * np.sin(ang)
        rows['z'] = -coarse[idxs] * cfg['drift_velocity_liquid']
        rows['amp'] = cnt
        return rows

    # ---- one Pulse call -----------------------------------------------------------------------
    def pulse_call(self, rows):
        """sim_data (rawdata.py:166-202) for one run: returns dict with pulses, last pulse end,
        truth-row pieces and spawned secondaries."""
        cfg = self.cfg
        typ = int(rows['type'][0])
        te = None
        if typ == 1:
            t, ch = self.s1_photons(rows)
        else:
            t, ch, te = self.s2_photons(rows)
        t, dpe, gain = self.pmt_stage(t, ch)
        self.n_photons_made += len(t)
        pulses = [det.pulse_call(cfg, self.templates, t[ch >= 0], ch[ch >= 0], gain[ch >= 0])]
        if cfg.get('enable_pmt_afterpulses', True) and self.pmt_ap and len(t):
            at, ach, ag = self.pmt_afterpulses(t, ch, dpe)
            if len(at):
                pulses.append(det.pulse_call(cfg, self.templates, at, ach, ag))
        P = det.Pulses.concat(pulses)
        spawned = rows[:0].copy()
        if typ == 2 and cfg.get('enable_electron_afterpulses', True) and self.ele_ap is not None:
            spawned = self.photoionization(t, rows)
        return dict(pulses=P, t=t, te=te, truth=self.truth_counters(t, ch, gain, dpe), spawned=spawned,
                    end=(int(P.right.max()) * self.dt if len(P) else None))

    # ---- scheduler (rawdata.py:38-157) + truth rows (rawdata.py:313-375) ----------------------
    def noise_offset(self, first_sample, span):
        """randint(0, high) of rawdata.py:407-417 (the stochastic oracle draws it from its generator)."""
        high = len(self.noise) - span - 1
        if high < 0:
            high = len(self.noise) - 1
        return int(self.rng.integers(0, high)) if high > 0 else 0

    def simulate(self, instructions, truth_dtype=None, ids=None):
        """rawdata.py:38-157.  Every instruction row carries an identity `_id` (primaries: `ids`, default
        their index; secondaries: what pulse_call gave them, or fresh numbers); `runs` lists the Pulse calls
        in execution order as (type, ids, digitisation group)."""
        cfg = self.cfg
        v, rext = cfg['drift_velocity_liquid'], cfg['right_raw_extension']
        save_full = cfg.get('save_full_truth', True)
        wdt = np.dtype([(n, instructions.dtype[n]) for n in instructions.dtype.names if n != '_id'] + [('_id', np.int64)])
        work = np.zeros(len(instructions), wdt)
        for n in instructions.dtype.names:
            work[n] = instructions[n]
        work['_id'] = np.arange(len(work)) if ids is None else ids
        instructions = work
        next_id = int(instructions['_id'].max()) + 1 if len(instructions) else 0

        def sig_time(rows):
            zf = rows['z'].astype(np.float32) / np.float32(v)
            k = (rows['type'].astype(np.int8) % 2 - 1).astype(np.float32)
            return rows['time'].astype(np.int64) + (zf * k).astype(np.int64)
        st = sig_time(instructions)
        order = np.argsort(st, kind='stable')
        cuts = np.flatnonzero(np.diff(st[order]) > rext) + 1
        queue = np.split(order, cuts)
        buf = instructions[:0].copy()
        last_end, cache = None, []
        groups, truth_rows, records, runs = [], [], [], []

        def flush():
            nonlocal cache
            if not cache:
                return
            P = det.Pulses.concat(cache)
            cache = []
            ix = 0
            if cfg.get('enable_noise', True) and self.noise is not None:
                span = int(P.right.max() - P.left.min()) + 2 * cfg['trigger_window']
                ix = self.noise_offset(int(P.left.min()) - cfg['trigger_window'], span)
            out = det.digitize_zle(cfg, P, noise=self.noise, ix_rand=ix)
            groups.append((out[5][0], out[5][1], len(out[0])))
            records.append(det.pack_records(cfg, *out[:5]))
        k = 0
        finished = False
        while not finished:
            if k < len(queue):
                buf = np.concatenate([buf, instructions[queue[k]]])
                k += 1
            bt = sig_time(buf)
            o = np.argsort(bt, kind='stable')
            buf, bt = buf[o], bt[o]
            if last_end is not None and len(buf) and bt[0] - last_end > rext:
                flush()
            clusters = np.split(np.arange(len(buf)), np.flatnonzero(np.diff(bt) > rext) + 1) if len(buf) else []
            keep = np.ones(len(buf), bool)
            spawned_all = []
            for cl in clusters:
                stop = False
                for ptype in (1, 2, 4, 6):
                    sel = cl[buf['type'][cl] == ptype]
                    if not len(sel):
                        continue
                    if ptype in (1, 2):
                        stop = True
                        if save_full:
                            sets = [sel[i:i + 1] for i in range(len(sel))]
                        else:
                            gap = 100 if ptype == 1 else int(0.2 / v)
                            sets = np.split(sel, np.flatnonzero(np.diff(bt[sel]) > gap) + 1)
                    else:
                        sets = [sel]
                    for s in sets:
                        rows = buf[s]
                        res = self.pulse_call(rows)
                        runs.append((ptype, rows['_id'].copy(), len(groups)))
                        if len(res['pulses']):
                            cache.append(res['pulses'])
                            last_end = res['end'] if last_end is None else max(last_end, res['end'])
                        if len(res['spawned']):
                            sp = res['spawned']
                            if not res.get('spawned_have_ids'):
                                sp = sp.copy()
                                sp['_id'] = np.arange(next_id, next_id + len(sp))
                                next_id += len(sp)
                            spawned_all.append(sp)
                        row = self.truth_row(rows, res, truth_dtype)
                        if row is not None:
                            truth_rows.append(row)
                        keep[s] = False
                if stop:
                    break
                flush()
            buf = buf[keep]
            if spawned_all:
                buf = np.concatenate([buf] + spawned_all)
            finished = k == len(queue) and len(buf) == 0
        flush()
        rec = np.concatenate(records) if records else np.zeros(0, det.raw_record_dtype())
        rec = det.sort_by_time(rec)
        truth = np.concatenate(truth_rows) if truth_rows else None
        return dict(records=rec, truth=truth, groups=groups, runs=runs)

    def truth_row(self, rows, res, truth_dtype):
        if truth_dtype is None:
            return None
        typ = int(rows['type'][0])
        t, te = res['t'], res['te']
        if len(t) == 0 and typ not in (1, 2):
            return None
        tb = np.zeros(1, truth_dtype)
        for q, times in (('photon', t), ('electron', te if te is not None else [])):
            if len(times):
                tb[f'n_{q}'] = len(times)
                tb[f't_mean_{q}'] = np.mean(times)
                tb[f't_first_{q}'] = np.min(times)
                tb[f't_last_{q}'] = np.max(times)
                tb[f't_sigma_{q}'] = np.std(times)
            else:
                tb[f'n_{q}'] = 0
                for f in ('t_mean_', 't_first_', 't_last_', 't_sigma_'):
                    tb[f + q] = np.nan
        xy = res.get('xy_obs')         # rawdata.py:377-390: only S2 calls under a field-distortion model
        tb['x_mean_electron'] = np.nan if xy is None else np.mean(xy[:, 0])
        tb['y_mean_electron'] = np.nan if xy is None else np.mean(xy[:, 1])
        if np.isnan(tb['t_last_photon'][0]):
            tb['endtime'] = rows['time'][0]
        else:
            tb['endtime'] = tb['t_last_photon'] + (cfg_n(self.cfg)) * self.dt
        for k, val in res['truth'].items():
            tb[k] = val
        for f in INSTR_FIELDS:
            if f not in rows.dtype.names:
                continue
            val = rows[f]
            if len(rows) > 1 and f in 'xyz':
                tb[f] = np.mean(val)
            elif len(rows) > 1 and f == 'amp':
                tb[f] = np.sum(val)
            else:
                tb[f] = val[0]
        return tb


def cfg_n(cfg):
    return cfg['samples_before_pulse_center'] + cfg['samples_after_pulse_center'] + 1


# ------------------------------------------------------------------------------------------
# chunker: strax_interface.py:368-497 applied to the group stream
# ------------------------------------------------------------------------------------------
def chunk_boundaries(cfg, t_min_instr, groups, time_zero=None):
    """Emulates the chunk_time bookkeeping of ChunkRawRecords.__call__ given, per digitisation
    group, (left, right, n_intervals).  Returns the list of (chunk_time_pre, chunk_time)."""
    dt = cfg['sample_duration']
    rext = int(cfg['right_raw_extension'])
    cksz = int(cfg['chunk_size'] * 1e9)
    pre = (time_zero - rext) if time_zero else (t_min_instr - rext)
    ct = pre + cksz
    cur_right = last_right = 0
    out = []
    for left, right, n_itv in groups:
        for _ in range(max(int(n_itv), 0)):
            if right != cur_right:
                last_right, cur_right = cur_right, right
            if left * dt > ct + rext:
                if (last_right + 1) * dt > ct:
                    ct += (last_right + 1) * dt - ct
                out.append((pre, ct))
                pre = ct
                ct += cksz
            else:
                break
    last_right = cur_right
    ct = max((last_right + 1) * dt, pre + dt)
    out.append((pre, ct))
    return out


class ReplayOracle(OracleSimulator):
    """Scheduler, Pulse calls, digitiser, ZLE, record packing and truth rows on PRESET stage outputs:
    the photons (time after transit-time spread, channel, gain, double-pe flag, PMT-afterpulse flag) and
    electron times of every instruction, and the secondary instructions every S2 spawned.  Nothing is
    sampled, so the result is a deterministic function of the presets -- comparable bit for bit with the
    reference run on the same presets (tests/golden/make_golden_sched.py) and with the CUDA path run on
    the photons it generated itself (tests/test_gpu_replay.py).

      photons      structured array: id, t, channel, gain, dpe, ap  (id = instruction identity)
      electrons    structured array: id, t
      secondaries  instruction rows + `_id` + `_parent` (identity of the S2 that spawned each)
      noise_seed   Philox seed of the noise start offsets (keyed by the first sample of the group)"""

    def __init__(self, cfg, photons, electrons=None, secondaries=None, noise=None, noise_seed=0, xy_obs=None):
        super().__init__(cfg, spe_table=np.zeros((len(cfg['gains']), 2001)), noise=noise)
        self.ph = photons[np.argsort(photons['id'], kind='stable')]
        self.el = None if electrons is None else electrons[np.argsort(electrons['id'], kind='stable')]
        self.sec = secondaries
        self.noise_seed = noise_seed
        self.xy_obs = xy_obs

    def noise_offset(self, first_sample, span):
        from .philox import noise_offset
        return noise_offset(self.noise_seed, first_sample, span, len(self.noise))

    @staticmethod
    def _take(arr, ids):
        parts = []
        for i in ids:
            a, b = np.searchsorted(arr['id'], i, 'left'), np.searchsorted(arr['id'], i, 'right')
            parts.append(arr[a:b])
        return np.concatenate(parts) if parts else arr[:0]

    def pulse_call(self, rows):
        cfg = self.cfg
        typ = int(rows['type'][0])
        ids = rows['_id']
        ph = self._take(self.ph, ids)
        main, ap = ph[~ph['ap'].astype(bool)], ph[ph['ap'].astype(bool)]
        te = None
        if typ != 1:
            te = self._take(self.el, ids)['t'] if self.el is not None else np.zeros(0, np.int64)
        pulses = []
        if not cfg.get('enable_pmt_afterpulses', True):
            ap = ap[:0]
        for part in (main, ap):           # the Pulse call itself, then its PMT afterpulses (rawdata.py:176-190)
            live = part['channel'] >= 0
            if live.any():
                pulses.append(det.pulse_call(cfg, self.templates, part['t'][live], part['channel'][live],
                                             part['gain'][live]))
        P = det.Pulses.concat(pulses)
        spawned = rows[:0].copy()
        if typ == 2 and self.sec is not None and len(self.sec) and len(main):      # rawdata.py:193-201
            sel = np.isin(self.sec['_parent'], ids) & (
                ((self.sec['type'] == 4) & bool(cfg.get('enable_electron_afterpulses', True))) |
                ((self.sec['type'] == 6) & bool(cfg.get('enable_gate_afterpulses', False))))
            if sel.any():
                spawned = np.zeros(int(sel.sum()), rows.dtype)
                for n in rows.dtype.names:
                    spawned[n] = self.sec[n][sel]
        xy = None
        if typ == 2 and self.xy_obs is not None:
            xy = np.array([self.xy_obs[int(i)] for i in ids])
        return dict(pulses=P, t=main['t'], te=te,
                    truth=self.truth_counters(main['t'], main['channel'], main['gain'], main['dpe'].astype(bool)),
                    spawned=spawned, spawned_have_ids=True, xy_obs=xy,
                    end=(int(P.right.max()) * self.dt if len(P) else None))
