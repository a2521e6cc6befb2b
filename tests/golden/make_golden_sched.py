"""Scheduler / truth golden vectors: the UNMODIFIED reference scheduler on PRESET stage outputs.

The reference's `RawData.__call__` (rawdata.py:38-157), `sim_data` (:166-202), `digitize_pulse_cache`,
`ZLE`, `get_truth` (:313-390), `Pulse.__call__` / `add_truth` / `add_current` (pulse.py:39-144, 229-318)
and `ChunkRawRecords` (strax_interface.py:353-504) run exactly as they are; only the SAMPLING is replaced:
the S1 / S2 / PhotoIonization_Electron / PhotoElectric_Electron / PMT_Afterpulse objects in
`RawData.pulses` are subclasses whose `__call__` loads preset photons (time, channel, gain) of the
instruction identities they are handed (carried in the `g4id` field) and then calls the reference's own
`Pulse.__call__`, and whose `generate_instruction` returns preset secondary instructions.  With the gains
preset the reference skips the transit-time spread and the SPE draw (pulse.py:53,95,105-107), so the run is
a deterministic function of the presets and can be compared bit for bit with
`oracle.wfsim_oracle_sim.ReplayOracle` and, for the grouping decisions, with the library's host scheduler
(`wfs_schedule`).

Cases cover: secondaries landing inside their parent's group, between groups, bridging two groups and
behind later primaries; pile-up of primaries under right_raw_extension; clusters of secondaries only;
save_full_truth on and off (S1s within 100 ns / S2s within 0.2 cm merged into one Pulse call); PMT
afterpulse calls extending the last pulse end; photo-electric (type 6) secondaries; dead PMTs.

A second, independent golden pins `Pulse.add_truth` with double photo-electron emission (`n_double_pe`
> 0, the `above_threshold[:n_double_pe]` rule of pulse.py:255): the method is called directly.

Run through tests/golden/make_golden.py:   python tests/golden/make_golden.py sched
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

PH_DT = np.dtype([('id', np.int64), ('t', np.int64), ('channel', np.int32), ('gain', np.float64),
                  ('dpe', np.int8), ('ap', np.int8)])
EL_DT = np.dtype([('id', np.int64), ('t', np.int64)])


def sig_time(rows, v):
    zf = rows['z'].astype(np.float32) / np.float32(v)
    k = (rows['type'].astype(np.int8) % 2 - 1).astype(np.float32)
    return rows['time'].astype(np.int64) + (zf * k).astype(np.int64)


def build_case(cfg, idt, seed, n_events, merge_clusters=False, gate=False):
    """Preset stage outputs of one synthetic run (all integers / exactly representable)."""
    rng = np.random.default_rng(seed)
    v = cfg['drift_velocity_liquid']
    gains = np.asarray(cfg['gains'])
    n_ch = len(gains)
    prim = []
    t = 5_000_000
    since_quiet = 0
    for ev in range(n_events):
        u = rng.random()
        if since_quiet >= 4 or u > 0.7:
            gap, since_quiet = int(rng.integers(2_000_000, 6_000_000)), 0
        elif u > 0.3:
            gap, since_quiet = int(rng.integers(150_000, 900_000)), since_quiet + 1
        else:
            gap, since_quiet = int(rng.integers(5_000, 80_000)), since_quiet + 1
        t += gap
        z = -float(rng.uniform(1, 96))
        n_scatter = int(rng.integers(2, 4)) if (merge_clusters and rng.random() < 0.5) else 1
        for k in range(n_scatter):
            for typ in (1, 2):
                r = np.zeros(1, idt)
                r['event_number'] = ev
                r['type'] = typ
                r['time'] = t + k * int(rng.integers(10, 60))
                rr, th = np.sqrt(rng.uniform(0, 45 ** 2)), rng.uniform(-np.pi, np.pi)
                r['x'], r['y'] = rr * np.cos(th), rr * np.sin(th)
                r['z'] = z - 0.03 * k * rng.random()           # < 0.2 cm apart: one S2 call when merging
                r['amp'] = int(rng.integers(5, 400))
                r['recoil'] = 7
                r['e_dep'] = float(rng.uniform(1, 100))
                r['local_field'] = 82.0
                r['vol_id'] = -1
                prim.append(r)
    prim = np.concatenate(prim)
    prim = prim[rng.permutation(len(prim))]             # the scheduler must not rely on the input order
    prim['g4id'] = np.arange(len(prim))
    n_prim = len(prim)
    st = sig_time(prim, v)
    photons, electrons, secs = [], [], []

    def some_photons(ident, t0, n, width, ap=False):
        ph = np.zeros(n, PH_DT)
        ph['id'] = ident
        ph['t'] = t0 + rng.integers(0, width, n)
        ph['channel'] = rng.integers(0, n_ch, n)
        if n > 8:                                         # pile several photons on a few PMTs
            ph['channel'][: n // 3] = rng.integers(0, 12, n // 3) * 37 % n_ch
        g = gains[ph['channel']] * rng.uniform(0.3, 2.2, n)
        g[rng.random(n) < 0.03] *= -0.2                   # the SPE table has negative entries
        ph['gain'] = g
        ph['ap'] = ap
        return ph

    next_id = n_prim
    for i in range(n_prim):
        typ = int(prim['type'][i])
        if typ == 1:
            n = 0 if rng.random() < 0.08 else int(rng.integers(3, 40))    # an S1 without detected photons
            ph = some_photons(i, int(st[i]) + 40, n, 250)
        else:
            n_e = int(rng.integers(0, 25)) if rng.random() < 0.9 else 0
            te = int(st[i]) + cfg['drift_time_gate'] + rng.integers(-400, 800, n_e)
            el = np.zeros(n_e, EL_DT)
            el['id'], el['t'] = i, te
            electrons.append(el)
            ph = np.concatenate([some_photons(i, int(x) + 30, int(rng.integers(4, 22)), 400) for x in te]) \
                if n_e else np.zeros(0, PH_DT)
        photons.append(ph)
        if len(ph) and rng.random() < 0.35:              # PMT afterpulses of this Pulse call
            photons.append(some_photons(i, int(ph['t'].max()), int(rng.integers(1, 6)), 9000, ap=True))
        if typ == 2 and len(ph) and rng.random() < 0.7:
            for _ in range(int(rng.integers(1, 7))):
                u = rng.random()
                delay = rng.uniform(100, 50_000) if u < 0.3 else rng.uniform(50_000, 700_000) if u < 0.8 \
                    else rng.uniform(700_000, 1_800_000)
                s = np.repeat(prim[i:i + 1], 1)
                s['type'] = 6 if (gate and rng.random() < 0.3) else 4
                t0 = int(ph['t'][rng.integers(0, len(ph))])
                s['time'] = t0 + cfg['drift_time_gate'] if s['type'][0] == 6 else t0 - cfg['drift_time_gate']
                rr, th = np.sqrt(rng.uniform(0, 50 ** 2)), rng.uniform(-np.pi, np.pi)
                s['x'], s['y'] = rr * np.cos(th), rr * np.sin(th)
                s['z'] = -delay * v
                s['amp'] = 1 if s['type'][0] == 6 else int(rng.integers(1, 5))
                s['g4id'] = next_id
                secs.append((s, i))
                sst = int(sig_time(s, v)[0])
                n_e = int(s['amp'][0]) if rng.random() < 0.9 else 0      # a secondary that yields nothing
                te = sst + cfg['drift_time_gate'] + rng.integers(-300, 600, n_e)
                el = np.zeros(n_e, EL_DT)
                el['id'], el['t'] = next_id, te
                electrons.append(el)
                if n_e:
                    sp = np.concatenate([some_photons(next_id, int(x) + 30, int(rng.integers(3, 18)), 400) for x in te])
                    photons.append(sp)
                    if rng.random() < 0.2:
                        photons.append(some_photons(next_id, int(sp['t'].max()), 2, 6000, ap=True))
                next_id += 1
    photons = np.concatenate(photons)
    electrons = np.concatenate(electrons) if electrons else np.zeros(0, EL_DT)
    sec_rows = np.concatenate([s for s, _ in secs]) if secs else prim[:0].copy()
    sec_parent = np.array([p for _, p in secs], np.int64)
    return prim, sec_rows, sec_parent, photons, electrons


def make_stub_rawdata(ref, cfg, prim, sec_rows, sec_parent, photons, electrons):
    """RawData with preset-loading pulse objects; every scheduling / digitising method is the reference's."""
    ph_by = {}
    for k in np.unique(photons['id']):
        ph_by[int(k)] = photons[photons['id'] == k]
    el_by = {int(k): electrons['t'][electrons['id'] == k] for k in np.unique(electrons['id'])}

    def load(pulse, ids, ap):
        parts = [ph_by[int(i)] for i in ids if int(i) in ph_by]
        ph = np.concatenate(parts) if parts else np.zeros(0, PH_DT)
        ph = ph[ph['ap'].astype(bool) == ap]
        order = np.argsort(ph['channel'], kind='stable')       # s1.py:110-113 / s2.py:131-134 / afterpulse.py:245
        pulse._photon_timings = ph['t'][order].copy()
        pulse._photon_channels = ph['channel'][order].astype(np.int64)
        pulse._photon_gains = ph['gain'][order].copy()

    def rows_of(instruction):
        return np.array([instruction]) if len(instruction.shape) < 1 else instruction

    class PS1(ref.S1):
        def __call__(self, instruction):
            self._last_ids = rows_of(instruction)['g4id']
            load(self, self._last_ids, False)
            ref.Pulse.__call__(self)

    class _S2Like:
        def __call__(self, instruction):
            self._last_ids = rows_of(instruction)['g4id']
            load(self, self._last_ids, False)
            te = [el_by[int(i)] for i in self._last_ids if int(i) in el_by]
            self._electron_timings = np.concatenate(te) if te else np.zeros(0, np.int64)
            ref.Pulse.__call__(self)

    class PS2(_S2Like, ref.S2):
        pass

    class _Spawner(_S2Like):
        sec_type = 4

        def generate_instruction(self, signal_pulse, signal_pulse_instruction):
            if len(signal_pulse._photon_timings) == 0:          # afterpulse.py:24-27, 102-104
                return []
            ids = rows_of(signal_pulse_instruction)['g4id']
            sel = np.isin(sec_parent, ids) & (sec_rows['type'] == self.sec_type)
            return sec_rows[sel].copy()

    class PPi(_Spawner, ref.PhotoIonization_Electron):
        sec_type = 4

    class PPe(_Spawner, ref.afterpulse.PhotoElectric_Electron):
        sec_type = 6

    class PAp(ref.PMT_Afterpulse):
        def __call__(self, signal_pulse):
            if len(signal_pulse._photon_timings) == 0:          # afterpulse.py:161-164
                self.clear_pulse_cache()
                return
            load(self, signal_pulse._last_ids, True)
            ref.Pulse.__call__(self)

    class StubRawData(ref.RawData):
        def __init__(self, config, **kw):
            self.config = config
            self.pulses = dict(s1=PS1(config), s2=PS2(config), pi_el=PPi(config), pe_el=PPe(config),
                               pmt_ap=PAp(config))
            self.resource = ref.load_resource.load_config(config)
            self.log_runs, self.log_groups = [], []

        def sim_data(self, instruction, **kw):
            self.log_runs.append((int(instruction['type'][0]), instruction['g4id'].copy(), len(self.log_groups)))
            yield from super().sim_data(instruction, **kw)

        def digitize_pulse_cache(self):
            self._had = len(self._pulses_cache) > 0
            super().digitize_pulse_cache()
            if self._had:
                self.log_groups.append([int(self.left), int(self.right), 0])

        def ZLE(self):
            for x in super().ZLE():
                self.log_groups[-1][2] += 1
                yield x
    return StubRawData


def run_case(ref, cfg, name, seed, n_events, **kw):
    idt = np.dtype(ref.strax_interface.instruction_dtype)
    prim, sec_rows, sec_parent, photons, electrons = build_case(cfg, idt, seed, n_events, **kw)
    Stub = make_stub_rawdata(ref, cfg, prim, sec_rows, sec_parent, photons, electrons)
    # pass 1: RawData on its own -> truth rows in execution order, runs, groups
    tdt = np.dtype(ref.strax_interface.instruction_dtype + ref.strax_interface.truth_extra_dtype + [('fill', bool)])
    tb = np.zeros(4000, tdt)
    rd = Stub(dict(cfg))
    n_itv = sum(1 for _ in rd(prim.copy(), tb, progress_bar=False))
    truth = tb[tb['fill']]
    tdt2 = np.dtype(ref.strax_interface.instruction_dtype + ref.strax_interface.truth_extra_dtype)
    truth_out = np.zeros(len(truth), tdt2)
    for n in tdt2.names:
        truth_out[n] = truth[n]
    runs, groups = rd.log_runs, rd.log_groups
    # pass 2: through the reference chunker / record packer
    crr = ref.ChunkRawRecords(dict(cfg), rawdata_generator=Stub)
    crr.record_buffer = np.zeros(400000, dtype=crr.record_buffer.dtype)
    chunks, rr, rr_he, ctruth = [], [], [], []
    for res in crr(prim.copy(), progress_bar=False):
        chunks.append((int(crr.chunk_time_pre), int(crr.chunk_time), len(res['raw_records']), len(res['truth'])))
        rr.append(res['raw_records'].copy())
        rr_he.append(res['raw_records_he'].copy())
        ctruth.append(res['truth'].copy())
        assert len(res['raw_records_aqmon']) == 0
    rr, rr_he, ctruth = np.concatenate(rr), np.concatenate(rr_he), np.concatenate(ctruth)
    assert [g[:2] for g in crr.rawdata.log_groups] == [g[:2] for g in groups]
    print(f'{name}: {len(prim)} primaries, {len(sec_rows)} secondaries, {len(photons)} photons -> {len(runs)} Pulse '
          f'calls, {len(groups)} groups, {n_itv} intervals, {len(rr)} records, {len(truth)} truth rows, '
          f'{len(chunks)} chunks')
    run_len = np.array([len(r[1]) for r in runs], np.int32)
    return {
        f'{name}_prim': prim.view(np.uint8), f'{name}_sec': sec_rows.view(np.uint8),
        f'{name}_sec_parent': sec_parent, f'{name}_photons': photons.view(np.uint8),
        f'{name}_electrons': electrons.view(np.uint8),
        f'{name}_run_type': np.array([r[0] for r in runs], np.int8), f'{name}_run_len': run_len,
        f'{name}_run_ids': np.concatenate([r[1] for r in runs]).astype(np.int64),
        f'{name}_run_group': np.array([r[2] for r in runs], np.int32),
        f'{name}_groups': np.array(groups, np.int64).reshape(-1, 3),
        f'{name}_truth': truth_out.view(np.uint8), f'{name}_chunk_truth': ctruth.view(np.uint8),
        f'{name}_rr': rr.view(np.uint8), f'{name}_rr_he': rr_he.view(np.uint8),
        f'{name}_chunks': np.array(chunks, np.int64).reshape(-1, 4),
    }


CASES = {
    # name: (cfg overrides, seed, n_events, build kwargs)
    'full': (dict(), 101, 26, dict()),
    'merged': (dict(save_full_truth=False), 102, 22, dict(merge_clusters=True)),
    'gate': (dict(enable_gate_afterpulses=True), 103, 18, dict(gate=True)),
    'noap': (dict(enable_pmt_afterpulses=False, enable_electron_afterpulses=False, zle_threshold=40), 104, 14, dict()),
}
BASE = dict(enable_pmt_afterpulses=True, enable_electron_afterpulses=True,
            photon_ap_cdfs='synthetic_ap.json.gz', ele_ap_pdfs='synthetic_ele_ap.dill', chunk_size=0.02)


def add_truth_golden(ref, cfg):
    """Pulse.add_truth (pulse.py:229-271) called directly with double photo-electrons."""
    rng = np.random.default_rng(7)
    cfg = dict(cfg, special_thresholds={'7': 40, '255': 5})
    p = ref.Pulse(dict(cfg))
    p._truth_buffer = {}
    for field in 'n_photon n_pe n_photon_trigger n_pe_trigger raw_area raw_area_trigger'.split():
        p._truth_buffer[field] = 0
        p._truth_buffer[field + '_bottom'] = 0
    gains = np.asarray(cfg['gains'])
    rows = []
    for ch in (0, 7, 100, 255, 300, 420, 493, 3):
        n = int(rng.integers(1, 30))
        t = rng.integers(1000, 3000, n).astype(np.int64)
        g = gains[ch] * rng.uniform(0.05, 2.5, n)
        ndpe = int(rng.integers(0, n + 1))
        p.add_truth(t, g, cfg['sample_duration'], ch, ndpe)
        ph = np.zeros(n, PH_DT)
        ph['t'], ph['channel'], ph['gain'] = t, ch, g
        ph['dpe'][:ndpe] = 1                                 # the slice order IS the rule: first n_dpe count
        rows.append(ph)
    out = {'addtruth_photons': np.concatenate(rows).view(np.uint8),
           'addtruth_cfg': np.array(json.dumps({'special_thresholds': {'7': 40, '255': 5}}))}
    for k, val in p._truth_buffer.items():
        out['addtruth_' + k] = np.array(val, np.float64)
    return out


def main(ref, c0_config):
    out = {}
    for name, (extra, seed, n_events, kw) in CASES.items():
        over = dict(BASE)
        over.update(extra)
        cfg, _, _ = c0_config(**over)
        gains = cfg['gains'].copy()
        gains[[3, 100, 300]] = 0                             # dead PMTs (pulse.py:89-90)
        cfg['gains'] = gains
        out.update(run_case(ref, cfg, name, seed, n_events, **kw))
        out[f'{name}_cfg'] = np.array(json.dumps(over))
        if name == 'full':
            out.update(add_truth_golden(ref, cfg))
    np.savez_compressed(os.path.join(HERE, 'sched.npz'), **out)
