#!/usr/bin/env python
"""Benchmark of the WFSim hot path (wfsim_instructions -> raw_records) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path, BASELINE config C1
    python bench.py --config C2|C3|C4 [--events n] ...       # the other BASELINE.json configs
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Workloads (synthetic instructions from tests/golden/synth_instructions.py):
  C1 (default, the config the metric is quoted on)  1e5 low-energy (1-50 keV) events at 1 kHz, XENONnT 494
      channels, fax_config = the reference's shipped test config (dummy maps).  Weak scaling: every rank simulates
      its own 1e5 events (different instruction seed); no data-path collective (events are independent).
  C2  S2-heavy events (1e4-1e5 extracted electrons) with PMT afterpulses + photo-ionisation (synthetic AP tables).
  C3  mixed stream (95 % C1-like, 5 % C2-like), noise + ZLE, cut into 8 chunks in time: chunk k runs on rank
      k mod N (strong scaling; no collective), chunk after chunk into one recycled page-locked record arena.
  C4  pulse-superposition microbench: pulses of 1e6 photons over 494 channels through wfs_simulate_photons,
      template sweep (11 / 22 / 32 samples per template; longer ones are refused by the library); metric photons/s.
One step = one pass of the whole path over the workload.  The default run also carries bounded samples of
C2, C3 and C4 in `configs` (rank 0, N = 1 only).

Prints ONE JSON line (rank 0).  value = the metric with the instructions already planned and resident
(wfs_stage_instructions / wfs_run_staged; records stay in HBM); e2e = the same metric through the public API with
host buffers (instructions H2D, records + truth D2H into the caller's page-locked numpy array) inside the timed
region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RECORD_BYTES, TRUTH_BYTES, INSTR_BYTES, PHOTON_BYTES = 244, 218, 70, 24

WORKLOAD_NAME = ('C1: low-energy S1+S2 recoils 1-50 keV, 1 kHz, XENONnT 494 ch, '
                 'XENONnT_wfsim_config.json + test_load_nt dummy maps')
WORKLOADS = {
    'C1': WORKLOAD_NAME,
    'C2': ('C2: S2-heavy events (1e4-1e5 extracted electrons, S1 = 2 x S2 quanta) with PMT afterpulses + '
           'photo-ionisation (synthetic AP tables of SURVEY 8d), same detector config'),
    'C3': ('C3: mixed stream, 95 % C1-like + 5 % C2-like events, enable_noise (synthetic noise_data round(N(0,2)) '
           '[2^16, 494]) + ZLE, 8 chunks in time, chunk k on rank k mod N'),
    'C4': ('C4: pulse superposition, pulses of 1e6 photons (t ~ N(t0_k, 1 us), 62.7 % top / 37.3 % bottom, '
           'gain ~ gains[ch] (0.3 + Exp(0.7))) through wfs_simulate_photons, ZLE threshold 0'),
}


def load_config(**extra):
    from tests.conftest import load_c0_config
    return load_c0_config(**extra)


def spe_tables():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c0_tables.npz'))
    return z['spe_unique'], z['spe_row'][:494]


def workload(n_events, seed):
    from tests.golden.synth_instructions import c1_like
    return c1_like(n_events, seed=seed)


def heavy_events(n_events, seed):
    from tests.test_gpu_configs import heavy_s2_events
    return heavy_s2_events(n_events, seed=seed, n_e=(10_000, 100_000))


def mixed_stream(n_events, seed):
    """95 % low-energy + 5 % heavy events, interleaved in time: every heavy event takes the place half a
    millisecond behind a randomly chosen light one."""
    lo = workload(max(int(n_events * 0.95), 1), seed)
    n_hi = max(n_events - int(n_events * 0.95), 1)
    hi = heavy_events(n_hi, seed + 1)
    rng = np.random.default_rng(seed + 2)
    lo_times = np.unique(lo['time'])
    slot = rng.choice(len(lo_times), size=n_hi, replace=n_hi > len(lo_times))
    ev, inv = np.unique(hi['event_number'], return_inverse=True)
    t_first = np.full(len(ev), np.iinfo(np.int64).max)
    np.minimum.at(t_first, inv, hi['time'])
    hi['time'] = hi['time'] - t_first[inv] + lo_times[slot[inv % n_hi]] + 500_000
    inst = np.concatenate([lo, hi])
    return inst[np.argsort(inst['time'], kind='stable')]


def ncu_traffic():
    """DRAM bytes per device batch of the fused back end's kernels from the committed ncu capture."""
    path = os.path.join(ROOT, 'profiles', 'r2_fused_ncu.json')
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


def host_array(n, dtype):
    """Host array for the records as a plugin holds it (the chunker's arena), touched once so that page
    faults are not timed."""
    from wfsim_b200.simulator import host_records
    a = host_records(n, dtype)
    a.view(np.uint8)[::4096] = 0
    return a


def algorithmic_bytes(c):
    """every compulsory stream counted once."""
    return (PHOTON_BYTES * c['n_photons'] + RECORD_BYTES * c['n_records_total']
            + TRUTH_BYTES * c['n_truth'] + INSTR_BYTES * c['n_instructions'])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                      '-i', str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith('active'):
                        self.reasons.add(nm)
            except Exception:
                pass
            # (an nvidia-smi query takes the driver's locks: with one sampler per rank, 8 ranks, keep them rare)
            self._halt.wait(0.2 if int(os.environ.get('WORLD_SIZE', 1)) == 1 else 1.0)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'samples': len(self.samples)}


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)'
    return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


# ------------------------------------------------------------------------------------------------
# CPU baselines on the box's host cores.  kind "reference": the UNMODIFIED reference modules (staged by
# oracle/make_ref.sh into the git-ignored oracle/_ref, or /root/reference in the build container) driven through
# their own ChunkRawRecords; kind "port": the C / numpy restatement under oracle/.  Both are single-threaded like
# the reference (strax_interface.py:546), so P worker processes run disjoint event slices.
# ------------------------------------------------------------------------------------------------
def _port_worker(args):
    n_events, seed = args
    from oracle import wfsim_oracle_sim as osim
    from wfsim_b200.dtypes import truth_dtype
    cfg = load_config()
    uniq, row = spe_tables()
    inst = workload(n_events, seed)
    sim = osim.OracleSimulator(cfg, uniq[row], seed=seed)
    t0 = time.perf_counter()
    out = sim.simulate(inst, truth_dtype=truth_dtype())
    dt = time.perf_counter() - t0
    tr = out['truth']
    return dict(seconds=dt, n_pe=int(tr['n_pe'].sum()), n_records=len(out['records']), n_events=n_events)


def _reference_worker(args):
    n_events, seed = args
    from oracle import reference_runner as rr
    r = rr.run_events(load_config(), workload(n_events, seed), seed=seed)
    r['n_events'] = n_events
    return r


def reference_available():
    from oracle import ref_loader as RL
    try:
        import numba  # noqa: F401
    except ImportError:
        return False
    return RL.available()


def _run_workers(fn, events_per_worker, workers):
    import multiprocessing as mp
    t0 = time.perf_counter()
    if workers > 1:
        with mp.get_context('fork').Pool(workers) as pool:
            res = pool.map(fn, [(events_per_worker, 1000 + w) for w in range(workers)])
    else:
        res = [fn((events_per_worker, 1000))]
    wall = time.perf_counter() - t0
    return res, wall


def cpu_baseline_port(events_per_worker, workers):
    from oracle import wfsim_oracle as orc
    orc.build()
    _port_worker((2, 999))          # warm caches / library load outside the timed runs
    res, wall = _run_workers(_port_worker, events_per_worker, workers)
    n_pe = sum(r['n_pe'] for r in res)
    n_rec = sum(r['n_records'] for r in res)
    return dict(value=n_pe / wall, unit='pe/s', cores=workers, kind='port',
                sample=f'{events_per_worker} events x {workers} worker processes of the same '
                       f'C1 workload ({wall:.1f} s wall)',
                raw_records_gbs=n_rec * RECORD_BYTES / wall / 1e9, seconds=wall)


_ref_warm = [False]


def cpu_baseline_reference(events_per_worker, workers):
    """The reference itself: numba-compiled in this process by a throw-away run of two events, then forked into
    `workers` processes (the compiled code is inherited), each timing its own slice."""
    import logging
    logging.disable(logging.WARNING)      # the reference logs every aux file the stand-in loader cannot serve
    try:
        if not _ref_warm[0]:
            _reference_worker((2, 999))
            _ref_warm[0] = True
        res, wall = _run_workers(_reference_worker, events_per_worker, workers)
    finally:
        logging.disable(logging.NOTSET)
    n_pe = sum(r['n_pe'] for r in res)
    n_rec = sum(r['n_records'] for r in res)
    per_core = float(np.mean([r['n_pe'] / r['seconds'] for r in res]))
    return dict(value=n_pe / wall, unit='pe/s', cores=workers, kind='reference',
                sample=f'{events_per_worker} events x {workers} worker processes of the same C1 workload through the '
                       f'unmodified reference modules (RawData + ChunkRawRecords, numba JIT-warmed; {wall:.1f} s wall)',
                raw_records_gbs=n_rec * RECORD_BYTES / wall / 1e9, seconds=wall, pe_per_s_per_core=per_core)


def run_reference(args, rank, world):
    if rank != 0:
        return
    workers = len(os.sched_getaffinity(0))
    use_ref = reference_available() and not os.environ.get('WFS_BENCH_REF_PORT')
    per = max(2, int(args.ref_events if use_ref else args.port_events))
    fn = cpu_baseline_reference if use_ref else cpu_baseline_port
    vals, recs, secs = [], [], []
    b = None
    for i in range(args.warmup + args.steps):
        if 0 < i < args.warmup:
            continue          # one warm-up pass is enough for a CPU code (JIT, page cache): more would only burn minutes
        b = fn(per, workers)
        if i >= args.warmup:
            vals.append(b['value']); recs.append(b['raw_records_gbs']); secs.append(b['seconds'])
    v = float(np.mean(vals))
    base = {k: b[k] for k in ('unit', 'cores', 'kind', 'sample', 'pe_per_s_per_core') if k in b}
    base['value'] = v
    if not use_ref:
        base['sample'] += '; oracle/ port (no reference modules staged: run oracle/make_ref.sh in the build container)'
    line = {
        'impl': 'reference', 'metric': 'photoelectrons_per_s', 'value': v, 'unit': 'pe/s',
        'raw_records_gbs': float(np.mean(recs)),
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': float(np.mean(secs)) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_NAME, 'sample_events_per_step': per * workers},
        'cpu_baseline': base,
        'e2e': {'value': v, 'unit': 'pe/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing: barrier + max / sum over ranks (NCCL); no data-path collective."""

    def __init__(self, world, local_rank):
        import torch
        self.torch = torch
        self.world = world
        torch.cuda.set_device(local_rank)
        if world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, values, op='sum'):
        if self.world == 1:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], device='cuda', dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == 'max' else self.dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    def gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def one_lane(fn):
    """Runs fn with WFS_LANES=1, so that CUDA events around a kernel bracket that kernel only."""
    before = os.environ.get('WFS_LANES')
    os.environ['WFS_LANES'] = '1'
    try:
        return fn()
    finally:
        if before is None:
            del os.environ['WFS_LANES']
        else:
            os.environ['WFS_LANES'] = before


def measure_path(sim, inst, args, D, clocks_gpu=None, rng_id=None, e2e=True):
    """Device-resident leg (stage + run_staged) and end-to-end leg (simulate into the caller's page-locked array)
    of one instruction set on this rank.  Times are this rank's; the caller reduces over ranks."""
    from wfsim_b200.dtypes import raw_record_dtype
    sim.stage(inst)
    for _ in range(args.warmup):
        c = sim.run_staged(seed=1)
    D.barrier()
    sampler = ClockSampler(clocks_gpu) if clocks_gpu is not None and not os.environ.get('WFS_BENCH_NO_CLOCKS') else None
    if sampler:
        sampler.start()
    ms, launches = [], 0
    phases = np.zeros(12)
    for k in range(args.steps):
        c = sim.run_staged(seed=1)
        ms.append(c['ms_total']); launches += c['gpu_launches']
        phases += np.array(c['ms_phase'])
    D.barrier()
    clocks = sampler.stop() if sampler else None

    def alone():
        sim.run_staged(seed=1)
        return [sim.run_staged(seed=1)['ms_phase'] for _ in range(args.steps)]
    ph_alone = np.mean(np.array(one_lane(alone)), axis=0)
    D.barrier()
    r = dict(counts=c, t_dev=float(np.sum(ms)) / 1e3, launches=launches, phases=phases / args.steps,
             phases_alone=ph_alone, clocks=clocks)
    if not e2e or getattr(args, 'no_e2e', False):
        return r
    cap = int(c['n_records_total'] * 1.02) + 1024
    records_out = host_array(cap, raw_record_dtype())
    pinned = (not os.environ.get('WFS_BENCH_PAGEABLE')) and sim.pin(records_out)
    e2e_ms, e2e_lib = [], []
    h2d = inst.nbytes + 494 * 4 * 2 + len(inst) * (8 * 3 + 4)
    d2h = 0
    for k in range(args.warmup + args.steps):
        D.barrier()
        t0 = time.perf_counter()
        sim.simulate(inst, seed=1, cap_records=cap, records_out=records_out, rng_id=rng_id)
        D.torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        lc = sim.last_counts
        if k >= args.warmup:
            e2e_ms.append(dt * 1e3)
            e2e_lib.append([lc['ms_total']] + list(lc['ms_phase'][8:10]) + [lc['n_plain_records']])
        # record bytes that crossed PCIe (compact headers + differing sample blocks and / or plain rows)
        # + truth rows + group info
        d2h = lc['d2h_bytes'] + lc['n_truth'] * TRUTH_BYTES + lc['n_groups'] * 24
    r.update(t_e2e=float(np.sum(e2e_ms)) / 1e3, h2d=int(h2d), d2h=int(d2h), pinned=bool(pinned),
             ms_device=float(np.mean([x[0] for x in e2e_lib])), ms_d2h=float(np.mean([x[1] for x in e2e_lib])),
             ms_expand=float(np.mean([x[2] for x in e2e_lib])),
             plain_share=[round(x[3] / max(c['n_records_total'], 1), 3) for x in e2e_lib])
    del records_out
    return r


def roofline_of(c, ph_alone, ph_lanes, peak, peak_src):
    """The back end's dominant kernel(s) against the HBM roofline, from the one-lane pass (kernel alone on the GPU)."""
    nb = max(int(c['n_batches']), 1)
    if c['n_fused_batches'] == c['n_batches']:
        # group-resident fused back end: k_group_analyse (all size classes, side by side) + k_group_records
        # take the photons (24 B each) and leave the records (244 B each)
        t_an, t_rec = float(ph_alone[3]), float(ph_alone[6])
        byt = PHOTON_BYTES * c['n_photons'] + RECORD_BYTES * c['n_records_total']
        traffic = ncu_traffic()
        ms = t_an + t_rec
        return {'bound': 'hbm', 'kernel': 'k_group_analyse + k_group_records (fused back end: photons in, records out)',
                'achieved': byt / (ms / 1e3) / 1e9 if ms > 0 else None, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': byt / (ms / 1e3) / 1e9 / peak if ms > 0 else None,
                'traffic': (traffic or {}).get('dram_bytes_per_batch'),
                'traffic_source': (traffic or {}).get('source'),
                'launches_per_step': nb, 'algorithmic_bytes_per_launch': int(byt / nb),
                'algorithmic_bytes_per_step': int(byt), 'kernel_ms_per_step': ms,
                'kernels': {'k_group_analyse': {'ms_per_step': t_an, 'ms_per_step_in_timed_region': float(ph_lanes[3])},
                            'k_group_records': {'ms_per_step': t_rec, 'ms_per_step_in_timed_region': float(ph_lanes[6]),
                                                'achieved_gbs': RECORD_BYTES * c['n_records_total'] / (t_rec / 1e3) / 1e9 if t_rec > 0 else None}},
                'timing': 'CUDA events on the library stream around the launches of every device batch, summed per step, '
                          'in a pass of the same steps with one lane (kernels alone on the GPU); a launch = the back end of one '
                          'device batch (the size classes of k_group_analyse / k_group_analyse_small side by side, scan, k_group_records)'}
    ms = float(ph_alone[3])
    byt = PHOTON_BYTES * c['n_photons'] + 2 * c['n_samples']
    return {'bound': 'hbm', 'kernel': 'k_digitize (multi-pass back end: photons in, dense int16 samples out)',
            'achieved': byt / (ms / 1e3) / 1e9 if ms > 0 else None, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
            'frac': byt / (ms / 1e3) / 1e9 / peak if ms > 0 else None, 'traffic': None,
            'launches_per_step': nb, 'algorithmic_bytes_per_launch': int(byt / nb), 'algorithmic_bytes_per_step': int(byt),
            'kernel_ms_per_step': ms, 'kernel_ms_per_step_in_timed_region': float(ph_lanes[3]),
            'fused_batches': int(c['n_fused_batches']), 'batches': int(c['n_batches']),
            'timing': 'CUDA events on the library stream around every k_digitize launch, one-lane pass'}


PHASE_NAMES = ['frontend', 'photon_sort', 'windows', 'digitize_or_group_analyse', 'zle', 'record_sort',
               'record_pack_or_group_records', 'host_scheduler_truth']


def path_line(name, r, D, args, weak=True, extra_config=None):
    """JSON fields of an instruction-driven config from this rank's measurement `r` (reduced over ranks)."""
    c = r['counts']
    t_dev, = D.reduce([r['t_dev']], 'max')
    n_pe_all, n_rec_all, launches_all, n_ph_all = D.reduce([c['n_pe'], c['n_records_total'], r['launches'], c['n_photons']])
    peak, peak_src = measured_peak()
    balg = algorithmic_bytes(c)
    line = {
        'metric': 'photoelectrons_per_s', 'value': n_pe_all * args.steps / t_dev, 'unit': 'pe/s',
        'raw_records_gbs': n_rec_all * RECORD_BYTES * args.steps / t_dev / 1e9,
        'n_gpus': D.world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak' if weak else 'strong',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': dict({'workload': WORKLOADS[name], 'instructions_per_gpu': int(c['n_instructions']),
                        'photons_per_step_per_gpu': int(c['n_photons']),
                        'records_per_step_per_gpu': int(c['n_records_total']),
                        'device_batches_per_step': int(c['n_batches']), 'fused_batches_per_step': int(c['n_fused_batches']),
                        'l2': 'inputs and outputs of every batch exceed the 126 MB L2 (no flush needed)'},
                       **(extra_config or {})),
    }
    if D.world > 1:
        line['ms_per_step_per_rank'] = [round(x['t_dev'] / args.steps * 1e3, 2) for x in D.gather_objects({'t_dev': r['t_dev']})]
    if r.get('clocks') is not None:
        line['clocks'] = r['clocks']
    if 't_e2e' in r:
        t_e2e, = D.reduce([r['t_e2e']], 'max')
        line['e2e'] = {'value': n_pe_all * args.steps / t_e2e, 'unit': 'pe/s', 'h2d_bytes_per_step': r['h2d'],
                       'd2h_bytes_per_step': r['d2h'], 'ms_per_step': t_e2e / args.steps * 1e3,
                       'raw_records_gbs': n_rec_all * RECORD_BYTES * args.steps / t_e2e / 1e9,
                       # inside the call (rank 0, per step): device work of all batches; wall clock summed over batches
                       # of (batch shipped -> its D2H done) and (D2H done -> expanded by host threads)
                       'ms_device': r['ms_device'], 'ms_batches_d2h': r['ms_d2h'], 'ms_batches_expand': r['ms_expand'],
                       'destination': 'caller-owned numpy array, ' + ('page-locked (wfs_host_register) as the plugin\'s '
                                                                      'record arenas are' if r['pinned'] else 'pageable'),
                       'plain_record_share': r['plain_share']}
    line['gpu_launches'] = int(launches_all)
    line['ms_phase_per_step'] = dict(zip(PHASE_NAMES, np.round(r['phases'][:8], 3).tolist()))
    line['path_hbm'] = {'algorithmic_bytes_per_step': int(balg), 'achieved_gbs': balg / (t_dev / args.steps) / 1e9,
                        'frac_of_measured_peak': balg / (t_dev / args.steps) / 1e9 / peak}
    line['roofline'] = roofline_of(c, r['phases_alone'], r['phases'], peak, peak_src)
    return line


def make_sim(name, local_rank):
    from wfsim_b200.resource import Resource
    from wfsim_b200.simulator import Simulator
    uniq, row = spe_tables()
    if name == 'C2':
        from tests.golden.synth_tables import EleApHist, pmt_ap_tables
        cfg = load_config(enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
        res = Resource(cfg, spe_ppf=uniq, spe_row=row, uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist())
    elif name == 'C3':
        cfg = load_config(enable_noise=True)
        noise = np.round(np.random.default_rng(3).normal(0, 2.0, (1 << 16, 494)))
        res = Resource(cfg, spe_ppf=uniq, spe_row=row, noise_data=noise)
    else:
        cfg = load_config()
        res = Resource(cfg, spe_ppf=uniq, spe_row=row)
    return Simulator(cfg, resource=res, device=local_rank), cfg


def run_c2(args, D, rank, local_rank, n_events, clocks=True):
    sim, cfg = make_sim('C2', local_rank)
    inst = heavy_events(n_events, seed=200 + rank)
    r = measure_path(sim, inst, args, D, clocks_gpu=local_rank if clocks else None)
    line = path_line('C2', r, D, args, extra_config={
        'events_per_gpu': n_events,
        'stand_ins': 'uniform_to_pmt_ap / uniform_to_ele_ap are the synthetic tables of SURVEY 8d (the real files are not in the tree)'})
    sim.close()
    return line


def run_c3(args, D, rank, local_rank, n_events, clocks=True, n_chunks=8):
    """One stream, the same on every rank, cut into `n_chunks` time ranges at quiet gaps; rank r simulates chunks
    r, r + N, ... one after the other into one recycled page-locked arena.  Strong scaling.  Check: the records of a
    closed prefix of every chunk, hashed by the rank that made them, against a separate single-GPU call on rank 0."""
    import hashlib
    from wfsim_b200.dtypes import raw_record_dtype
    from wfsim_b200.sharding import shard_instructions, signal_time
    sim, cfg = make_sim('C3', local_rank)
    inst = mixed_stream(n_events, seed=300)
    gap = sim.quiet_gap()
    parts = shard_instructions(inst, n_chunks, cfg, min_gap=gap)
    mine = [k for k in range(n_chunks) if k % D.world == rank and len(parts[k])]
    st = signal_time(inst, cfg['drift_velocity_liquid'])

    def prefix_of(idx, m=400):          # a prefix of the chunk that ends at a quiet gap
        s = st[idx]
        ok = np.flatnonzero(np.diff(s) > gap) + 1
        ok = ok[ok >= min(m, len(idx) - 1)]
        return (idx[:ok[0]], int(s[ok[0]])) if len(ok) else (idx, None)

    # device-resident leg, chunk after chunk
    t_dev = 0.0
    keys = ('n_pe', 'n_records_total', 'n_photons', 'n_truth', 'n_instructions', 'n_batches', 'n_fused_batches', 'n_samples')
    tot = dict.fromkeys(keys, 0)
    launches, phases, ph_alone = 0, np.zeros(12), np.zeros(12)
    sampler = ClockSampler(local_rank) if clocks else None
    if sampler:
        sampler.start()
    caps = {}
    c = None
    for k in mine:
        sim.stage(inst[parts[k]])          # (Philox identities inside a staged chunk are local: the timing leg only)
        for _ in range(min(args.warmup, 1)):
            sim.run_staged(seed=1)
        for _ in range(args.steps):
            c = sim.run_staged(seed=1)
            t_dev += c['ms_total'] / 1e3
            launches += c['gpu_launches']
            phases += np.array(c['ms_phase'])
        caps[k] = int(c['n_records_total'] * 1.02) + 1024
        for key in keys:
            tot[key] += c[key]
        ph_alone += np.array(one_lane(lambda: sim.run_staged(seed=1))['ms_phase'])
    clk = sampler.stop() if sampler else None
    D.barrier()
    # end-to-end leg
    arena = host_array(max(caps.values()) if caps else 1, raw_record_dtype())
    pinned = sim.pin(arena)
    hashes, t_e2e, d2h, h2d = {}, 0.0, 0, 0
    for step in range(1 + args.steps):        # one warm-up pass over the rank's chunks
        D.barrier()
        t0 = time.perf_counter()
        for k in mine:
            idx = parts[k]
            out = sim.simulate(inst[idx], seed=1, records_out=arena, rng_id=idx.astype(np.uint64))
            if step == 1:
                lc = sim.last_counts
                d2h += lc['d2h_bytes'] + lc['n_truth'] * TRUTH_BYTES + lc['n_groups'] * 24
                h2d += inst[idx].nbytes + len(idx) * 28
                pre, t_cut = prefix_of(idx)
                rec = out['raw_records']
                sel = rec if t_cut is None else rec[rec['time'] < t_cut - int(cfg['right_raw_extension'])]
                hashes[k] = (len(pre), len(sel), hashlib.md5(sel.tobytes()).hexdigest())
            del out
        D.torch.cuda.synchronize()
        if step >= 1:
            t_e2e += time.perf_counter() - t0
    D.barrier()
    all_hashes = {}
    for h in D.gather_objects(hashes):
        all_hashes.update(h)
    check = None
    if rank == 0:
        bad = 0
        for k, (n_pre, n_sel, md5) in sorted(all_hashes.items()):
            idx = parts[k][:n_pre]
            out = sim.simulate(inst[idx], seed=1, rng_id=idx.astype(np.uint64))
            bad += int(hashlib.md5(out['raw_records'].tobytes()).hexdigest() != md5 or len(out['raw_records']) != n_sel)
        check = {'chunks_checked': len(all_hashes), 'mismatching': bad,
                 'what': 'records of a closed prefix (>= 400 instructions, cut at a quiet gap) of every chunk, as produced inside '
                         'the chunk run on its rank, byte-compared (md5) with a separate single-GPU call of that prefix on rank 0'}
    # reduce
    t_dev_m, t_e2e_m = D.reduce([t_dev, t_e2e], 'max')
    sums = D.reduce([tot['n_pe'], tot['n_records_total'], launches, tot['n_photons'], tot['n_truth'], tot['n_instructions']])
    n_pe_all, n_rec_all, launches_all, n_ph_all, n_tr_all, n_in_all = sums
    peak, peak_src = measured_peak()
    balg = PHOTON_BYTES * n_ph_all + RECORD_BYTES * n_rec_all + TRUTH_BYTES * n_tr_all + INSTR_BYTES * n_in_all
    steps = args.steps
    s_dev, s_e2e = max(t_dev_m / steps, 1e-9), max(t_e2e_m / steps, 1e-9)
    line = {
        'metric': 'photoelectrons_per_s', 'value': n_pe_all / s_dev,
        'unit': 'pe/s', 'raw_records_gbs': n_rec_all * RECORD_BYTES / s_dev / 1e9,
        'n_gpus': D.world, 'steps': steps, 'warmup': min(args.warmup, 1), 'ms_per_step': s_dev * 1e3,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOADS['C3'], 'events': n_events, 'instructions': int(len(inst)), 'chunks': n_chunks,
                   'chunks_on_rank0': len(mine), 'photons_per_step': int(n_ph_all), 'records_per_step': int(n_rec_all),
                   'device_batches_on_rank0': int(tot['n_batches']), 'fused_batches_on_rank0': int(tot['n_fused_batches']),
                   'l2': 'inputs and outputs of every batch exceed the 126 MB L2 (no flush needed)'},
        'e2e': {'value': n_pe_all / s_e2e, 'unit': 'pe/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': s_e2e * 1e3,
                'raw_records_gbs': n_rec_all * RECORD_BYTES / s_e2e / 1e9,
                'destination': 'one recycled record arena per rank, ' + ('page-locked' if pinned else 'pageable')},
        'gpu_launches': int(launches_all),
        'ms_phase_per_step': dict(zip(PHASE_NAMES, np.round(phases[:8] / steps, 3).tolist())),
        'path_hbm': {'algorithmic_bytes_per_step': int(balg), 'achieved_gbs': balg / s_dev / 1e9,
                     'frac_of_measured_peak': balg / s_dev / 1e9 / peak},
        'shard_check': check,
    }
    if clk is not None:
        line['clocks'] = clk
    ctot = dict(tot, n_batches=max(tot['n_batches'], 1))
    line['roofline'] = roofline_of(ctot, ph_alone, phases / steps, peak, peak_src)
    del arena
    sim.close()
    return line


def stretched_templates(cfg, taps):
    """The shipped single-pe pulse shape stretched in time to `taps` = (samples before, samples after the pulse centre)
    (pulse.py:146-187: the template spans before + after samples; shipped: 2 + 20)."""
    before, after = taps
    if (before, after) == (2, 20):
        return cfg
    cfg = dict(cfg)
    cfg['pe_pulse_ts'] = (np.asarray(cfg['pe_pulse_ts'], float) * (before + after) / 22.0).tolist()
    cfg['samples_before_pulse_center'], cfg['samples_after_pulse_center'] = before, after
    return cfg


def run_c4(args, D, rank, local_rank, n_pulses, per_pulse=1_000_000, pulses_per_call=50,
           taps=((2, 20), (1, 10), (3, 29), (4, 40))):
    """Photons generated on the device (torch: plumbing for the synthetic input), pulses split over the ranks
    (weak: every rank its own `n_pulses`).  value: photons/s with the photons resident (on_device call, CUDA-event
    time of the call); e2e: the same call with host arrays in and a host record buffer out."""
    import ctypes as C
    import torch
    from wfsim_b200 import lib as wlib
    from wfsim_b200.dtypes import raw_record_dtype
    from wfsim_b200.simulator import Simulator, _ptr
    base = load_config(zle_threshold=0)
    peak, peak_src = measured_peak()
    dev = torch.device('cuda', local_rank)
    results = {}
    for tp in taps:
        factor = sum(tp)
        cfg = stretched_templates(base, tp)
        try:
            sim = Simulator(cfg, device=local_rank)
        except Exception as e:      # the library takes templates of up to 32 samples (k_digitize's owner layers): 44 is refused
            results[f'taps_{factor}'] = {'error': str(e)[:200]}
            continue
        gains = torch.tensor(np.asarray(cfg['gains'], np.float64), device=dev)
        K = min(pulses_per_call, n_pulses)
        n = K * per_pulse
        gen = torch.Generator(device=dev)
        gen.manual_seed(1234 + rank)

        def make(k0):
            t0 = (torch.arange(k0, k0 + K, device=dev, dtype=torch.int64) * 1_000_000).repeat_interleave(per_pulse)
            t = t0 + torch.round(torch.randn(n, generator=gen, device=dev, dtype=torch.float64) * 1000.0).to(torch.int64) + 100_000
            top = torch.rand(n, generator=gen, device=dev) < 0.627
            ch = torch.where(top, torch.randint(0, 253, (n,), generator=gen, device=dev),
                             torch.randint(253, 494, (n,), generator=gen, device=dev)).to(torch.int32)
            g = gains[ch.long()] * (0.3 - 0.7 * torch.log1p(-torch.rand(n, generator=gen, device=dev, dtype=torch.float64)))
            pc = torch.arange(K, device=dev, dtype=torch.int32).repeat_interleave(per_pulse)
            return t, ch, g, pc
        group_of = np.arange(K, dtype=np.int32)
        groups = np.zeros(K, dtype=[('left', np.int64), ('right', np.int64), ('n_intervals', np.int64)])
        cap = 80 * K * 494
        d_rec = torch.empty(cap * RECORD_BYTES, dtype=torch.uint8, device=dev)
        counts = wlib.Counts()

        def call_device(t, ch, g, pc):
            rc = sim.lib.wfs_simulate_photons(sim.handle, n, C.c_void_p(t.data_ptr()), C.c_void_p(ch.data_ptr()),
                                              C.c_void_p(g.data_ptr()), C.c_void_p(pc.data_ptr()), K, _ptr(group_of), K,
                                              None, 1, 1, C.c_void_p(d_rec.data_ptr()), cap, _ptr(groups), C.byref(counts))
            if rc != 0:
                sim._raise(rc)
            return counts.as_dict()
        try:
            n_calls = max(n_pulses // K, 1)
            batches = [make(i * K) for i in range(min(n_calls, 2))]      # two resident input sets, alternated (1.2 GB each > L2)
            torch.cuda.synchronize()
            for _ in range(max(args.warmup, 1)):
                call_device(*batches[0])
            D.barrier()
            ms, ms_digi, launches, nrec, nsamp = 0.0, 0.0, 0, 0, 0
            phases = np.zeros(12)
            for step in range(args.steps):
                for i in range(n_calls):
                    c = call_device(*batches[i % len(batches)])
                    ms += c['ms_total']; ms_digi += c['ms_digitize']; launches += c['gpu_launches']
                    nrec, nsamp = c['n_records_total'], c['n_samples']
                    phases += np.array(c['ms_phase'])
            D.barrier()
            # e2e: host arrays in, host records out
            t, ch, g, pc = (x.cpu().numpy() for x in batches[0])
            rec_host = host_array(int(nrec * 1.02) + 1024, raw_record_dtype())
            sim.pin(rec_host)
            e2e_s, d2h_bytes = 0.0, 0
            for step in range(1 + args.steps):
                D.barrier()
                t0 = time.perf_counter()
                for i in range(n_calls):
                    rc = sim.lib.wfs_simulate_photons(sim.handle, n, _ptr(t), _ptr(ch), _ptr(g), _ptr(pc), K, _ptr(group_of), K,
                                                      None, 1, 0, _ptr(rec_host), len(rec_host), _ptr(groups), C.byref(counts))
                    if rc != 0:
                        sim._raise(rc)
                torch.cuda.synchronize()
                if step >= 1:
                    e2e_s += time.perf_counter() - t0
                d2h_bytes = int(counts.d2h_bytes) * n_calls
            t_dev, t_e2e = D.reduce([ms / 1e3, e2e_s], 'max')
            n_ph_all, launches_all = D.reduce([n * n_calls * args.steps, launches])
            byt = PHOTON_BYTES * n * n_calls + 2 * nsamp * n_calls          # photons in, int16 window samples out
            s_digi = ms_digi / args.steps / 1e3
            results[f'taps_{factor}'] = {
                'template_taps': int(sim.params.template_length), 'photons_per_s': n_ph_all / t_dev, 'ms_per_step': t_dev / args.steps * 1e3,
                'e2e_photons_per_s': n_ph_all / t_e2e, 'e2e_ms_per_step': t_e2e / args.steps * 1e3,
                'h2d_bytes_per_step': int(PHOTON_BYTES * n * n_calls), 'd2h_bytes_per_step': d2h_bytes,
                'records_per_call': int(nrec), 'window_samples_per_call': int(nsamp), 'gpu_launches': int(launches_all),
                'ms_phase_per_step': dict(zip(PHASE_NAMES, np.round(phases[:8] / args.steps, 3).tolist())),
                'path_hbm_frac': byt / (t_dev / args.steps) / 1e9 / peak,
                'k_digitize': {'ms_per_step': ms_digi / args.steps,
                               'achieved_gbs': byt / s_digi / 1e9 if s_digi > 0 else None,
                               'frac': byt / s_digi / 1e9 / peak if s_digi > 0 else None}}
            del batches, rec_host
        except Exception as e:
            results[f'taps_{factor}'] = {'error': f'{type(e).__name__}: {str(e)[:200]}'}
        del d_rec
        sim.close()
        torch.cuda.empty_cache()
    first = results.get('taps_22', {})
    line = {
        'metric': 'photons_per_s', 'value': first.get('photons_per_s'), 'unit': 'photons/s', 'n_gpus': D.world,
        'steps': args.steps, 'warmup': max(args.warmup, 1), 'ms_per_step': first.get('ms_per_step'), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic (generated on the device)',
        'config': {'workload': WORKLOADS['C4'], 'pulses_per_gpu': n_pulses, 'photons_per_pulse': per_pulse,
                   'pulses_per_call': min(pulses_per_call, n_pulses),
                   'l2': 'two resident input sets of 1.2 GB each are alternated (larger than the 126 MB L2)'},
        'e2e': {'value': first.get('e2e_photons_per_s'), 'unit': 'photons/s', 'h2d_bytes_per_step': first.get('h2d_bytes_per_step'),
                'd2h_bytes_per_step': first.get('d2h_bytes_per_step')},
        'gpu_launches': first.get('gpu_launches'),
        'roofline': {'bound': 'hbm', 'kernel': 'k_digitize', 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                     'achieved': (first.get('k_digitize') or {}).get('achieved_gbs'), 'frac': (first.get('k_digitize') or {}).get('frac'),
                     'traffic': None},
        'template_sweep': results,
    }
    return line


def brief(line):
    """What a config contributes to the `configs` block of the default line."""
    keep = ('metric', 'value', 'unit', 'ms_per_step', 'raw_records_gbs', 'scaling', 'config', 'e2e', 'path_hbm', 'roofline',
            'shard_check', 'template_sweep', 'ms_phase_per_step', 'gpu_launches')
    return {k: line[k] for k in keep if k in line}


def run_b200(args, rank, world, local_rank):
    D = Dist(world, local_rank)
    name = args.config
    if name == 'C2':
        line = run_c2(args, D, rank, local_rank, args.events or 2000)
    elif name == 'C3':
        line = run_c3(args, D, rank, local_rank, args.events or 1_000_000)
    elif name == 'C4':
        line = run_c4(args, D, rank, local_rank, args.events or 1000)
    else:
        sim, cfg = make_sim('C1', local_rank)
        n_events = args.events or 100000
        inst = workload(n_events, seed=100 + rank)
        r = measure_path(sim, inst, args, D, clocks_gpu=local_rank)
        line = path_line('C1', r, D, args, extra_config={'events_per_gpu': n_events})
        sim.close()
        if world == 1 and rank == 0 and not args.no_configs:
            # bounded samples of the other BASELINE.json configs (full sizes: --config C2|C3|C4)
            small = argparse.Namespace(**vars(args))
            small.steps, small.warmup = 2, 1
            cfgs = {}
            for key, fn in (('C2', lambda: run_c2(small, D, rank, local_rank, 300, clocks=False)),
                            ('C3', lambda: run_c3(small, D, rank, local_rank, 40000, clocks=False)),
                            ('C4', lambda: run_c4(small, D, rank, local_rank, 50, pulses_per_call=50))):
                try:
                    cfgs[key] = brief(fn())
                except Exception as e:      # a sample that fails must not take the C1 line with it
                    cfgs[key] = {'error': f'{type(e).__name__}: {str(e)[:300]}'}
            line['configs'] = cfgs
        if world == 1 and rank == 0 and not args.no_cpu_baseline:
            workers = len(os.sched_getaffinity(0))
            port = cpu_baseline_port(args.port_events, workers)
            if reference_available():
                line['cpu_baseline'] = cpu_baseline_reference(args.ref_events, workers)
                line['cpu_baseline']['port'] = port       # the restatement under oracle/, as a second stated figure
            else:
                line['cpu_baseline'] = port
    if rank == 0:
        print(json.dumps(line), flush=True)
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C1', choices=['C1', 'C2', 'C3', 'C4'])
    ap.add_argument('--events', type=int, default=int(os.environ.get('WFS_BENCH_EVENTS', 0)),
                    help='events per GPU (C1: 100000, C2: 2000), events of the stream (C3: 1000000), pulses per GPU (C4: 1000)')
    ap.add_argument('--ref-events', type=int, default=int(os.environ.get('WFS_BENCH_REF_EVENTS', 200)),
                    help='events per worker process and step of the reference CPU arm')
    ap.add_argument('--port-events', type=int, default=int(os.environ.get('WFS_BENCH_PORT_EVENTS', 1500)),
                    help='events per worker process of the oracle-port CPU figure')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='C1 only: skip the bounded C2 / C3 / C4 samples')
    ap.add_argument('--no-e2e', action='store_true', help='device-resident leg only (diagnostics)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
