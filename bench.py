#!/usr/bin/env python
"""Benchmark of the WFSim hot path (wfsim_instructions -> raw_records) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

Workload (config.workload): BASELINE.json config[1] -- 1e5 low-energy (1-50 keV) events at 1 kHz,
XENONnT 494 channels, fax_config = the reference's shipped test config (dummy maps), synthetic
instructions (tests/golden/synth_instructions.py:c1_like).  One step = one pass of the whole path
over those instructions.  Weak scaling: every rank simulates its own 1e5 events (different
instruction seed), no data-path collective (events are independent).

Prints ONE JSON line (rank 0).  value = photoelectrons/s with the instructions already planned
and resident (wfs_stage_instructions / wfs_run_staged; records stay in HBM); e2e = same metric
through the public API with host buffers (instructions H2D, records + truth D2H into pinned
memory) inside the timed region.
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


def load_config():
    from tests.conftest import load_c0_config
    return load_c0_config()


def spe_tables():
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'c0_tables.npz'))
    return z['spe_unique'], z['spe_row'][:494]


def workload(n_events, seed):
    from tests.golden.synth_instructions import c1_like
    return c1_like(n_events, seed=seed)


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (None if there is none)."""
    path = os.path.join(ROOT, 'profiles', f'r1p_{kernel}_ncu.json')
    try:
        with open(path) as f:
            return float(json.load(f)['dram_bytes_per_launch'])
    except (OSError, KeyError, ValueError):
        return None


def host_array(n, dtype):
    """Ordinary pageable host array for the records (what a plugin would allocate), on transparent
    huge pages where the kernel offers them, touched once so that page faults are not timed."""
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
            self._halt.wait(0.2)

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
# CPU baseline: the oracle port of the reference path (single-threaded like the reference;
# P worker processes over disjoint event slices)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
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


def cpu_baseline(events_per_worker, workers):
    import multiprocessing as mp
    from oracle import wfsim_oracle as orc
    orc.build()
    _cpu_worker((2, 999))          # warm caches / library load outside the timed runs
    t0 = time.perf_counter()
    if workers > 1:
        with mp.get_context('fork').Pool(workers) as pool:
            res = pool.map(_cpu_worker, [(events_per_worker, 1000 + w) for w in range(workers)])
    else:
        res = [_cpu_worker((events_per_worker, 1000))]
    wall = time.perf_counter() - t0
    n_pe = sum(r['n_pe'] for r in res)
    n_rec = sum(r['n_records'] for r in res)
    return dict(value=n_pe / wall, unit='pe/s', cores=workers, kind='port',
                sample=f'{events_per_worker} events x {workers} worker processes of the same '
                       f'C1 workload ({wall:.1f} s wall)',
                raw_records_gbs=n_rec * RECORD_BYTES / wall / 1e9, seconds=wall)


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    workers = len(os.sched_getaffinity(0))
    per = max(2, int(args.ref_events))
    vals, recs, secs = [], [], []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(per, workers)
        if i >= args.warmup:
            vals.append(b['value']); recs.append(b['raw_records_gbs']); secs.append(b['seconds'])
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': 'photoelectrons_per_s', 'value': v, 'unit': 'pe/s',
        'raw_records_gbs': float(np.mean(recs)),
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': float(np.mean(secs)) * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_NAME, 'sample_events_per_step': per * workers},
        'cpu_baseline': {'value': v, 'unit': 'pe/s', 'cores': workers, 'kind': 'port',
                         'sample': f'{per} events x {workers} processes per step; oracle/ port of the '
                                   'reference path (the reference is pure Python + numba and cannot be '
                                   'compiled into oracle/_ref)'},
        'e2e': {'value': v, 'unit': 'pe/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


WORKLOAD_NAME = ('C1: low-energy S1+S2 recoils 1-50 keV, 1 kHz, XENONnT 494 ch, '
                 'XENONnT_wfsim_config.json + test_load_nt dummy maps')


def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from wfsim_b200.resource import Resource
    from wfsim_b200.simulator import Simulator, PinnedArray
    from wfsim_b200 import lib as wlib
    from wfsim_b200.dtypes import raw_record_dtype
    import ctypes as C

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    cfg = load_config()
    uniq, row = spe_tables()
    res = Resource(cfg, spe_ppf=uniq, spe_row=row)
    sim = Simulator(cfg, resource=res, device=local_rank)
    inst = workload(args.events, seed=100 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg -----------------------------------------------------------------
    sim.stage(inst)
    for _ in range(args.warmup):
        c = sim.run_staged(seed=1)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, ms_digi, launches = [], [], 0
    phases = np.zeros(12)
    for k in range(args.steps):
        c = sim.run_staged(seed=1)
        ms.append(c['ms_total']); ms_digi.append(c['ms_digitize']); launches += c['gpu_launches']
        phases += np.array(c['ms_phase'])
    barrier()
    clocks = sampler.stop()
    t_dev = float(np.sum(ms)) / 1e3
    if world > 1:
        tt = torch.tensor([t_dev], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev = float(tt.item())
        tot = torch.tensor([c['n_pe'], c['n_records_total'], launches], device='cuda', dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        n_pe_all, n_rec_all, launches_all = (float(x) for x in tot.tolist())
    else:
        n_pe_all, n_rec_all, launches_all = float(c['n_pe']), float(c['n_records_total']), float(launches)
    value = n_pe_all * args.steps / t_dev
    rec_gbs = n_rec_all * RECORD_BYTES * args.steps / t_dev / 1e9

    # ---- kernel-alone pass for the roofline: one lane, so that the CUDA events around k_digitize
    # bracket that kernel only (with two lanes the other lane's kernels share the GPU with it) ----
    lanes_before = os.environ.get('WFS_LANES')
    os.environ['WFS_LANES'] = '1'
    try:
        sim.run_staged(seed=1)
        ms_digi_alone = [sim.run_staged(seed=1)['ms_digitize'] for _ in range(args.steps)]
    finally:
        if lanes_before is None:
            del os.environ['WFS_LANES']
        else:
            os.environ['WFS_LANES'] = lanes_before
    barrier()

    # ---- end-to-end leg: public API, host buffers, H2D + D2H inside the timed region ----------
    cap = int(c['n_records_total'] * 1.02) + 1024
    # caller-owned destination: an ordinary (pageable) numpy array, touched once, as a plugin would
    # hold it -- the library's host threads expand the compact records straight into it
    records_out = host_array(cap, raw_record_dtype())
    dest_pinned = (not os.environ.get('WFS_BENCH_PAGEABLE')) and sim.pin(records_out)
    e2e_ms, e2e_lib = [], []
    h2d = inst.nbytes + 494 * 4 * 2 + len(inst) * (8 * 3 + 4)
    d2h = 0
    out = None
    for k in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        out = sim.simulate(inst, seed=1, cap_records=cap, records_out=records_out)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if k >= args.warmup:
            e2e_ms.append(dt * 1e3)
            e2e_lib.append([sim.last_counts['ms_total']] + list(sim.last_counts['ms_phase'][8:10])
                           + [sim.last_counts['n_plain_records']])
        # record bytes that crossed PCIe (compact transport: headers + non-baseline sample blocks,
        # expanded to 244-byte records by the library's host threads) + truth rows + group info
        d2h = (sim.last_counts['d2h_bytes'] + sim.last_counts['n_truth'] * TRUTH_BYTES
               + sim.last_counts['n_groups'] * 24)
    t_e2e = float(np.sum(e2e_ms)) / 1e3
    if world > 1:
        tt = torch.tensor([t_e2e], device='cuda', dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
    e2e_value = n_pe_all * args.steps / t_e2e

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peak()
    balg = algorithmic_bytes(c)
    digi_bytes = PHOTON_BYTES * c['n_photons'] + 2 * c['n_samples']
    digi_ms = float(np.mean(ms_digi_alone))
    digi_ms_lanes = float(np.mean(ms_digi))
    line = {
        'metric': 'photoelectrons_per_s', 'value': value, 'unit': 'pe/s',
        'raw_records_gbs': rec_gbs,
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_dev / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_NAME, 'events_per_gpu': args.events,
                   'instructions_per_gpu': int(len(inst)), 'photons_per_step_per_gpu': int(c['n_photons']),
                   'records_per_step_per_gpu': int(c['n_records_total']),
                   'device_batches_per_step': int(c['n_batches']),
                   'l2': 'inputs and outputs of every batch exceed the 126 MB L2 (no flush needed)'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'pe/s', 'h2d_bytes_per_step': int(h2d),
                'd2h_bytes_per_step': int(d2h), 'ms_per_step': t_e2e / args.steps * 1e3,
                'raw_records_gbs': n_rec_all * RECORD_BYTES * args.steps / t_e2e / 1e9,
                # inside the call (rank 0, per step): device work of all batches; wall clock summed over
                # batches of (batch shipped -> its compact D2H done) and (D2H done -> expanded by host threads)
                'ms_device': float(np.mean([x[0] for x in e2e_lib])),
                'ms_batches_d2h': float(np.mean([x[1] for x in e2e_lib])),
                'ms_batches_expand': float(np.mean([x[2] for x in e2e_lib])),
                # split transport: the destination is the caller's numpy array, page-locked with wfs_host_register
                # as the plugin's record arenas are; this share of the records arrived as plain rows by DMA
                'destination': 'caller-owned numpy array, ' + ('page-locked (wfs_host_register)' if dest_pinned else 'pageable'),
                'plain_record_share': [round(x[3] / max(c['n_records_total'], 1), 3) for x in e2e_lib]},
        'gpu_launches': int(launches_all),
        'ms_phase_per_step': dict(zip(['frontend', 'photon_sort', 'windows', 'digitize', 'zle', 'record_sort',
                                       'record_pack', 'host_scheduler_truth'], (phases[:8] / args.steps).round(3).tolist())),
        'path_hbm': {'algorithmic_bytes_per_step': int(balg),
                     'achieved_gbs': balg / (t_dev / args.steps) / 1e9,
                     'frac_of_measured_peak': balg / (t_dev / args.steps) / 1e9 / peak},
        'roofline': {'bound': 'hbm', 'kernel': 'k_digitize',
                     'achieved': digi_bytes / (digi_ms / 1e3) / 1e9 if digi_ms > 0 else None,
                     'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                     'frac': (digi_bytes / (digi_ms / 1e3) / 1e9 / peak) if digi_ms > 0 else None,
                     # DRAM bytes per launch (read + write) of the committed ncu --set full capture
                     'traffic': ncu_traffic('k_digitize'),
                     'traffic_source': 'profiles/r1p_k_digitize_ncu.json (one launch = one device batch of the same '
                                       'size class as here; dram__bytes_read.sum + dram__bytes_write.sum)',
                     'launches_per_step': int(c['n_batches']),
                     'algorithmic_bytes_per_launch': int(digi_bytes / max(int(c['n_batches']), 1)),
                     'algorithmic_bytes_per_step': int(digi_bytes), 'kernel_ms_per_step': digi_ms,
                     'timing': 'CUDA events on the library stream around every k_digitize launch, summed per step, '
                               'in a pass of the same steps with one lane (kernel alone on the GPU)',
                     'kernel_ms_per_step_in_timed_region': digi_ms_lanes},
    }
    if world == 1 and not args.no_cpu_baseline:
        workers = len(os.sched_getaffinity(0))
        line['cpu_baseline'] = cpu_baseline(args.ref_events, workers)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--events', type=int, default=int(os.environ.get('WFS_BENCH_EVENTS', 100000)))
    ap.add_argument('--ref-events', type=int, default=int(os.environ.get('WFS_BENCH_REF_EVENTS', 1500)),
                    help='events per worker process and step of the CPU baseline')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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
