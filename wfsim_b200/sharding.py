"""Chunk/shard-level parallelism over the GPUs of one box: events are independent,
so the instruction set is cut into contiguous time ranges at cluster gaps, one range per GPU, and
the per-GPU results are concatenated in time order on the host.  No collective is involved."""
import numpy as np


def signal_time(instructions, drift_velocity_liquid):
    """rawdata.py:61 (float32 arithmetic as numpy evaluates it)."""
    zf = instructions['z'].astype(np.float32) / np.float32(drift_velocity_liquid)
    k = (instructions['type'].astype(np.int8) % 2 - 1).astype(np.float32)
    return instructions['time'].astype(np.int64) + (zf * k).astype(np.int64)


def shard_instructions(instructions, n_shards, config, min_gap=None):
    """Indices of `instructions` per shard: contiguous in signal time, cut only at gaps larger than
    `min_gap` (default: right_raw_extension plus the longest photo-ionisation delay when electron
    afterpulses are enabled), balanced by sum(amp) as a photon-count proxy.  Shards may be empty."""
    n = len(instructions)
    if n == 0:
        return [np.zeros(0, np.int64) for _ in range(n_shards)]
    st = signal_time(instructions, config['drift_velocity_liquid'])
    order = np.argsort(st, kind='stable')
    if min_gap is None:
        min_gap = int(config.get('right_raw_extension', 100000))
        if config.get('enable_electron_afterpulses', False):
            min_gap += int(config.get('tpc_length', 150) / config['drift_velocity_liquid']) + 100000
    gaps = np.diff(st[order])
    cut_ok = np.flatnonzero(gaps > min_gap) + 1            # positions where a cut is allowed
    w = np.cumsum(instructions['amp'][order].astype(np.float64))
    total = w[-1]
    cuts = []
    for k in range(1, n_shards):
        if len(cut_ok) == 0:
            break
        target = total * k / n_shards
        pos = int(np.searchsorted(w, target))
        j = int(np.argmin(np.abs(cut_ok - pos)))
        cuts.append(int(cut_ok[j]))
    cuts = sorted(set(cuts))
    parts = np.split(order, cuts)
    parts += [np.zeros(0, np.int64)] * (n_shards - len(parts))
    return parts


def merge_results(results):
    """Concatenate per-shard outputs of Simulator.simulate (shards are disjoint in time and given in
    time order)."""
    keys = ('raw_records', 'raw_records_he', 'raw_records_aqmon', 'truth', 'groups')
    results = [r for r in results if r is not None]
    return {k: np.concatenate([r[k] for r in results]) for k in keys}


class ShardedSimulator:
    """One Simulator handle per visible GPU, driven from one process by a thread per device
    (ctypes releases the GIL during the library call)."""

    def __init__(self, config, resource=None, devices=None):
        from .simulator import Simulator
        from . import lib as wlib
        n = wlib.load().wfs_device_count()
        self.devices = list(range(n)) if devices is None else list(devices)
        if not self.devices:
            raise RuntimeError('no CUDA device visible: wfsim_b200 has no CPU fallback')
        self.config = config
        self.sims = [Simulator(config, resource=resource, device=d) for d in self.devices]

    def simulate(self, instructions, seed=0):
        from concurrent.futures import ThreadPoolExecutor
        parts = shard_instructions(instructions, len(self.sims), self.config)

        def run(k):
            idx = parts[k]
            if len(idx) == 0:
                return None
            return self.sims[k].simulate(instructions[idx], seed=seed, rng_id=idx.astype(np.uint64))
        with ThreadPoolExecutor(len(self.sims)) as ex:
            outs = list(ex.map(run, range(len(self.sims))))
        return merge_results(outs)

    def close(self):
        for s in self.sims:
            s.close()
