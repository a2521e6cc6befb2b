"""Chunk/shard-level parallelism over the GPUs of one box: events are independent,
so the instruction set is cut into contiguous time ranges at cluster gaps, one range per GPU, and
the per-GPU results are concatenated in time order on the host.  No collective is involved."""
import numpy as np


def signal_time(instructions, drift_velocity_liquid):
    """rawdata.py:61 (float32 arithmetic as numpy evaluates it)."""
    zf = instructions['z'].astype(np.float32) / np.float32(drift_velocity_liquid)
    k = (instructions['type'].astype(np.int8) % 2 - 1).astype(np.float32)
    return instructions['time'].astype(np.int64) + (zf * k).astype(np.int64)


def default_quiet_gap(config):
    """Upper bound of Simulator.quiet_gap() (wfs_quiet_gap) from the config alone, for callers that hold
    no handle: the clustering gap plus the longest delay of a secondary an S2 can spawn -- photo-ionisation
    electrons up to the end of the delay histogram of the afterpulse file (bounded here by 1 ms or two full
    drift lengths, whichever is longer), photo-electric electrons 6 sigma behind their centre."""
    gap = int(config.get('right_raw_extension', 100000))
    extra = 0
    if config.get('enable_electron_afterpulses', False):
        extra = max(int(2 * config.get('tpc_length', 150) / config['drift_velocity_liquid']), 1000000) + 50000
    if config.get('enable_gate_afterpulses', False):
        extra = max(extra, int(config.get('photoelectric_t_center', 0) + config.get('drift_time_gate', 0)
                               + 6 * config.get('photoelectric_t_spread', 0)) + 50000)
    # photons trail the signal time by the S2 width and the PMT-afterpulse delays (the library reads the longest
    # delay from the afterpulse tables; without them: 100 us)
    reach = 50000 + (100000 if config.get('enable_pmt_afterpulses', False) else 0)
    return gap + extra + reach


def shard_instructions(instructions, n_shards, config, min_gap=None):
    """Indices of `instructions` per shard: contiguous in signal time, cut only at gaps larger than
    `min_gap` (Simulator.quiet_gap(): what the library cuts its device batches at; default:
    default_quiet_gap(config)), balanced by sum(amp) as a photon-count proxy.  Shards may be empty."""
    n = len(instructions)
    if n == 0:
        return [np.zeros(0, np.int64) for _ in range(n_shards)]
    st = signal_time(instructions, config['drift_velocity_liquid'])
    order = np.argsort(st, kind='stable')
    if min_gap is None:
        min_gap = default_quiet_gap(config)
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
    out = {k: np.concatenate([r[k] for r in results]) for k in keys}
    for k in keys[:3]:      # shards are cut at quiet gaps; should one ever reach back, restore the order
        r = out[k]
        if len(r) > 1 and (np.diff(r['time']) < 0).any():
            out[k] = r[np.lexsort((r['channel'], r['time']))]
    return out


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
        parts = shard_instructions(instructions, len(self.sims), self.config, min_gap=self.sims[0].quiet_gap())

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
