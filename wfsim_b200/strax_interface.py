"""Drop-in host side of the hot path: the interface of wfsim/strax_interface.py with the
simulator swapped for the B200 library.

Mirrors (names, argument meaning, error behaviour; citations relative to the reference):
* `ChunkRawRecords`        strax_interface.py:353-504   (chunk cutting, truth selection)
* `SimulatorPlugin`        strax_interface.py:506-663   (options, set_config, _sort_check, is_ready)
* `RawRecordsFromFaxNT`    strax_interface.py:665-714   (provides, _setup, check_instructions, compute)
* `instruction_from_csv`   strax_interface.py:336-350
* dtypes                   strax_interface.py:25-116     (wfsim_b200/dtypes.py)

With strax/straxen installed the plugin classes derive from strax.Plugin and register in a
strax.Context exactly like the reference's; without them (this build container) a minimal
stand-in base class is used so the same compute loop can be exercised by the tests.
The simulation itself always runs in libwfsim_b200.so (ctypes); there is no CPU fallback.
"""
import logging
import os

import numpy as np

from . import config as wcfg
from .dtypes import (RECORD_LENGTH, extra_truth_dtype_per_pmt, instruction_dtype, optical_extra_dtype, raw_record_dtype,
                     truth_extra_dtype)
from .resource import Resource, resource_from_reference  # noqa: F401  (re-exported for deployments)

log = logging.getLogger('wfsim_b200.interface')

try:                                    # pragma: no cover - depends on the deployment
    import strax
    import straxen
    HAVE_STRAX = True
except ImportError:
    strax = straxen = None
    HAVE_STRAX = False

__all__ = ['instruction_dtype', 'truth_extra_dtype', 'extra_truth_dtype_per_pmt', 'ChunkRawRecords',
           'SimulatorPlugin', 'RawRecordsFromFaxNT', 'RawRecordsFromFaxOpticalNT', 'optical_extra_dtype',
           'instruction_from_csv', 'chunk_boundaries', 'ChunkClock']


def instruction_from_csv(filename):
    """Return wfsim instructions from a csv (strax_interface.py:336-350)."""
    import pandas as pd
    df = pd.read_csv(filename)
    recs = np.zeros(len(df), dtype=instruction_dtype)
    for column in df.columns:
        recs[column] = df[column]
    expected_dtype = np.dtype(instruction_dtype)
    assert recs.dtype == expected_dtype, \
        f"CSV {filename} produced wrong dtype. Got {recs.dtype}, expected {expected_dtype}."
    return recs


class ChunkClock:
    """The chunk_time bookkeeping of ChunkRawRecords.__call__ (strax_interface.py:378-440) as a state
    machine that is fed the digitisation groups piece by piece.

    `feed(groups)` takes (left, right, n_intervals) per group in time order -- what the reference
    reads from RawData.left / RawData.right while iterating the ZLE intervals -- and returns the
    chunks [(chunk_time_pre, chunk_time), ...] that became complete; `finish()` returns the last one."""

    RECORD_BUFFER = 5000000      # records the reference's chunker holds (strax_interface.py:360)

    def __init__(self, config, t_min_instruction, time_zero=None, record_buffer=None):
        self.dt = config['sample_duration']
        self.rext = int(config['right_raw_extension'])
        self.cksz = int(config['chunk_size'] * 1e9)
        self.pre = (time_zero - self.rext) if time_zero else (int(t_min_instruction) - self.rext)
        self.ct = self.pre + self.cksz
        self.cur_right = self.last_right = 0
        self.buffer_length = int(record_buffer or self.RECORD_BUFFER)
        self.blevel = 0              # records held since the last chunk left (the reference's buffer level)
        self.borrowed = False        # peek() has made the cut that the next interval would have made

    def feed(self, groups, n_records=None):
        """`n_records`: records per group (all data types).  With it the record-buffer rule of
        strax_interface.py:409-418 applies: a group that does not fit behind the records held closes the chunk
        at the end of the previous group, whatever the chunk clock says."""
        out = []
        dt, rext = self.dt, self.rext
        for g, (left, right, n_itv) in enumerate(groups):
            for _ in range(max(int(n_itv), 0)):
                if right != self.cur_right:
                    self.last_right, self.cur_right = self.cur_right, right
                if self.borrowed:
                    # the reference closes at most ONE chunk per ZLE interval (strax_interface.py:401-410); the chunk
                    # this interval would have closed has left through peek() already
                    self.borrowed = False
                    continue
                if left * dt > self.ct + rext:
                    if (self.last_right + 1) * dt > self.ct:
                        self.ct += (self.last_right + 1) * dt - self.ct
                    out.append((self.pre, self.ct))
                    self.pre = self.ct
                    self.ct += self.cksz
                    self.blevel = 0      # every record of the earlier groups has left
                else:
                    break
            if n_records is not None and int(n_itv) > 0:
                n = int(n_records[g])
                if self.blevel + n > self.buffer_length:
                    log.warning('Chunck size too large, insufficient record buffer \n'
                                'No longer in sync if simulating nVeto with TPC \n'
                                'Consider reducing the chunk size')
                    self.ct = (self.last_right + 1) * dt
                    out.append((self.pre, self.ct))
                    self.pre = self.ct
                    self.ct += self.cksz
                    self.blevel = 0
                    if n > self.buffer_length:
                        # strax_interface.py:420-422 drops the pulses that do not fit ("skipping pulse"); a group
                        # of more than 5e6 records is not simulated here either way
                        raise ValueError('Pulse length too large, insufficient record buffer')
                self.blevel += n
        return out

    def peek(self, t_next):
        """Every group still to come starts after `t_next` [ns]: if the first of them is going to close the
        current chunk, close it now (same bounds: they depend on the groups seen so far only) so that its
        records can leave before the next piece is simulated."""
        if self.borrowed or t_next <= self.ct + self.rext:
            return []
        if (self.cur_right + 1) * self.dt > self.ct:      # cur_right is what last_right will be by then
            self.ct += (self.cur_right + 1) * self.dt - self.ct
        out = [(self.pre, self.ct)]
        self.pre = self.ct
        self.ct += self.cksz
        self.blevel = 0
        self.borrowed = True
        return out

    def finish(self):
        self.last_right = self.cur_right
        self.ct = max((self.last_right + 1) * self.dt, self.pre + self.dt)
        return (self.pre, self.ct)


def chunk_boundaries(config, t_min_instruction, groups, time_zero=None, n_records=None, record_buffer=None):
    """All chunks of one run: [(chunk_time_pre, chunk_time), ...] in the order the reference yields
    them (see ChunkClock)."""
    clock = ChunkClock(config, t_min_instruction, time_zero, record_buffer=record_buffer)
    out = clock.feed(groups, n_records)
    out.append(clock.finish())
    return out


class _ArenaPool:
    """Record arenas of one chunker.  take(n) hands out an arena of at least n records: one of its own
    that nobody else references any more (the consumer dropped every chunk view into it), else a new one."""

    def __init__(self, pin=None):
        self.arenas = []
        self.pin = pin          # page-locks a new arena (Simulator.pin): part of the records then arrive by plain DMA

    def take(self, n, dtype):
        import sys
        # an arena is free when the list slot is the only reference left (+ the call argument)
        free = [i for i in range(len(self.arenas)) if sys.getrefcount(self.arenas[i]) <= 2]
        fit = [i for i in free if len(self.arenas[i]) >= n]
        if fit:
            chosen = self.arenas[fit[0]]
            self.arenas = [x for i, x in enumerate(self.arenas) if i == fit[0] or i not in free]
            return chosen
        self.arenas = [x for i, x in enumerate(self.arenas) if i not in free]      # too small: let them go
        # generous, quantised sizes: the estimates move a little from chunk to chunk and an arena that is
        # released and allocated again pays for unmapping and for the page faults of fresh memory
        quantum = 1 << 22
        self.arenas.append(np.empty((int(n * 1.25) // quantum + 1) * quantum, dtype))
        if self.pin is not None and not os.environ.get('WFS_NO_PINNED_ARENAS'):
            self.pin(self.arenas[-1])
        return self.arenas[-1]


def _pieces(instructions, config, piece_max, time_zero=None, min_gap=None):
    """Index arrays of the pieces the run is simulated in: contiguous in signal time, cut only at quiet
    gaps (sharding.shard_instructions explains which), first where the chunk clock will cut -- so that the
    records of a chunk come from one library call and are handed on without a copy -- then further down to
    at most ~piece_max instructions."""
    from .sharding import signal_time
    n = len(instructions)
    st = signal_time(instructions, config['drift_velocity_liquid'])
    order = np.argsort(st, kind='stable')
    sts = st[order]
    if min_gap is None:
        from .sharding import default_quiet_gap
        min_gap = default_quiet_gap(config)
    cut_ok = np.flatnonzero(np.diff(sts) > min_gap) + 1
    cuts = set()
    if len(cut_ok):
        rext, cksz = int(config['right_raw_extension']), int(config['chunk_size'] * 1e9)
        t0 = (time_zero - rext) if time_zero else int(np.min(instructions['time'])) - rext
        for t_cut in np.arange(t0 + cksz, sts[-1], max(cksz, 1)):
            j = np.searchsorted(cut_ok, np.searchsorted(sts, t_cut + rext, side='right'))
            if j < len(cut_ok):
                cuts.add(int(cut_ok[j]))
        edges = [0] + sorted(cuts) + [n]
        for a, b in zip(edges[:-1], edges[1:]):            # pieces that are still too big
            for k in range(1, -(-(b - a) // piece_max)):
                j = int(np.argmin(np.abs(cut_ok - (a + k * (b - a) // -(-(b - a) // piece_max)))))
                if a < cut_ok[j] < b:
                    cuts.add(int(cut_ok[j]))
    return [p for p in np.split(order, sorted(cuts)) if len(p)]


def _with_optical_columns(truth, instructions, dtype):
    """Truth rows of optical instructions get the `_first` / `_last` columns of the instruction they
    describe (one Pulse call per instruction; matched on the copied instruction fields)."""
    out = np.zeros(len(truth), dtype=dtype)
    for n in truth.dtype.names:
        out[n] = truth[n]
    key = ('time', 'type', 'amp', 'event_number', 'x', 'y', 'z')
    where = {}
    for j, row in enumerate(instructions):
        where.setdefault(tuple(row[k].item() for k in key), []).append(j)
    for i, row in enumerate(truth):
        js = where.get(tuple(row[k].item() for k in key))
        if js:
            j = js.pop(0) if len(js) > 1 else js[0]
            out['_first'][i], out['_last'][i] = instructions['_first'][j], instructions['_last'][j]
    return out


class ChunkRawRecords(object):
    """Same protocol as the reference class: calling the object with the instructions returns a
    generator of dict(raw_records, raw_records_he, raw_records_aqmon, truth); `chunk_time_pre`
    and `chunk_time` hold the bounds of the chunk just yielded; `source_finished()`."""

    def __init__(self, config, rawdata_generator=None, resource=None, device=0, seed=None, **kwargs):
        from .simulator import Simulator
        self.config = config
        if resource is None:
            resource = Resource(config, **{k: kwargs[k] for k in list(kwargs) if k in (
                'spe_ppf', 'spe_row', 'photon_area_distribution', 'noise_data', 'uniform_to_pmt_ap',
                'uniform_to_ele_ap')})
        self.simulator = rawdata_generator(config, resource=resource, device=device) \
            if rawdata_generator is not None else Simulator(config, resource=resource, device=device)
        truth_per_n_pmts = self._n_channels if config.get('per_pmt_truth') else False
        self.truth_dtype = extra_truth_dtype_per_pmt(truth_per_n_pmts)
        # externally supplied photons (ChunkRawRecords(..., rawdata_generator=RawDataOptical, channels=,
        # timings=), strax_interface.py:726-729): the instructions then carry `_first` / `_last`
        self.channels, self.timings = kwargs.get('channels'), kwargs.get('timings')
        if self.channels is not None:
            self.truth_dtype = optical_extra_dtype + self.truth_dtype
        self.seed = int(seed if seed is not None else (config.get('seed') or 0))
        self._arena_pool = _ArenaPool(pin=getattr(self.simulator, 'pin', None))
        self._finished = False
        self.chunk_time_pre = self.chunk_time = 0

    def __call__(self, instructions, time_zero=None, **kwargs):
        samples_per_record = RECORD_LENGTH
        if len(instructions) == 0:      # Empty (strax_interface.py:374-377)
            yield from np.array([], dtype=raw_record_dtype(samples_per_record=samples_per_record))
            self._finished = True
            return
        cfg = self.config
        # The run is simulated PIECE BY PIECE (contiguous in signal time, cut only at quiet gaps, so the
        # pieces are independent) and the chunks are yielded as soon as they are complete: host memory is
        # bounded by one piece (config 'b200_piece_instructions', default 40000) instead of the whole run.
        # Philox identities are the instruction indices and the noise draw of a digitisation group is keyed
        # by its first sample, so the records do not depend on how the run is cut.
        piece = max(int(cfg.get('b200_piece_instructions', 40000)), 1)
        # pieces are cut at the gaps the library itself cuts its device batches at (wfs_quiet_gap)
        quiet = self.simulator.quiet_gap() if hasattr(self.simulator, 'quiet_gap') else None
        parts = _pieces(instructions, cfg, piece, time_zero, min_gap=quiet)
        clock = ChunkClock(cfg, np.min(instructions['time']), time_zero, record_buffer=cfg.get('b200_record_buffer'))
        keys = ('raw_records', 'raw_records_he', 'raw_records_aqmon')
        side = {k: [] for k in keys[1:]}    # he / aqmon records not yet delivered (few): arrays in time order
        tdt = np.dtype(instruction_dtype + self.truth_dtype)
        rdt = raw_record_dtype(samples_per_record=samples_per_record)
        held_truth = None
        n_groups = 0
        empty = np.zeros(0, rdt)
        # TPC records live in an ARENA: every piece is simulated straight behind the records still held
        # (records_out = arena[used:]), chunks are views arena[a0:a0 + stop] -- no concatenation, no copy.
        # Arenas whose views the consumer has dropped are recycled (already-touched pages: the first-touch
        # page faults of fresh memory cost ten times the simulation).
        pool = self._arena_pool
        arena, a0, used = None, 0, 0
        per_instr = None                     # records per instruction, learnt from the pieces

        # one arena per chunk: it is changed right after a chunk has left (little to carry over), sized for
        # the instructions of one chunk; should the estimate prove too small it is doubled on the way
        span = float(np.max(instructions['time']) - np.min(instructions['time'])) + 1.0
        instr_per_chunk = int(len(instructions) * min(1.0, cfg['chunk_size'] * 1e9 / span)) + 1

        def room_for(n_instr, fresh=False):
            nonlocal arena, a0, used
            need = int((per_instr or 700.0) * n_instr * 1.3) + 65536
            if arena is not None and not fresh and len(arena) - used >= need:
                return
            held_n = used - a0
            new = pool.take(held_n + need if fresh else max(2 * held_n + need, 2 * need), rdt)
            if held_n:
                new[:held_n] = arena[a0:used]
            arena, a0, used = new, 0, held_n

        def cut(ct, last):
            """Records and truth rows up to chunk time `ct` leave the held arrays."""
            nonlocal held_truth, a0
            res = {}
            rec = arena[a0:used] if arena is not None else empty
            if len(rec) > 1 and not last and (np.diff(rec['time']) < 0).any():
                # pieces are cut at quiet gaps, so their records follow each other in time; should a piece
                # ever reach back behind its predecessor, restore the (time, channel) order in place
                rec[:] = rec[np.lexsort((rec['channel'], rec['time']))]
            stop = len(rec) if last else int(np.searchsorted(rec['time'], ct, side='right'))
            res['raw_records'] = rec[:stop]
            a0 += stop
            for k in keys[1:]:
                parts_k = [x for x in side[k] if len(x)]
                rec = empty if not parts_k else parts_k[0] if len(parts_k) == 1 else np.concatenate(parts_k)
                if len(rec) > 1 and (np.diff(rec['time']) < 0).any():
                    rec = rec[np.lexsort((rec['channel'], rec['time']))]
                stop = len(rec) if last else int(np.searchsorted(rec['time'], ct, side='right'))
                res[k], side[k] = rec[:stop], [rec[stop:]]
            # truth rows of this chunk (strax_interface.py:458-483)
            truth = held_truth
            tfp = truth['t_first_photon']
            sel = (tfp <= ct) | (np.isnan(tfp) & (truth['time'] <= ct))
            t, held_truth = truth[sel], truth[~sel]
            t = t[np.argsort(t['time'], kind='stable')]
            _truth = np.zeros(len(t), dtype=tdt)
            for name in _truth.dtype.names:
                _truth[name] = t[name]
            has = ~np.isnan(_truth['t_first_photon'])
            _truth['time'][has] = _truth['t_first_photon'][has].astype(int)
            res['truth'] = _truth[np.argsort(_truth['time'], kind='stable')]
            return res

        def deliver(res):
            if cfg['detector'] in ('XENON1T', 'XENONnT_neutron_veto'):
                return dict(raw_records=res['raw_records'], truth=res['truth'])
            return res

        delivered = True
        from .sharding import signal_time
        next_start = [signal_time(instructions[p[:1]], cfg['drift_velocity_liquid'])[0] for p in parts]
        for i_part, idx in enumerate(parts):
            optical = None if self.channels is None else (self.channels, self.timings)
            room_for(max(instr_per_chunk, len(idx)) if delivered else len(idx), fresh=delivered and arena is not None)
            delivered = False
            out = self.simulator.simulate(instructions[idx], seed=self.seed, rng_id=idx.astype(np.uint64),
                                          optical=optical, records_out=arena[used:])
            if optical is not None:
                out['truth'] = _with_optical_columns(out['truth'], instructions[idx], tdt)
            n_groups += len(out['groups'])
            rr = out['raw_records']
            in_arena = len(rr) == 0 or \
                arena.ctypes.data <= rr.ctypes.data < arena.ctypes.data + arena.nbytes
            if not in_arena:                                     # the estimate was too small: the call used
                per_instr = max(per_instr or 0.0, len(rr) / max(len(idx), 1))        # a buffer of its own
                room_for(len(idx))
                arena[used:used + len(rr)] = rr
            else:
                per_instr = max(0.7 * (per_instr or 0.0), len(rr) / max(len(idx), 1))
                for k in keys[1:]:          # they sit behind the TPC records in the arena: move them out
                    out[k] = np.array(out[k])
            used += len(rr)
            for k in keys[1:]:
                side[k].append(np.asarray(out[k]))
            held_truth = out['truth'] if held_truth is None or not len(held_truth) \
                else np.concatenate([held_truth, out['truth']])
            # records per digitisation group (all data types), for the record-buffer rule of the chunk clock
            g = out['groups']
            n_rec_group = np.zeros(len(g), np.int64)
            if len(g):
                lo, hi = g['left'] * cfg['sample_duration'], (g['right'] + 1) * cfg['sample_duration']
                for k in keys:
                    t_k = np.asarray(out[k]['time']) if k != 'raw_records' else rr['time']
                    if len(t_k):
                        if len(t_k) > 1 and (np.diff(t_k) < 0).any():
                            t_k = np.sort(t_k)
                        n_rec_group += np.searchsorted(t_k, hi, side='left') - np.searchsorted(t_k, lo, side='left')
            done = clock.feed(g, n_rec_group)
            if i_part + 1 < len(parts) and n_groups:
                # nothing of the next piece can start before its first signal time minus the longest
                # backward reach of a pulse (left margin, trigger window, diffusion of the drift)
                done += clock.peek(int(next_start[i_part + 1]) - 200000)
            for pre, ct in done:
                self.chunk_time_pre, self.chunk_time = pre, ct
                yield deliver(cut(ct, last=False))
                delivered = True
        self.chunk_time_pre, self.chunk_time = clock.finish()
        self._finished = True
        yield deliver(cut(self.chunk_time, last=True))

    def source_finished(self):
        return self._finished

    @property
    def _n_channels(self):
        return len(self.config.get('gains', np.arange(wcfg.N_TPC_PMTS)))


# --------------------------------------------------------------------------------------------------
# plugin layer
# --------------------------------------------------------------------------------------------------
_OPTION_DEFAULTS = dict(
    detector='XENONnT', event_rate=1000, chunk_size=100, n_chunk=10, per_pmt_truth=False,
    fax_file=None, fax_config='fax_config_nt_design.json', fax_config_override=None,
    fax_config_override_from_cmt=None, channel_map=None, n_tpc_pmts=wcfg.N_TPC_PMTS,
    n_top_pmts=wcfg.N_TOP_PMTS, right_raw_extension=100000, seed=False)

if HAVE_STRAX:                                                          # pragma: no cover
    _PluginBase = strax.Plugin

    def _takes_config(cls):
        from immutabledict import immutabledict
        opts = [strax.Option(k, default=v, track=k in ('detector', 'per_pmt_truth'), infer_type=False)
                for k, v in _OPTION_DEFAULTS.items() if k not in ('channel_map', 'n_tpc_pmts', 'n_top_pmts')]
        opts += [strax.Option('channel_map', track=False, type=immutabledict),
                 strax.Option('n_tpc_pmts', track=False, infer_type=False),
                 strax.Option('n_top_pmts', track=False, infer_type=False)]
        return strax.takes_config(*opts)(cls)
else:
    class _PluginBase:
        """Tiny stand-in for strax.Plugin: a config dict and `chunk()` returning a plain dict."""
        provides = tuple()

        def __init__(self, config=None, run_id='0'):
            self.config = dict(_OPTION_DEFAULTS)
            self.config.update(config or {})
            self.run_id = run_id

        def chunk(self, *, start, end, data, data_type=None):
            return dict(start=start, end=end, data=data, data_type=data_type)

    def _takes_config(cls):
        return cls


@_takes_config
class SimulatorPlugin(_PluginBase):
    compressor = 'zstd'
    depends_on = tuple()
    rechunk_on_save = False     # Cannot arbitrarily rechunk records inside events
    parallel = False            # the simulator is a stateful generator (strax_interface.py:544-546)
    last_chunk_time = -999999999999999
    input_timeout = 3600
    gain_model_mc = None        # to_pe per channel; straxen.URLConfig in a straxen deployment

    def setup(self):
        self.set_config()
        self.get_instructions()
        self.check_instructions()
        self._setup()

    def set_config(self):
        fax = self.config['fax_config']
        if isinstance(fax, str):
            fax = straxen.get_resource(fax, fmt='json') if HAVE_STRAX else wcfg.load_fax_config(fax)
        to_pe = self.config.get('gain_model_mc', self.gain_model_mc)
        if to_pe is None or isinstance(to_pe, str):
            raise ValueError('gain_model_mc must resolve to the to_pe array of the TPC PMTs')
        self.to_pe = np.asarray(to_pe, dtype=np.float64)
        opts = {k: self.config[k] for k in _OPTION_DEFAULTS
                if k in self.config and k not in ('channel_map', 'fax_config', 'fax_config_override')}
        if self.config.get('channel_map'):
            opts['channel_map'] = dict(self.config['channel_map'])
        merged = wcfg.plugin_config(fax, overrides=self.config.get('fax_config_override'), to_pe=self.to_pe,
                                    **opts)
        for k, v in self.config.items():
            merged.setdefault(k, v)
        self.config = merged
        if self.config['seed']:
            np.random.seed(self.config['seed'])
        if self.config.get('fax_config_override_from_cmt') is not None:
            raise NotImplementedError('CMT overrides need straxen.get_correction_from_cmt')

    def _setup(self):
        pass

    def get_instructions(self):
        pass

    def check_instructions(self):
        pass

    def _sort_check(self, results):
        if not isinstance(results, list):
            results = [results]
        last_chunk_time = self.last_chunk_time
        for result in results:
            if len(result) == 0:
                continue
            if result['time'][0] < self.last_chunk_time + 1000:
                raise RuntimeError(
                    "Simulator returned chunks with insufficient spacing. "
                    f"Last chunk's max time was {self.last_chunk_time}, "
                    f"this chunk's first time is {result['time'][0]}.")
            if len(result) == 1:
                continue
            if np.diff(result['time']).min() < 0:
                raise RuntimeError("Simulator returned non-sorted records!")
            last_chunk_time = max(result['time'].max(), self.last_chunk_time)
        self.last_chunk_time = last_chunk_time

    def is_ready(self, chunk_i):
        """Flip-flop: False to check source finished, True to get the next chunk."""
        if 'ready' not in self.__dict__:
            self.ready = False
        self.ready ^= True
        return self.ready

    def source_finished(self):
        return self.sim.source_finished()

    @property
    def _n_channels(self):
        return len(self.config.get('gains', np.arange(wcfg.N_TPC_PMTS)))

    @property
    def _truth_dtype(self):
        truth_per_n_pmts = self._n_channels if self.config.get('per_pmt_truth') else False
        return extra_truth_dtype_per_pmt(truth_per_n_pmts)


class RawRecordsFromFaxNT(SimulatorPlugin):
    provides = ('raw_records', 'raw_records_he', 'raw_records_aqmon', 'truth')
    data_kind = dict(zip(provides, provides))
    resource_overrides = None    # extra Resource pieces (spe table, noise, afterpulse tables)
    device = 0

    def _setup(self):
        self.sim = ChunkRawRecords(self.config, device=self.device, **(self.resource_overrides or {}))
        self.sim_iter = self.sim(self.instructions)

    def get_instructions(self):
        if self.config['fax_file']:
            assert not self.config['fax_file'].endswith('root'), \
                'None optical g4 input is deprecated use EPIX instead'
            assert self.config['fax_file'].endswith('csv'), 'Only csv input is supported'
            self.instructions = instruction_from_csv(self.config['fax_file'])
        elif getattr(self, 'instructions', None) is None:
            # strax_interface.py:680; nestpy yields when nestpy is installed, else the fixed-yield stand-in
            from .instructions import rand_instructions
            self.instructions = rand_instructions(self.config, seed=self.config.get('seed') or None)

    def check_instructions(self):
        # Let below cathode S1 instructions pass but remove S2 instructions
        m = (self.instructions['z'] < - self.config['tpc_length']) & (self.instructions['type'] == 2)
        self.instructions = self.instructions[~m]
        r_instr = np.sqrt(self.instructions['x']**2 + self.instructions['y']**2)
        assert np.all((r_instr < self.config['tpc_radius']) | np.isclose(r_instr, self.config['tpc_radius'])), \
            "Interaction is outside the TPC (radius)"
        assert np.all(self.instructions['z'] < 0.25), "Interaction is outside the TPC (in Z)"
        assert np.all(self.instructions['amp'] > 0), "Interaction has zero size"

    def infer_dtype(self):
        dtype = {data_type: raw_record_dtype(samples_per_record=RECORD_LENGTH)
                 for data_type in self.provides if data_type != 'truth'}
        dtype['truth'] = instruction_dtype + self._truth_dtype
        return dtype

    def compute(self):
        try:
            result = next(self.sim_iter)
        except StopIteration:
            raise RuntimeError("Bug in chunk count computation")
        self._sort_check(result[self.provides[0]])
        return {data_type: self.chunk(
            start=self.sim.chunk_time_pre,
            end=self.sim.chunk_time,
            data=result[data_type],
            data_type=data_type) for data_type in self.provides}


class RawRecordsFromFaxOpticalNT(RawRecordsFromFaxNT):
    """Mirror of strax_interface.py:722-751: photons (channel, arrival time) come from outside -- G4 optical
    output, neutron-veto style inputs -- and only the PMT stage (transit time, double photo-electrons, SPE
    gains, afterpulses), the digitiser and the ZLE are simulated.  Set `instructions` (instruction_dtype +
    optical_extra_dtype, `_first` / `_last` indexing the lists), `channels` and `timings` before setup();
    reading the G4 files themselves (`read_optical`, uproot) is not part of this package."""

    def _setup(self):
        self.sim = ChunkRawRecords(self.config, device=self.device, channels=self.channels, timings=self.timings,
                                   **(self.resource_overrides or {}))
        self.sim_iter = self.sim(self.instructions)

    def get_instructions(self):
        if getattr(self, 'instructions', None) is None or getattr(self, 'channels', None) is None \
                or getattr(self, 'timings', None) is None:
            raise RuntimeError('optical input: set instructions (with _first/_last), channels and timings; '
                               'reading G4 optical files needs uproot (third party)')

    def infer_dtype(self):
        dtype = super().infer_dtype()
        dtype['truth'] = instruction_dtype + optical_extra_dtype + self._truth_dtype
        return dtype
