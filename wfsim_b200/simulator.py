"""Python face of one per-GPU simulator handle (ctypes -> libwfsim_b200.so)."""
import ctypes as C

import numpy as np

from . import lib as wlib
from . import params as wparams
from .dtypes import raw_record_dtype, instruction_dtype, truth_dtype
from .resource import Resource, evaluate_instruction_maps


class SimulatorError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data if a is not None else None


def host_records(n, dtype=None):
    """Uninitialised host array for `n` records.  (Asking for transparent huge pages here was tried:
    the GPU boxes of this pool do not grant them and every page fault of a madvise(MADV_HUGEPAGE) range
    then also pays for a compaction attempt -- 0.47 s instead of 0.34 s for a 1.9 GB output.)"""
    return np.empty(int(n), np.dtype(raw_record_dtype() if dtype is None else dtype))


def pin_array(lib, array):
    """Page-locks the memory of a numpy array the caller owns (wfs_host_register) until the array is
    garbage-collected.  Into such a destination the library moves part of the records by plain DMA.  Returns
    False (and leaves the array usable as ordinary memory) when the driver refuses."""
    import weakref
    if array.nbytes == 0 or not array.flags['C_CONTIGUOUS'] or not array.flags['OWNDATA']:
        return False
    ptr = array.ctypes.data
    if lib.wfs_host_register(ptr, array.nbytes) != 0:
        return False
    weakref.finalize(array, lib.wfs_host_unregister, ptr)
    return True


class PinnedArray:
    """numpy view over pinned host memory obtained from the library (plain DMA target)."""

    def __init__(self, lib, n, dtype):
        self._lib = lib
        dtype = np.dtype(dtype)
        self.nbytes = max(int(n) * dtype.itemsize, 1)
        self.ptr = lib.wfs_host_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(f'cannot pin {self.nbytes} bytes')
        buf = (C.c_uint8 * self.nbytes).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(n))

    def free(self):
        if self.ptr:
            self.array = None
            self._lib.wfs_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Simulator:
    """One handle per GPU.  `config` is the dict the reference simulator classes receive
    (fax_config merged with the plugin-derived keys, see wfsim_b200.config.plugin_config)."""

    def __init__(self, config, resource=None, device=0):
        self.lib = wlib.load()
        if self.lib.wfs_device_count() <= 0:
            raise SimulatorError('no CUDA device visible: wfsim_b200 has no CPU fallback')
        self.config = config
        self.resource = resource
        self.params = wparams.build_params(config)
        if resource is not None:
            wparams.set_resource_scalars(self.params, config, resource)
        self.tables = wparams.build_tables(config, resource)
        h = C.c_void_p()
        rc = self.lib.wfs_create(C.byref(self.params), C.byref(self.tables.struct), device, C.byref(h))
        if rc != 0:
            raise SimulatorError(f'wfs_create failed ({rc}): {self.lib.wfs_last_error(None).decode()}')
        self.handle = h
        self.device = device
        self.last_counts = None
        self._pinned_cache = None      # reusable pinned record buffer (pinning is slow)

    def pin(self, array):
        """Page-lock a caller-owned record array (see pin_array)."""
        return pin_array(self.lib, array)

    def close(self):
        if getattr(self, '_pinned_cache', None) is not None:
            self._pinned_cache.free()
            self._pinned_cache = None
        if getattr(self, 'handle', None):
            self.lib.wfs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def quiet_gap(self):
        """Smallest signal-time gap [ns] at which a run may be cut into independent calls (pieces,
        shards) -- the bound the library itself cuts its device batches at (wfs_quiet_gap)."""
        return int(self.lib.wfs_quiet_gap(self.handle))

    def _raise(self, rc):
        msg = self.lib.wfs_last_error(self.handle).decode()
        if rc == wlib.E_PULSE_CACHE_TOO_LONG:
            raise AssertionError('Pulse cache too long')      # rawdata.py:219
        raise SimulatorError(f'libwfsim_b200 error {rc}: {msg}')

    def _split(self, rec, counts):
        n0, n1, n2 = counts.n_records
        return dict(raw_records=rec[:n0], raw_records_he=rec[n0:n0 + n1],
                    raw_records_aqmon=rec[n0 + n1:n0 + n1 + n2])

    def simulate_photons(self, t_ns, channel, gain, pulse_call, group_of, ix_rand=None, seed=0,
                         cap_records=None, pinned=False):
        """Deterministic entry (see wfs_simulate_photons in the header).  Returns a dict with
        raw_records / raw_records_he / raw_records_aqmon (strax.raw_record_dtype arrays, each
        sorted by (time, channel)) and `groups` (left, right, n_intervals per group)."""
        t_ns = np.ascontiguousarray(t_ns, np.int64)
        channel = np.ascontiguousarray(channel, np.int32)
        gain = np.ascontiguousarray(gain, np.float64)
        pulse_call = np.ascontiguousarray(pulse_call, np.int32)
        group_of = np.ascontiguousarray(group_of, np.int32)
        n, n_pc = len(t_ns), len(group_of)
        if not (len(channel) == len(gain) == len(pulse_call) == n):
            raise ValueError('photon arrays must have equal length')
        n_groups = int(group_of.max()) + 1 if n_pc else 0
        if ix_rand is not None:
            ix_rand = np.ascontiguousarray(ix_rand, np.int64)
            if len(ix_rand) != n_groups:
                raise ValueError('ix_rand needs one entry per group')
        groups = np.zeros(n_groups, dtype=[('left', np.int64), ('right', np.int64), ('n_intervals', np.int64)])
        counts = wlib.Counts()
        cap = int(cap_records) if cap_records is not None else max(1024, n // 2)
        while True:
            holder = PinnedArray(self.lib, cap, raw_record_dtype()) if pinned else None
            rec = holder.array if pinned else np.zeros(cap, raw_record_dtype())
            rc = self.lib.wfs_simulate_photons(
                self.handle, n, _ptr(t_ns), _ptr(channel), _ptr(gain), _ptr(pulse_call), n_pc,
                _ptr(group_of), n_groups, _ptr(ix_rand), int(seed), 0, _ptr(rec), cap,
                _ptr(groups) if n_groups else None, C.byref(counts))
            if rc == wlib.E_CAPACITY:
                cap = int(counts.need_records)
                if holder is not None:
                    holder.free()
                continue
            if rc != 0:
                self._raise(rc)
            break
        self.last_counts = counts.as_dict()
        out = self._split(rec[:counts.n_records_total], counts)
        out['groups'] = groups
        out['_pinned'] = holder
        return out

    # -----------------------------------------------------------------------------------------
    def _maps_struct(self, instructions, maps=None, rng_id=None, seed=0, optical=None):
        if maps is None:
            if self.resource is None or isinstance(self.resource, dict):
                raise SimulatorError('simulate() needs a Resource (maps) -- pass resource= to Simulator')
            maps = evaluate_instruction_maps(self.config, self.resource, instructions, seed=seed, rng_id=rng_id)
        keep = {k: np.ascontiguousarray(v, dtype=(np.float32 if k == 'pattern' else
                                                   np.int32 if k in ('pattern_row', 'gg_lo_row', 'gg_hi_row')
                                                   else np.float64))
                for k, v in maps.items()}
        m = wlib.InstrMaps()
        m.s1_lce = _ptr(keep['s1_lce'])
        m.s2_sc_gain = _ptr(keep['s2_sc_gain'])
        m.s2_cy_extra = _ptr(keep['s2_cy_extra'])
        m.pattern = _ptr(keep['pattern'])
        m.pattern_row = _ptr(keep['pattern_row'])
        m.n_pattern_rows = keep['pattern'].shape[0]
        m.s2_sc_gain_default = 0.0
        if optical is not None:
            # externally supplied photons (RawDataOptical, rawdata.py:461-495): optical = (first, last,
            # channels, timings); `first` / `last` are the `_first` / `_last` columns of the instructions
            first, last, channels, timings = optical
            keep['opt_first'] = np.ascontiguousarray(first, np.int64)
            keep['opt_last'] = np.ascontiguousarray(last, np.int64)
            keep['opt_channels'] = np.ascontiguousarray(channels, np.int32)
            keep['opt_timings'] = np.ascontiguousarray(timings, np.int64)
            if len(keep['opt_first']) != len(instructions) or len(keep['opt_last']) != len(instructions):
                raise ValueError('optical first / last need one entry per instruction')
            if len(keep['opt_channels']) != len(keep['opt_timings']):
                raise ValueError('optical channels and timings must have equal length')
            m.opt_first, m.opt_last = _ptr(keep['opt_first']), _ptr(keep['opt_last'])
            m.opt_channels, m.opt_timings = _ptr(keep['opt_channels']), _ptr(keep['opt_timings'])
            m.n_opt = len(keep['opt_channels'])
            m.opt_time_cutoff = int(self.config.get('nveto_time_max_cutoff', int(1e6)))
        for k in ('drift_velocity', 'diffusion_long', 'x_obs', 'y_obs', 'gg_lo_row', 'gg_hi_row', 'gg_frac',
                  'hdiff_sigma_r', 'hdiff_sigma_a', 'lum_gap', 'lum_e0'):
            if k in keep:
                setattr(m, k, _ptr(keep[k]))
        if rng_id is not None:
            keep['rng_id'] = np.ascontiguousarray(rng_id, np.uint64)
            if len(keep['rng_id']) != len(instructions):
                raise ValueError('rng_id needs one entry per instruction')
            m.rng_id = _ptr(keep['rng_id'])
        return m, keep

    def _pinned_records(self, cap):
        """Pinned host buffer for `cap` records, reused across calls (grow-only).  The arrays
        returned by simulate(pinned=True) are views into it and are overwritten by the next call."""
        c = self._pinned_cache
        if c is None or len(c.array) < cap:
            if c is not None:
                c.free()
            self._pinned_cache = c = PinnedArray(self.lib, int(cap * 1.05) + 1024, raw_record_dtype())
        return c

    def simulate(self, instructions, seed=0, maps=None, cap_records=None, pinned=False, rng_id=None,
                 per_pmt_truth=None, records_out=None, optical=None):
        """Full path for one set of instructions (see wfs_simulate in the header).

        Returns dict(raw_records, raw_records_he, raw_records_aqmon, truth, groups); records of
        each data type are sorted by (time, channel); truth rows are in Pulse-call execution
        order with `time` still the instruction time (the chunker sets it to t_first_photon,
        strax_interface.py:481-482).  With per_pmt_truth (default: config['per_pmt_truth']) the
        truth rows carry the `*_per_pmt` fields of extra_truth_dtype_per_pmt instead of `*_bottom`.
        `records_out`: a caller-owned raw_record array the records are written into (any host
        memory; the returned record arrays are views of it) -- grown if it turns out too small."""
        if per_pmt_truth is None:
            per_pmt_truth = bool(self.config.get('per_pmt_truth', False))
        n_pmt = int(self.params.n_tpc_pmts)
        instructions, optical = self._split_optical(instructions, optical)
        if instructions.dtype.itemsize != 70:
            raise ValueError('instructions must have the packed 70-byte instruction_dtype')
        n = len(instructions)
        m, keep = self._maps_struct(instructions, maps, rng_id, seed=seed, optical=optical)
        counts = wlib.Counts()
        cap_rec = int(cap_records) if cap_records is not None else \
            (len(records_out) if records_out is not None else max(4096, 1500 * n))
        cap_truth, cap_groups, cap_batches = 2 * n + 64, n + 64, 4096
        tdt = truth_dtype()
        gdt = np.dtype([('left', np.int64), ('right', np.int64), ('n_intervals', np.int64)])
        while True:
            holder = self._pinned_records(cap_rec) if pinned else None
            if records_out is not None and len(records_out) >= cap_rec:
                if records_out.dtype != raw_record_dtype() or not records_out.flags['C_CONTIGUOUS']:
                    raise ValueError('records_out must be a C-contiguous raw_record_dtype array')
                rec = records_out
            else:
                rec = holder.array if pinned else host_records(cap_rec)
            cap_rec = len(rec)
            truth = np.zeros(cap_truth, tdt)
            groups = np.zeros(cap_groups, gdt)
            batch_records = np.zeros((cap_batches, 3), np.int64)
            pmt_counts = np.zeros((cap_truth, 4, n_pmt), np.int32) if per_pmt_truth else None
            pmt_areas = np.zeros((cap_truth, 2, n_pmt), np.float64) if per_pmt_truth else None
            out = wlib.Outputs(_ptr(rec), cap_rec, _ptr(truth), cap_truth, _ptr(groups), cap_groups,
                               _ptr(batch_records), cap_batches, _ptr(pmt_counts), _ptr(pmt_areas))
            rc = self.lib.wfs_simulate(self.handle, _ptr(instructions.view(np.uint8)), n, C.byref(m),
                                       int(seed), C.byref(out), C.byref(counts))
            if rc == wlib.E_CAPACITY:
                cap_rec = max(cap_rec, int(counts.need_records))
                cap_truth = max(cap_truth, int(counts.need_truth))
                cap_groups = max(cap_groups, int(counts.need_groups))
                cap_batches = max(cap_batches, int(counts.need_batches))
                continue
            if rc != 0:
                self._raise(rc)
            break
        self.last_counts = counts.as_dict()
        nb = counts.n_batches
        br = batch_records[:nb]
        total = int(counts.n_records_total)
        rec = rec[:total]
        if nb <= 1 or (br[:, 1:].sum() == 0):
            n0 = int(br[:, 0].sum())
            res = dict(raw_records=rec[:n0], raw_records_he=rec[n0:n0 + int(br[:, 1].sum())],
                       raw_records_aqmon=rec[n0 + int(br[:, 1].sum()):total])
        else:
            offs = np.concatenate([[0], np.cumsum(br.sum(axis=1))])
            parts = [[], [], []]
            for b in range(nb):
                o = offs[b]
                for k in range(3):
                    parts[k].append(rec[o:o + br[b, k]])
                    o += br[b, k]
            res = dict(zip(('raw_records', 'raw_records_he', 'raw_records_aqmon'),
                           (np.concatenate(p) for p in parts)))
        # Device batches are cut at quiet gaps, so their records follow each other in time.  Only when the
        # stream has no quiet gap within twice the batch budget (far beyond physical TPC rates) is a cut
        # forced, and delayed secondaries of one batch may then lie behind the start of the next: restore
        # the (time, channel) order the consumer relies on (strax_interface.py:453, 622-640).
        if nb > 1:
            starts = np.cumsum(br, axis=0)
            for k, name in enumerate(('raw_records', 'raw_records_he', 'raw_records_aqmon')):
                r = res[name]
                cuts = starts[:-1, k]
                cuts = cuts[(cuts > 0) & (cuts < len(r))]
                if len(cuts) and (r['time'][cuts] < r['time'][cuts - 1]).any():
                    order = np.lexsort((r['channel'], r['time']))
                    res[name] = r[order]
        truth = truth[:counts.n_truth]
        if per_pmt_truth:
            full = np.zeros(len(truth), truth_dtype(n_pmt))
            for name in full.dtype.names:
                if name in truth.dtype.names:
                    full[name] = truth[name]
            for k, f in enumerate(('n_photon', 'n_pe', 'n_photon_trigger', 'n_pe_trigger')):
                full[f + '_per_pmt'] = pmt_counts[:len(truth), k]
            for k, f in enumerate(('raw_area', 'raw_area_trigger')):
                full[f + '_per_pmt'] = pmt_areas[:len(truth), k]
            truth = full
        res['truth'] = truth
        groups = groups[:counts.n_groups]
        # n_intervals < 0 marks a group index without pulses (trailing Pulse calls that made nothing, at the
        # end of a device batch): the reference digitises nothing for it (rawdata.py:207-208)
        res['groups'] = groups[groups['n_intervals'] >= 0]
        res['_pinned'] = None
        return res

    @staticmethod
    def _split_optical(instructions, optical):
        """Instructions with the optical extra columns (`_first`, `_last`, strax_interface.py:44-45) plus
        optical=(channels, timings) -> plain 70-byte instructions + (first, last, channels, timings)."""
        instructions = np.ascontiguousarray(instructions)
        if optical is not None and len(optical) == 2:
            if '_first' not in (instructions.dtype.names or ()):
                raise ValueError("optical photons need instructions with the '_first' / '_last' columns")
            optical = (instructions['_first'], instructions['_last'], optical[0], optical[1])
        if instructions.dtype.names and '_first' in instructions.dtype.names:
            plain = np.zeros(len(instructions), instruction_dtype)
            for name in plain.dtype.names:
                plain[name] = instructions[name]
            instructions = plain
        return instructions, optical

    def stage(self, instructions, maps=None):
        instructions = np.ascontiguousarray(instructions)
        m, keep = self._maps_struct(instructions, maps)
        rc = self.lib.wfs_stage_instructions(self.handle, _ptr(instructions.view(np.uint8)),
                                             len(instructions), C.byref(m))
        if rc != 0:
            self._raise(rc)

    def run_staged(self, seed=0):
        counts = wlib.Counts()
        rc = self.lib.wfs_run_staged(self.handle, int(seed), None, C.byref(counts))
        if rc != 0:
            self._raise(rc)
        self.last_counts = counts.as_dict()
        return self.last_counts

    PHOTON_DUMP_DTYPE = np.dtype([('t', np.int64), ('gain', np.float64), ('channel', np.int32),
                                  ('instruction', np.int32), ('flags', np.int32), ('secondary', np.int32)])

    SECONDARY_DUMP_DTYPE = np.dtype([('time', np.int64), ('x', np.float32), ('y', np.float32), ('amp', np.int32),
                                     ('parent', np.int32), ('type', np.int32), ('z', np.float32)])

    def sample_secondaries(self, instructions, seed=0, maps=None):
        """The secondary (type 4 / 6) instructions the S2 calls spawn (stage 4 of wfs_sample_stage): row k
        is the secondary the photon / electron dumps call `secondary == k`."""
        return self.sample_stage(instructions, stage=4, seed=seed, maps=maps).view(self.SECONDARY_DUMP_DTYPE)

    def sample_stage(self, instructions, stage=0, seed=0, maps=None, optical=None):
        """Front-end only: photons (stage 0) or emitters/electrons (stage 1) as a structured array
        (see wfs_sample_stage).  For the statistical parity tests."""
        instructions, optical = self._split_optical(instructions, optical)
        m, keep = self._maps_struct(instructions, maps, seed=seed, optical=optical)
        n_out = C.c_int64()
        cap = 1 << 16
        while True:
            out = np.zeros(cap, self.PHOTON_DUMP_DTYPE)
            rc = self.lib.wfs_sample_stage(self.handle, int(stage), _ptr(instructions.view(np.uint8)),
                                           len(instructions), C.byref(m), int(seed), _ptr(out), cap,
                                           C.byref(n_out))
            if rc == wlib.E_CAPACITY:
                cap = int(n_out.value)
                continue
            if rc != 0:
                self._raise(rc)
            return out[:n_out.value]
