"""Python face of one per-GPU simulator handle (ctypes -> libwfsim_b200.so)."""
import ctypes as C

import numpy as np

from . import lib as wlib
from . import params as wparams
from .dtypes import raw_record_dtype


class SimulatorError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data if a is not None else None


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
        self.params = wparams.build_params(config)
        self.tables = wparams.build_tables(config, resource)
        h = C.c_void_p()
        rc = self.lib.wfs_create(C.byref(self.params), C.byref(self.tables.struct), device, C.byref(h))
        if rc != 0:
            raise SimulatorError(f'wfs_create failed ({rc}): {self.lib.wfs_last_error(None).decode()}')
        self.handle = h
        self.device = device
        self.last_counts = None

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.wfs_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

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
