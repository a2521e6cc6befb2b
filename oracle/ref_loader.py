"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* WFSim reference for golden-vector generation.

Nothing in the product (`wfsim_b200/`) may import this module.  It is used by
`tests/golden/make_golden.py` (run in the build container, where the reference checkout is
mounted read-only at /root/reference) and by the CPU-side tests that cross-check the
restatement in `oracle/wfsim_oracle.py` against the reference itself.  On the GPU box the
reference checkout does not exist; `available()` returns False there and callers skip.

The reference package cannot be imported normally in this image because strax, straxen,
nestpy, immutabledict and uproot are absent.  The hot-path modules only
use a handful of names from strax/straxen, so we register two tiny stand-in modules and load
the eight hot-path files by path under a synthetic `wfsim` package, bypassing
`wfsim/__init__.py`.  No reference source is copied: the files are executed where they lie.
"""
import importlib.util
import json
import os
import re
import sys
import types

import numpy as np

def _reference_root():
    # the checkout in the build container; on the GPU box the modules staged by oracle/make_ref.sh (git-ignored)
    env = os.environ.get('WFSIM_REFERENCE_ROOT')
    if env:
        return env
    if os.path.isfile('/root/reference/wfsim/core/pulse.py'):
        return '/root/reference'
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


REFERENCE_ROOT = _reference_root()

N_TPC_PMTS = 494      # straxen.n_tpc_pmts
N_TOP_PMTS = 253      # straxen.n_top_pmts
RECORD_LENGTH = 110   # strax.DEFAULT_RECORD_LENGTH


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'wfsim', 'core', 'pulse.py'))


def raw_record_dtype(samples_per_record=RECORD_LENGTH):
    """Field layout of strax.raw_record_dtype (strax is a third-party dependency absent from
    /root/reference; this is its published 244-byte layout)."""
    return np.dtype([
        (('Start time since unix epoch [ns]', 'time'), np.int64),
        (('Length of the interval in samples', 'length'), np.int32),
        (('Width of one sample [ns]', 'dt'), np.int16),
        (('Channel/PMT number', 'channel'), np.int16),
        (('Length of pulse to which the record belongs (without zero-padding)', 'pulse_length'), np.int32),
        (('Fragment number in the pulse', 'record_i'), np.int16),
        (('Baseline determined by the digitizer (if this is supported)', 'baseline'), np.int16),
        (('Waveform data in raw ADC counts', 'data'), np.int16, samples_per_record)])


def _sort_by_time(x):
    if len(x) == 0:
        return x
    if 'channel' in x.dtype.names:
        key = (x['time'] - x['time'].min()) * (int(x['channel'].max()) + 1) + x['channel']
        return x[np.argsort(key, kind='mergesort')]
    return x[np.argsort(x['time'], kind='mergesort')]


def _exporter(export_self=False):
    all_ = []
    if export_self:
        all_.append('exporter')

    def decorator(obj):
        all_.append(obj.__name__)
        return obj
    return decorator, all_


def _deterministic_hash(thing):
    import hashlib
    import base64

    def norm(o):
        if isinstance(o, dict):
            return {str(k): norm(o[k]) for k in sorted(o, key=str)}
        if isinstance(o, (list, tuple)):
            return [norm(v) for v in o]
        if isinstance(o, np.ndarray):
            return ['ndarray', o.shape, hashlib.sha1(np.ascontiguousarray(o).tobytes()).hexdigest()]
        if isinstance(o, (np.integer,)):
            return int(o)
        if isinstance(o, (np.floating,)):
            return float(o)
        if isinstance(o, (str, int, float, bool, type(None))):
            return o
        return repr(o)
    digest = hashlib.sha1(json.dumps(norm(thing), sort_keys=True, default=repr).encode()).digest()
    return base64.b32encode(digest)[:10].decode().lower()


def _install_stubs(resource_hook=None):
    import pandas as pd

    strax = types.ModuleType('strax')
    strax.exporter = _exporter
    strax.deterministic_hash = _deterministic_hash
    strax.raw_record_dtype = raw_record_dtype
    strax.DEFAULT_RECORD_LENGTH = RECORD_LENGTH
    strax.sort_by_time = _sort_by_time
    strax_utils = types.ModuleType('strax.utils')

    class _NoBar:
        def __init__(self, iterable=None, *a, **k):
            self._it = iterable

        def __iter__(self):
            return iter(self._it)

        def update(self, *a):
            pass

        def close(self):
            pass
    strax_utils.tqdm = _NoBar
    strax.utils = strax_utils

    straxen = types.ModuleType('straxen')
    straxen.n_tpc_pmts = N_TPC_PMTS
    straxen.n_top_pmts = N_TOP_PMTS
    straxen.tpc_r = 66.4
    straxen.tpc_z = 148.6515

    def get_resource(path, fmt='text'):
        if resource_hook is not None:
            out = resource_hook(path, fmt)
            if out is not None:
                return out
        if fmt == 'csv':
            local = path
            if not os.path.isfile(local):
                local = os.path.join(REFERENCE_ROOT, 'files', os.path.basename(path))
            return pd.read_csv(local)
        raise FileNotFoundError(f'stub straxen.get_resource cannot serve {path} ({fmt})')
    straxen.get_resource = get_resource

    def _no_mongo(*a, **k):
        raise NameError('no utilix in the stub')
    straxen.MongoDownloader = _no_mongo
    sys.modules['strax'] = strax
    sys.modules['strax.utils'] = strax_utils
    sys.modules['straxen'] = straxen
    return strax, straxen


_loaded = {}


def load_reference(resource_hook=None):
    """Return a namespace with the reference's hot-path modules:
    .units .utils .load_resource .pulse .s1 .s2 .afterpulse .rawdata (+ classes)."""
    if 'ns' in _loaded and resource_hook is None:
        return _loaded['ns']
    if not available():
        raise RuntimeError(f'reference checkout not found at {REFERENCE_ROOT}')
    _install_stubs(resource_hook)
    pkg = types.ModuleType('wfsim')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'wfsim')]
    sys.modules['wfsim'] = pkg
    core = types.ModuleType('wfsim.core')
    core.__path__ = [os.path.join(REFERENCE_ROOT, 'wfsim', 'core')]
    sys.modules['wfsim.core'] = core
    pkg.core = core

    def load(modname, relpath):
        spec = importlib.util.spec_from_file_location(
            modname, os.path.join(REFERENCE_ROOT, relpath))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    ns = types.SimpleNamespace()
    ns.units = load('wfsim.units', 'wfsim/units.py')
    pkg.units = ns.units
    ns.utils = load('wfsim.utils', 'wfsim/utils.py')
    pkg.utils = ns.utils
    ns.load_resource = load('wfsim.load_resource', 'wfsim/load_resource.py')
    pkg.load_resource = ns.load_resource
    ns.pulse = load('wfsim.core.pulse', 'wfsim/core/pulse.py')
    ns.s1 = load('wfsim.core.s1', 'wfsim/core/s1.py')
    ns.s2 = load('wfsim.core.s2', 'wfsim/core/s2.py')
    ns.afterpulse = load('wfsim.core.afterpulse', 'wfsim/core/afterpulse.py')
    ns.rawdata = load('wfsim.core.rawdata', 'wfsim/core/rawdata.py')
    ns.Pulse = ns.pulse.Pulse
    ns.S1 = ns.s1.S1
    ns.S2 = ns.s2.S2
    ns.PMT_Afterpulse = ns.afterpulse.PMT_Afterpulse
    ns.PhotoIonization_Electron = ns.afterpulse.PhotoIonization_Electron
    ns.RawData = ns.rawdata.RawData
    ns.find_intervals_below_threshold = ns.utils.find_intervals_below_threshold
    pkg.RawData = ns.RawData
    pkg.load_config = ns.load_resource.load_config
    ns.strax_interface = _load_strax_interface(load)
    ns.ChunkRawRecords = ns.strax_interface.ChunkRawRecords
    if resource_hook is None:
        _loaded['ns'] = ns
    return ns


def _load_strax_interface(load):
    """wfsim/strax_interface.py imports uproot, immutabledict and a few strax/straxen plugin
    classes at module level; none is used by ChunkRawRecords (:353-504) itself, so empty
    stand-ins are enough to execute the file and obtain the reference record packer."""
    strax = sys.modules['strax']
    straxen = sys.modules['straxen']
    if 'immutabledict' not in sys.modules:
        m = types.ModuleType('immutabledict')
        m.immutabledict = dict
        sys.modules['immutabledict'] = m
    if 'uproot' not in sys.modules:
        sys.modules['uproot'] = types.ModuleType('uproot')

    class Option:
        def __init__(self, name, **kw):
            self.name = name
            self.kw = kw

    def takes_config(*options):
        def deco(cls):
            cls.takes_config = {o.name: o for o in options}
            return cls
        return deco

    class Plugin:
        def chunk(self, *, start, end, data, data_type=None):
            return dict(start=start, end=end, data=data, data_type=data_type)
    strax.Option = Option
    strax.takes_config = takes_config
    strax.Plugin = Plugin
    strax.OverlapWindowPlugin = Plugin
    strax.LoopPlugin = Plugin

    class URLConfig:
        def __init__(self, default=None, **kw):
            self.default = default

        def __set_name__(self, owner, name):
            self.name = name

        def __get__(self, obj, objtype=None):
            if obj is None:
                return self
            return obj.__dict__.get('_urlcfg_' + self.name, self.default)

        def __set__(self, obj, value):
            obj.__dict__['_urlcfg_' + self.name] = value
    straxen.URLConfig = URLConfig
    return load('wfsim.strax_interface', 'wfsim/strax_interface.py')


def seed_reference_rngs(seed):
    """Seed both numpy's global generator and numba's internal generator (the reference
    plugin seeds only the former)."""
    import numba

    @numba.njit
    def _seed(s):
        np.random.seed(s)
    np.random.seed(seed)
    _seed(seed)
