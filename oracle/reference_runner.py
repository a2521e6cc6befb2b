"""TEST / BENCHMARK INFRASTRUCTURE -- times the UNMODIFIED reference on the CPU.

`run_events` drives the reference's own ChunkRawRecords (strax_interface.py:353-504: RawData.__call__ ->
S1 / S2 / Pulse / digitize_pulse_cache / ZLE -> record packing) over a set of instructions, through the
stand-ins of oracle/ref_loader.py for strax / straxen (absent from this image), exactly as
tests/golden/make_golden*.py run it.  Used by `bench.py --impl reference` and `cpu_baseline`
("what is timed: the reference implementation itself").  Nothing under wfsim_b200/
imports this."""
import os
import time

import numpy as np

from . import ref_loader as RL


def _spe_hook(path, fmt):
    # the single-channel SPE csv widened to 494 identical columns (tests/test_wfsim.py:83-88)
    import pandas as pd
    if fmt == 'csv' and 'spe_distributions' in path:
        df = pd.read_csv(os.path.join(RL.REFERENCE_ROOT, 'files', 'XENONnT_spe_distributions_single_channel.csv'))
        cols = {str(i): df['0'] for i in range(1, 494)}
        return pd.concat([df, pd.DataFrame(cols)], axis=1)
    return None


_state = {}


def chunker(cfg):
    """One reference ChunkRawRecords per process (tables and numba compilations are reused)."""
    if 'crr' not in _state:
        ref = RL.load_reference(resource_hook=_spe_hook)
        _state['ref'] = ref
        _state['crr'] = ref.ChunkRawRecords(dict(cfg))
    return _state['ref'], _state['crr']


def run_events(cfg, instructions, seed=0):
    """-> dict(seconds, n_pe, n_records, n_truth).  `instructions`: packed instruction_dtype rows."""
    ref, crr = chunker(cfg)
    RL.seed_reference_rngs(seed)
    inst = np.zeros(len(instructions), dtype=ref.strax_interface.instruction_dtype)
    for n in inst.dtype.names:
        inst[n] = instructions[n]
    # a fresh chunker state for every call (the class keeps its buffers and chunk clock between calls)
    crr = ref.ChunkRawRecords(dict(cfg))
    n_pe = n_rec = n_truth = 0
    t0 = time.perf_counter()
    for res in crr(inst):
        n_rec += sum(len(res[k]) for k in res if k.startswith('raw_records'))
        n_pe += int(res['truth']['n_pe'].sum())
        n_truth += len(res['truth'])
    return dict(seconds=time.perf_counter() - t0, n_pe=n_pe, n_records=n_rec, n_truth=n_truth)
