"""CPU tests of the host-side mirror of the reference interface (no GPU needed):
chunk cutting vs golden vectors produced by the reference's ChunkRawRecords, dtypes, config
loader, tables vs the reference's tables, and that the C-ABI library exports every symbol
declared in include/wfsim_b200.h."""
import json
import os
import re

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT, load_c0_config


def test_chunk_boundaries_match_reference():
    from wfsim_b200.strax_interface import chunk_boundaries
    from oracle.wfsim_oracle_sim import chunk_boundaries as oracle_cb
    with open(os.path.join(GOLDEN, 'chunks.json')) as f:
        cases = json.load(f)
    for name, c in cases.items():
        cfg = load_c0_config()
        cfg['chunk_size'] = c['chunk_size']
        groups = [tuple(g) for g in c['groups']]
        n_records = [2 * g[2] for g in groups]          # 120 samples per interval -> 2 records
        got = chunk_boundaries(cfg, c['t_min_instruction'], groups, n_records=n_records, record_buffer=c.get('record_buffer'))
        assert [list(b) for b in got] == c['bounds'], name
        if not c.get('record_buffer'):
            assert oracle_cb(cfg, c['t_min_instruction'], groups) == got
            assert chunk_boundaries(cfg, c['t_min_instruction'], groups) == got      # 5e6 records are never reached here
        # records per chunk: an interval's record belongs to the first chunk whose end >= its time
        times = []
        for left, right, n in groups:
            for k in range(n):      # 120 samples -> 2 records (110 + 10)
                a = left + 60 + 200 * k
                times += [a * 10, (a + 110) * 10]
        times = np.sort(np.array(times))
        done, counts = 0, []
        for pre, ct in got:
            stop = done + int(np.searchsorted(times[done:], ct, side='right'))
            counts.append(stop - done)
            done = stop
        assert counts == c['records_per_chunk'], name


def test_chunk_clock_pieces_with_a_lagging_clock():
    """Random group sequences with chunks far shorter than the spacing of the groups (the reference closes at most one
    chunk per ZLE interval, so its clock lags) and short record buffers: feeding in pieces with peek() gives the
    one-pass bounds, as long as an interval follows the peek; if none ever does (a run that ends with groups without
    intervals) the only difference is one more, empty, chunk at the end."""
    import logging
    from tests.golden.make_golden_chunks import make_groups
    from wfsim_b200.strax_interface import ChunkClock
    logging.disable(logging.WARNING)
    try:
        rng = np.random.default_rng(0)
        cfg = load_c0_config()
        tails = 0
        for trial in range(400):
            cfg['chunk_size'] = float(rng.choice([0.05, 0.5, 2]))
            groups = make_groups(int(rng.integers(1 << 30)), int(rng.integers(3, 40)), float(rng.choice([0.05, 0.4, 1.0])))
            nrec = [2 * g[2] for g in groups]
            buf = int(rng.choice([12, 40, 100000]))
            t0 = groups[0][0] * 10 + 600
            c = ChunkClock(cfg, t0, record_buffer=buf)
            one = c.feed(groups, nrec)
            one.append(c.finish())
            cuts = sorted(set(rng.integers(1, len(groups), rng.integers(1, 6)).tolist()))
            pieces = [(groups[a:b], nrec[a:b]) for a, b in zip([0] + cuts, cuts + [len(groups)])]
            c = ChunkClock(cfg, t0, record_buffer=buf)
            got = []
            for k, (pg, pn) in enumerate(pieces):
                got += c.feed(pg, pn)
                if k + 1 < len(pieces):
                    got += c.peek(pieces[k + 1][0][0][0] * 10 - int(rng.integers(0, 50000)))
            got.append(c.finish())
            if any(g[2] > 0 for g in pieces[-1][0]) or got == one:
                assert got == one, trial
            else:
                tails += 1
                assert got[:-2] == one[:-1] and got[-2][0] == one[-1][0] and got[-1][0] == got[-2][1], trial
        assert tails < 40
    finally:
        logging.disable(logging.NOTSET)


def test_dtypes_and_config_loader():
    from wfsim_b200 import dtypes, config
    assert np.dtype(dtypes.instruction_dtype).itemsize == 70
    assert dtypes.truth_dtype().itemsize == 218
    assert dtypes.raw_record_dtype().itemsize == 244
    assert dtypes.truth_dtype(494).fields['n_photon_per_pmt'][0].shape == (494,)
    txt = '// c\n{ "a": 1, # x\n "url": "http://x//y#z", "l": [1, 2,], }\n'
    assert config.loads_tolerant(txt) == {'a': 1, 'url': 'http://x//y#z', 'l': [1, 2]}
    cfg = load_c0_config()
    assert cfg['channel_map']['sum_signal'] == 800 and len(cfg['channels_bottom']) == 241
    assert cfg['field_distortion_model'] == 'none'
    np.testing.assert_allclose(cfg['gains'], 2.25 / 2 ** 14 / 8.010882825e-09 / 0.008)


def test_tables_match_reference_tables():
    from wfsim_b200 import tables
    cfg = load_c0_config()
    z = np.load(os.path.join(GOLDEN, 'c0_tables.npz'))
    np.testing.assert_array_equal(tables.pmt_current_templates(cfg), z['templates'])
    np.testing.assert_array_equal(tables.template_maxima(z['templates']), z['current_max'])
    thr = tables.zle_thresholds(dict(cfg, special_thresholds={'7': 40}))
    assert thr[0] == 16000 - 15 - 1 and thr[7] == 16000 - 40 - 1 and len(thr) == 801
    # SPE table: rebuild from a synthetic pdf and compare with scipy's interp1d(kind='next')
    from scipy.interpolate import interp1d
    charge = np.arange(-5.5, 30.5)
    pdf = np.exp(-0.5 * ((charge - 8) / 4) ** 2)
    uniq, idx = tables.spe_ppf_rows(charge, [pdf, pdf, np.zeros_like(pdf)])
    cdf = np.cumsum(pdf) / pdf.sum()
    want = interp1d(cdf, charge, kind='next', bounds_error=False, fill_value=(charge[0], charge[-1]))(
        np.linspace(0, 1, 2001))
    np.testing.assert_array_equal(uniq[idx[0]], want)
    assert idx[0] == idx[1] != idx[2] and np.all(uniq[idx[2]] == 0)


def test_library_exports_every_declared_symbol():
    from wfsim_b200 import lib
    L = lib.load()          # raises if the .so is missing, a symbol is absent or a struct differs
    with open(os.path.join(ROOT, 'include', 'wfsim_b200.h')) as f:
        hdr = f.read()
    declared = set(re.findall(r'\b(wfs_[a-z_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    for name in declared:
        assert hasattr(L, name), name
    assert set(lib.EXPORTS) >= declared


def test_no_cpu_fallback_without_gpu():
    """The product path must fail loudly when no CUDA device is present."""
    from wfsim_b200 import lib
    from wfsim_b200.simulator import Simulator, SimulatorError
    if lib.load().wfs_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(SimulatorError):
        Simulator(load_c0_config())


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'wfsim_b200')
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dp, fn)) as f:
                    txt = f.read()
                assert 'oracle' not in txt.replace('# oracle', ''), f'{fn} mentions the oracle'


def test_plugin_set_config_and_check_instructions():
    """Host logic of the plugin mirror that needs no GPU (strax_interface.py:566-608, 682-693)."""
    from wfsim_b200.strax_interface import RawRecordsFromFaxNT
    from tests.golden.synth_instructions import c0_like
    import json as _json
    with open(os.path.join(GOLDEN, 'c0_config.json')) as f:
        fax = _json.load(f)
    p = RawRecordsFromFaxNT(config=dict(fax_config=fax, gain_model_mc=np.full(494, 0.008), chunk_size=3,
                                        fax_config_override={'zle_threshold': 20}))
    p.instructions = c0_like(5, seed=1)
    p.set_config()
    assert p.config['chunk_size'] == 3 and p.config['zle_threshold'] == 20
    assert p.config['channel_map']['sum_signal'] == 800
    assert p.infer_dtype()['truth'] is not None and p.provides[0] == 'raw_records'
    p.check_instructions()
    for field, val, msg in (('amp', 0, 'Interaction has zero size'), ('z', 1.0, 'outside the TPC \\(in Z\\)'),
                            ('x', 80.0, 'outside the TPC \\(radius\\)')):
        bad = p.instructions.copy()
        bad[field][1] = val
        q = RawRecordsFromFaxNT(config=dict(fax_config=fax, gain_model_mc=np.full(494, 0.008)))
        q.instructions = bad
        q.set_config()
        with pytest.raises(AssertionError, match=msg):
            q.check_instructions()
    # S2 instructions below the cathode are dropped, S1 kept (strax_interface.py:684-685)
    deep = p.instructions.copy()
    deep['z'][:2] = -120.0
    q = RawRecordsFromFaxNT(config=dict(fax_config=fax, gain_model_mc=np.full(494, 0.008)))
    q.instructions = deep
    q.set_config()
    q.check_instructions()
    assert len(q.instructions) == len(deep) - 1
    # flip-flop is_ready
    assert [q.is_ready(i) for i in range(4)] == [True, False, True, False]
    # _sort_check
    from wfsim_b200.dtypes import raw_record_dtype
    r = np.zeros(3, raw_record_dtype()); r['time'] = [10, 5, 20]
    q.last_chunk_time = -10**15
    with pytest.raises(RuntimeError, match='non-sorted'):
        q._sort_check(r)
    r['time'] = [10, 15, 20]
    q._sort_check(r)
    r['time'] += 500
    with pytest.raises(RuntimeError, match='insufficient spacing'):
        q._sort_check(r)


def test_chunk_clock_fed_in_pieces_with_peek_gives_the_reference_bounds():
    """ChunkClock: feeding the groups piece by piece and closing chunks early with peek() (a lower bound on
    where the next piece starts) gives the bounds of the one-pass bookkeeping, i.e. the reference's
    (tests/golden/chunks.json), however the run is cut."""
    from wfsim_b200.strax_interface import ChunkClock
    with open(os.path.join(GOLDEN, 'chunks.json')) as f:
        cases = json.load(f)
    rng = np.random.default_rng(5)
    for name, c in cases.items():
        cfg = load_c0_config()
        cfg['chunk_size'] = c['chunk_size']
        groups = [tuple(g) for g in c['groups']]
        for trial in range(20):
            cuts = sorted(set(rng.integers(1, max(len(groups), 2), rng.integers(0, 6)).tolist()))
            pieces = [groups[a:b] for a, b in zip([0] + cuts, cuts + [len(groups)])]
            clock = ChunkClock(cfg, c['t_min_instruction'], record_buffer=c.get('record_buffer'))
            got = []
            for k, piece in enumerate(pieces):
                got += clock.feed(piece, [2 * g[2] for g in piece])
                if k + 1 < len(pieces) and pieces[k + 1]:
                    # the next piece starts at its first group's left edge; any earlier time is a valid bound
                    t_next = pieces[k + 1][0][0] * cfg['sample_duration'] - int(rng.integers(0, 50000))
                    got += clock.peek(t_next)
            got.append(clock.finish())
            assert [list(b) for b in got] == c['bounds'], (name, cuts)


def test_committed_ncu_traffic_is_what_bench_reads():
    """bench.py takes roofline.traffic from the committed ncu summary of the fused back end (profiles/r2_fused_ncu.json,
    written by profiles/tools/fused_traffic.py): one k_group_records launch and the size-class launches per batch."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t = bench.ncu_traffic()
    assert t and t['dram_bytes_per_batch'] > 1e9 and 'ncu --set full' in t['source']
    kernels = [l['kernel'] for l in t['launches']]
    assert kernels.count('k_group_records') == 1 and any(k.startswith('k_group_analyse') for k in kernels)
    assert abs(sum(l['dram_read'] + l['dram_write'] for l in t['launches']) - t['dram_bytes_per_batch']) < 1.0
