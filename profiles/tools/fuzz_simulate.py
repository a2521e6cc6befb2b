"""Randomised small instruction sets through wfs_simulate: no errors, sorted records, and the records equal
the oracle's deterministic back end on the photons the GPU generated (exact).  Not a test (run on a GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import wfsim_oracle as orc
from tests.test_gpu_stochastic import make_sim, group_of_photons
from tests.test_gpu_configs import check_records_sorted_and_consistent
from wfsim_b200.dtypes import instruction_dtype

if os.environ.get('FUZZ_AP'):        # PMT afterpulses + photo-ionisation secondaries
    from tests.test_gpu_configs import make_sim as make_sim_res
    from tests.golden.synth_tables import EleApHist, pmt_ap_tables
    sim, cfg = make_sim_res(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                            enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
else:
    sim, cfg = make_sim()
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_iter = int(sys.argv[2]) if len(sys.argv) > 2 else 150
bad = 0
for it in range(n_iter):
    n = int(rng.integers(0, 40)) if it % 4 else int(rng.integers(100, 600))
    inst = np.zeros(n, instruction_dtype)
    inst['type'] = rng.choice([1, 2], n)
    inst['time'] = np.sort(rng.integers(0, int(10 ** rng.uniform(4, 9)), n)) if rng.random() < 0.7 else rng.integers(0, 10 ** 8, n)
    r = np.sqrt(rng.uniform(0, 55 ** 2, n)); th = rng.uniform(-np.pi, np.pi, n)
    inst['x'], inst['y'] = r * np.cos(th), r * np.sin(th)
    inst['z'] = rng.uniform(-110, 5, n)
    inst['amp'] = np.where(rng.random(n) < 0.2, rng.integers(1, 4, n), (10 ** rng.uniform(0, 3.7 if n < 100 else 2.5, n)).astype(int))
    inst['recoil'] = 7
    inst['local_field'] = 82.0
    inst['event_number'] = np.arange(n)
    inst = inst[inst['amp'] > 0]
    # forced batch cuts are a documented deviation once secondaries exist: keep the default there
    os.environ['WFS_BATCH_INSTRUCTIONS'] = str(int(rng.choice([3, 7, 400000]))) if not os.environ.get('FUZZ_AP') else '400000'
    try:
        out = sim.simulate(inst, seed=it)
        check_records_sorted_and_consistent(out, cfg)
        if len(inst) == 0:
            assert len(out['raw_records']) == 0 and len(out['truth']) == 0
            continue
        ph = sim.sample_stage(inst, stage=0, seed=it)
        ph = ph[ph['channel'] >= 0]
        prim = ph[(ph['flags'] >> 1) & 1 == 0]
        assert out['truth']['n_photon'].sum() == len(prim), (out['truth']['n_photon'].sum(), len(prim))
        if it % 3 == 0 and len(ph) and len(out['groups']):
            g_of_ph = group_of_photons(ph, out['groups'], cfg)
            # Pulse-call identity: (primary instruction | secondary cluster = its group) x afterpulse flag
            base = np.where((ph['flags'] & 4) != 0, 10_000_000 + g_of_ph, ph['instruction'])
            pcall = base * 2 + ((ph['flags'] >> 1) & 1)
            uniq, pc = np.unique(pcall, return_inverse=True)
            group_of = np.zeros(len(uniq), np.int32)
            group_of[pc] = g_of_ph
            want = orc.simulate_photons(cfg, pc.astype(np.int32), ph['channel'], ph['t'], ph['gain'], group_of)
            if out['raw_records'].tobytes() != want['raw_records'].tobytes():
                if os.environ.get('FUZZ_AP'):
                    # with secondaries the Pulse-call identity is only approximated here (all type-4 photons of a
                    # group = one call; the scheduler makes one call per instruction cluster): piled-up
                    # secondaries can legitimately differ.  The exact AP + PI case is tests/test_gpu_afterpulse_plugin.py
                    print('unverified (approximate Pulse-call reconstruction) at iteration', it, flush=True)
                    continue
                bad += 1
                print('MISMATCH at iteration', it, len(inst), len(ph), len(out['raw_records']), len(want['raw_records']), flush=True)
    except Exception as e:      # noqa
        bad += 1
        print('ERROR at iteration', it, repr(e)[:300], flush=True)
print('done', n_iter, 'iterations,', bad, 'problems')
