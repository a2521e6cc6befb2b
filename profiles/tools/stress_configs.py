"""Larger runs of BASELINE.json configs [2] and [3] on the GPU: invariants + throughput (not a test;
the parity cases at oracle-sized inputs live in tests/test_gpu_configs.py)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.test_gpu_configs import make_sim, check_records_sorted_and_consistent, heavy_s2_events
from tests.golden.synth_instructions import c0_like, c1_like
from tests.golden.synth_tables import EleApHist, pmt_ap_tables


def run(name, sim, cfg, inst, seed=1):
    t0 = time.perf_counter()
    out = sim.simulate(inst, seed=seed)
    dt = time.perf_counter() - t0
    check_records_sorted_and_consistent(out, cfg)
    c = sim.last_counts
    print(f'{name}: {len(inst)} instr, {c["n_photons"]:.3g} photons, {c["n_records_total"]:.3g} records, '
          f'{c["n_batches"]} batches, {len(out["truth"])} truth rows; wall {dt:.2f} s, device {c["ms_total"]:.0f} ms, '
          f'{c["n_pe"] / dt:.3g} pe/s e2e; samples {c["n_samples"]:.3g}; phases {[round(x) for x in c["ms_phase"][:10]]}; segment-sorted photon/record batches {c["ms_phase"][10]:.0f}/{c["ms_phase"][11]:.0f}',
          flush=True)
    return out


n2 = int(sys.argv[1]) if len(sys.argv) > 1 else 400
n3 = int(sys.argv[2]) if len(sys.argv) > 2 else 30000
# C2: S2-heavy events, PMT afterpulses + photo-ionisation
sim, cfg = make_sim(dict(uniform_to_pmt_ap=pmt_ap_tables(494), uniform_to_ele_ap=EleApHist()),
                    enable_pmt_afterpulses=True, enable_electron_afterpulses=True)
inst = heavy_s2_events(n2, seed=5, n_e=(10_000, 100_000))
for k in range(2):
    run('C2', sim, cfg, inst)
sim.close()
if n3 <= 0:
    sys.exit(0)
# C3: mixed stream (95 % low energy, 5 % heavy), noise + ZLE
rng = np.random.default_rng(3)
noise = np.round(rng.normal(0, 2.0, (1 << 16, 494)))
sim, cfg = make_sim(dict(noise_data=noise), enable_noise=True)
lo = c1_like(n3, seed=7)
hi = heavy_s2_events(max(n3 // 20, 1), seed=8)
hi['time'] += np.int64(3_000_000)          # interleave: shift the heavy events between the light ones
inst = np.concatenate([lo, hi])
inst = inst[np.argsort(inst['time'], kind='stable')]
for k in range(2):
    run('C3', sim, cfg, inst)
sim.close()
