"""BASELINE.json config [4] (pulse-superposition microbench) through wfs_simulate_photons: K pulses x
1e6 photons, t ~ round(t0_k + N(0, 1000 ns)), channels 62.7 % top / 37.3 % bottom, gain ~ gains[ch] (0.3 + Exp(0.7))
(SURVEY.md section 8d).  Prints device time (photons resident), phases and photons/s."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.conftest import load_c0_config
from wfsim_b200.simulator import Simulator

K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
cfg = load_c0_config(zle_threshold=0)      # ZLE effectively off: only the pulse windows themselves matter
gains = np.asarray(cfg['gains'])
rng = np.random.default_rng(0)
n = K * per
t = (np.repeat(np.arange(K, dtype=np.int64) * 1_000_000, per) + np.round(rng.normal(0, 1000, n)).astype(np.int64)
     + 100_000)
top = rng.random(n) < 0.627
ch = np.where(top, rng.integers(0, 253, n), rng.integers(253, 494, n)).astype(np.int32)
g = gains[ch] * (0.3 + rng.exponential(0.7, n))
pcall = np.repeat(np.arange(K, dtype=np.int32), per)
group_of = np.arange(K, dtype=np.int32)
sim = Simulator(cfg)
for rep in range(3):
    t0 = time.perf_counter()
    out = sim.simulate_photons(t, ch, g, pcall, group_of, cap_records=40 * K * 494)
    wall = time.perf_counter() - t0
    c = sim.last_counts
    print(f'C4 {K} pulses x {per} photons: device {c["ms_total"]:.1f} ms (h2d {c["ms_h2d"]:.1f}, d2h {c["ms_d2h"]:.1f}), '
          f'{n / c["ms_total"] * 1e3:.3g} photons/s, {c["n_records_total"]} records, {c["n_samples"]:.3g} samples, '
          f'phases {[round(x, 1) for x in c["ms_phase"][1:7]]}, wall {wall:.2f} s', flush=True)
