"""One fused call with PMT afterpulses (for compute-sanitizer --tool racecheck / memcheck)."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.golden.synth_instructions import c1_like
from tests.test_gpu_afterpulse_plugin import make_sim
sim, cfg = make_sim(enable_pmt_afterpulses=True)
inst = c1_like(int(sys.argv[1]) if len(sys.argv) > 1 else 400, seed=9)
for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1):
    out = sim.simulate(inst, seed=5)
    r = out['raw_records']
    print(len(r), hashlib.md5(r.tobytes()).hexdigest()[:10], sim.last_counts['n_fused_batches'], flush=True)
