"""Device-resident C2 run (heavy S2s, PMT afterpulses + photo-ionisation): per-phase times.  usage: c2_probe.py [events] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sim, cfg = bench.make_sim('C2', 0)
sim.stage(bench.heavy_events(n, seed=200))
for k in range(steps):
    c = sim.run_staged(seed=1)
    names = ['front', 'sort', 'win', 'digi', 'zle', 'rsort', 'pack', 'host']
    print(f"step {k}: {c['ms_total']:.1f} ms, batches {c['n_batches']} fused {c['n_fused_batches']}, photons {c['n_photons']:.3g}, "
          f"records {c['n_records_total']:.3g}, samples {c['n_samples']:.3g}; " + ' '.join(f'{a} {b:.1f}' for a, b in zip(names, c['ms_phase'][:8])), flush=True)
sim.close()
