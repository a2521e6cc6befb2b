"""Seeded synthetic wfsim_instructions (inputs only) following the distributions of the
reference's rand_instructions (strax_interface.py:155-231) with nestpy replaced by fixed
yields, as laid out in SURVEY.md section 8(d)."""
import numpy as np

from wfsim_b200.dtypes import instruction_dtype


def c0_like(n_events, seed=0, event_rate=1.0, tpc_radius=50.0, tpc_length=97.0, e_range=(1, 100),
            s1_per_kev=45.0, s2_per_kev=28.0, recoil=7, t_start=0.0, poisson=False):
    rng = np.random.default_rng(seed)
    inst = np.zeros(2 * n_events, dtype=instruction_dtype)
    total_time = n_events / event_rate
    times = t_start + total_time * (np.arange(n_events) + 0.5) / n_events
    inst['time'] = (np.repeat(times, 2) * 1e9).astype(np.int64)
    inst['event_number'] = np.repeat(np.arange(n_events), 2)
    inst['type'] = np.tile([1, 2], n_events)
    r = np.sqrt(rng.uniform(0, tpc_radius ** 2, n_events)) * 0.999
    th = rng.uniform(-np.pi, np.pi, n_events)
    inst['x'] = np.repeat(r * np.cos(th), 2)
    inst['y'] = np.repeat(r * np.sin(th), 2)
    inst['z'] = np.repeat(rng.uniform(-tpc_length, 0, n_events), 2)
    for f in 'xyz':
        inst[f + '_pri'] = inst[f]
    e = rng.uniform(*e_range, n_events)
    if poisson:
        a1, a2 = rng.poisson(s1_per_kev * e), rng.poisson(s2_per_kev * e)
    else:
        a1, a2 = np.floor(s1_per_kev * e), np.floor(s2_per_kev * e)
    amp = np.stack([a1, a2], axis=1).reshape(-1)
    inst['amp'] = amp
    inst['recoil'] = recoil
    inst['e_dep'] = np.repeat(e, 2)
    inst['local_field'] = 82.0
    inst['g4id'] = -1
    inst['vol_id'] = -1
    inst['n_excitons'] = 0
    return inst[inst['amp'] > 0]


def c1_like(n_events, seed=0):
    """1e5 low-energy NR-like events at 1 kHz (BASELINE config[1])."""
    return c0_like(n_events, seed=seed, event_rate=1000.0, e_range=(1, 50), s1_per_kev=6.0,
                   s2_per_kev=5.0, recoil=0, poisson=True)
