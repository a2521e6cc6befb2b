"""Instruction generation: the step in front of the hot path.

Mirror of rand_instructions / random_instructions / _rand_instructions (strax_interface.py:119-231): one S1 and
one S2 row per event, events equally spaced over the run, positions uniform in the cylinder, energies uniform
in `energy_range`, quanta from a yield model.  The reference draws the quanta event by event from nestpy (a
third-party package that is absent here and on the GPU boxes); the yield model is therefore a parameter:

    yields(energy_keV[n], interaction_type[n], drift_field) -> (photons[n], electrons[n], excitons[n])

`nest_yields` calls nestpy when it can be imported, `fixed_yields` is the stand-in BASELINE config C0 is defined
with (amp_S1 = floor(45 E), amp_S2 = floor(28 E)).  Everything else is vectorised numpy on a
seeded Generator (1e6 events take a fraction of a second; the reference's per-event Python loop takes minutes).
"""
import logging

import numpy as np

from .dtypes import instruction_dtype

log = logging.getLogger('wfsim_b200.instructions')

TPC_R, TPC_Z = 66.4, 148.6515        # straxen.tpc_r, straxen.tpc_z (defaults of the reference's signature)


def fixed_yields(energy, interaction_type, drift_field, s1_per_kev=45.0, s2_per_kev=28.0):
    """Deterministic stand-in for NEST: floor(45 E) photons, floor(28 E) electrons, no excitons."""
    energy = np.asarray(energy, np.float64)
    return np.floor(s1_per_kev * energy), np.floor(s2_per_kev * energy), np.zeros(len(energy))


def nest_yields(energy, interaction_type, drift_field):
    """The reference's yield model (strax_interface.py:193-221); needs nestpy."""
    import nestpy
    calc = nestpy.NESTcalc(nestpy.VDetector())
    A, Z, density = 131.293, 54., 2.862
    ph, el, ex = [], [], []
    for e, it in zip(energy, interaction_type):
        y = calc.GetYields(nestpy.INTERACTION_TYPE(int(it)), e, density, drift_field, A, Z)
        q = calc.GetQuanta(y, density)
        ph.append(q.photons); el.append(q.electrons); ex.append(q.excitons)
    return np.array(ph), np.array(el), np.array(ex)


def default_yields():
    try:
        import nestpy  # noqa: F401
        return nest_yields
    except ImportError:
        log.warning('nestpy is not installed: instructions get the fixed-yield stand-in (45 photons, 28 electrons per keV)')
        return fixed_yields


def _rand_instructions(event_rate, chunk_size, n_chunk, drift_field, energy_range, tpc_length=TPC_Z, tpc_radius=TPC_R,
                       nest_inst_types=None, yields=None, seed=None):
    """strax_interface.py:155-231.  `seed`: numpy Generator seed (None: fresh entropy, as the reference's
    unseeded global generator)."""
    if nest_inst_types is None:
        nest_inst_types = [7]
    rng = np.random.default_rng(seed)
    n_events = int(event_rate * chunk_size * n_chunk)
    total_time = chunk_size * n_chunk
    inst = np.zeros(2 * n_events, dtype=instruction_dtype)
    for name in inst.dtype.names:          # inst[:] = -1 of the reference: whatever is not filled below stays -1
        inst[name] = -1
    uniform_times = total_time * (np.arange(n_events) + 0.5) / n_events
    inst['time'] = np.repeat(uniform_times, 2) * int(1e9)
    inst['event_number'] = np.digitize(inst['time'], 1e9 * np.arange(n_chunk) * chunk_size) - 1
    inst['type'] = np.tile([1, 2], n_events)
    r = np.sqrt(rng.uniform(0, tpc_radius ** 2, n_events))
    t = rng.uniform(-np.pi, np.pi, n_events)
    inst['x'] = np.repeat(r * np.cos(t), 2)
    inst['y'] = np.repeat(r * np.sin(t), 2)
    inst['z'] = np.repeat(rng.uniform(-tpc_length, 0, n_events), 2)
    inst['x_pri'], inst['y_pri'], inst['z_pri'] = inst['x'], inst['y'], inst['z']
    energy = rng.uniform(*energy_range, n_events)
    itype = rng.choice(np.asarray(nest_inst_types), n_events)
    photons, electrons, excitons = (yields or default_yields())(energy, itype, drift_field)
    inst['amp'] = np.stack([photons, electrons], axis=1).ravel()
    inst['local_field'] = drift_field
    inst['n_excitons'] = np.stack([excitons, np.zeros(n_events)], axis=1).ravel()
    inst['recoil'] = np.repeat(itype, 2)
    inst['e_dep'] = np.repeat(energy, 2)
    for field in inst.dtype.names:
        if np.any(inst[field] == -1):
            log.warning(f'{field} is not (fully) filled')
    return inst


def random_instructions(**kwargs):
    """strax_interface.py:138-152."""
    return _rand_instructions(**kwargs)


def rand_instructions(c, yields=None, seed=None):
    """strax_interface.py:119-135: the generator the plugin falls back to when no instructions are given."""
    log.warning('rand_instructions is deprecated, please use wfsim.random_instructions')
    if 'drift_field' not in c:
        log.warning('drift field not specified!')
    return _rand_instructions(event_rate=c.get('event_rate', 10), chunk_size=c.get('chunk_size', 5),
                              n_chunk=c.get('n_chunk', 2), energy_range=[1, 100], drift_field=c.get('drift_field', 100),
                              tpc_radius=c.get('tpc_radius', TPC_R), tpc_length=c.get('tpc_length', TPC_Z),
                              nest_inst_types=[7], yields=yields, seed=seed)
