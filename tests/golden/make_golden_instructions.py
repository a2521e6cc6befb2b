"""Golden samples of the reference's instruction generator (strax_interface.py:155-231), run unmodified with a
stand-in `nestpy` module whose yields are floor(45 E) photons / floor(28 E) electrons (the real nestpy is a
third-party package absent from this image).  Stored: the deterministic columns of one call and samples of the
random ones, for the KS tests of wfsim_b200.instructions."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def install_nestpy_stub():
    m = types.ModuleType('nestpy')

    class VDetector:
        pass

    class Quanta:
        def __init__(self, e):
            self.photons, self.electrons, self.excitons = int(np.floor(45 * e)), int(np.floor(28 * e)), 0

    class NESTcalc:
        def __init__(self, det):
            pass

        def GetYields(self, interaction, energy, density, field, A, Z):
            return energy

        def GetQuanta(self, y, density):
            return Quanta(y)
    m.VDetector, m.NESTcalc, m.INTERACTION_TYPE = VDetector, NESTcalc, (lambda i: i)
    sys.modules['nestpy'] = m


def main(ref):
    install_nestpy_stub()
    np.random.seed(20260)
    si = ref.strax_interface
    kw = dict(event_rate=50, chunk_size=5, n_chunk=4, drift_field=82.0, energy_range=[1, 100], tpc_length=97.0,
              tpc_radius=50.0, nest_inst_types=[7, 0])
    inst = si._rand_instructions(**kw)
    c = dict(event_rate=3, chunk_size=2, n_chunk=3, drift_field=100, tpc_radius=40.0, tpc_length=80.0)
    inst_c = si.rand_instructions(c)
    out = dict(kwargs=np.array(repr(kw)), inst=inst.view(np.uint8), config=np.array(repr(c)), inst_config=inst_c.view(np.uint8))
    np.savez_compressed(os.path.join(HERE, 'rand_instructions.npz'), **out)
    print(len(inst), len(inst_c), inst.dtype.itemsize)


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import logging
    logging.disable(logging.WARNING)
    from oracle import ref_loader as RL
    main(RL.load_reference())
