"""Golden vectors for the optical pre-step (wfsim/utils.py:121-165): random photon lists through the UNMODIFIED
reference `optical_adjustment`; inputs and outputs stored for tests/test_optical.py."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case(idt, seed, n_inst, mean_photons, p_empty, late_frac):
    rng = np.random.default_rng(seed)
    n_ph = rng.poisson(mean_photons, n_inst)
    n_ph[rng.random(n_inst) < p_empty] = 0
    last = np.cumsum(n_ph)
    first = last - n_ph
    inst = np.zeros(n_inst, idt)
    inst['time'] = 1_000_000 * (1 + np.arange(n_inst))
    inst['type'] = 1
    inst['_first'], inst['_last'] = first, last
    inst['amp'] = n_ph
    inst['event_number'] = np.arange(n_inst)
    t = rng.exponential(60.0, last[-1] if n_inst else 0).astype(np.int64) + rng.integers(0, 400, 1)[0]
    late = rng.random(len(t)) < late_frac
    t[late] += rng.integers(900, 5000, late.sum())
    ch = rng.integers(0, 494, len(t)).astype(np.int64)
    return inst, t, ch


def main(ref):
    si = ref.strax_interface
    idt = np.dtype(si.instruction_dtype + si.optical_extra_dtype)
    out = {}
    for k, (seed, n, mean, p_empty, late) in enumerate([(1, 60, 30, 0.1, 0.02), (2, 40, 8, 0.3, 0.0), (3, 25, 200, 0.0, 0.05),
                                                        (4, 1, 5, 0.0, 0.5)]):
        inst, t, ch = make_case(idt, seed, n, mean, p_empty, late)
        out[f'in_inst_{k}'], out[f'in_t_{k}'], out[f'in_ch_{k}'] = inst.copy().view(np.uint8), t.copy(), ch.copy()
        res = ref.utils.optical_adjustment(inst, t, ch)
        out[f'out_inst_{k}'], out[f'out_t_{k}'], out[f'out_ch_{k}'] = np.asarray(res).view(np.uint8), t, ch
        print(k, len(inst), '->', len(res), 'instructions,', len(t), 'photons')
    out['n_cases'] = np.array(4)
    np.savez_compressed(os.path.join(HERE, 'optical_adjustment.npz'), **out)


if __name__ == '__main__':
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    import logging
    logging.disable(logging.WARNING)
    from oracle import ref_loader as RL
    main(RL.load_reference())
