"""Seeded synthetic photon sets for the deterministic parity tests (inputs only)."""
import numpy as np


def synth_photons(cfg, rng, n_groups, t_start=1_000_000_000, big=False):
    """Synthetic (pulse-call, channel, time, gain) photons grouped into digitisation groups."""
    gains = np.asarray(cfg['gains'])
    n_ch = len(gains)
    pcall, ch, t, g, group_of = [], [], [], [], []
    pc = 0
    t0 = t_start
    for grp in range(n_groups):
        n_pc = int(rng.integers(1, 4))
        for k in range(n_pc):
            kind = rng.integers(0, 3)
            if kind == 0:      # S1-like: few photons, 50 ns scale
                n = int(rng.integers(1, 120))
                tt = t0 + rng.exponential(45, n) + rng.normal(0, 5, n)
            elif kind == 1:    # S2-like: many photons, microsecond scale
                n = int(rng.integers(500, 6000 if not big else 60000))
                tt = t0 + 3000 * k + rng.normal(0, 600, n) + rng.exponential(150, n)
            else:              # afterpulse-like: sparse, late, large gains
                n = int(rng.integers(1, 40))
                tt = t0 + rng.uniform(500, 9000, n)
            c = rng.integers(0, n_ch, n)
            if kind == 1:
                # a hot top channel and coincident photons (equal-ns merging, pulse.py:301-318)
                c[: n // 10] = 7
                tt[: n // 20] = np.round(tt[: n // 20] / 10) * 10
            amp = 0.3 + rng.exponential(0.7, n)
            amp[rng.random(n) < 0.03] *= -0.4       # SPE table has negative charges
            if kind == 2:
                amp *= rng.integers(1, 40, n)       # drives some samples to the 0 clamp
            pcall.append(np.full(n, pc)); ch.append(c); t.append(np.floor(tt).astype(np.int64))
            g.append(gains[c] * amp)
            group_of.append(grp)
            pc += 1
        t0 += int(rng.integers(300_000, 2_000_000))
    return (np.concatenate(pcall).astype(np.int32), np.concatenate(ch).astype(np.int32),
            np.concatenate(t), np.concatenate(g), np.asarray(group_of, np.int32))
