"""Synthetic stand-ins for the resource files that are not in the reference tree:
PMT afterpulse CDFs (`photon_ap_cdfs`, load_resource.py:94), the electron-afterpulse delay
histogram (`ele_ap_pdfs`, load_resource.py:361) and a noise sample (`noise_file`, :99)."""
import numpy as np


def pmt_ap_tables(n_ch=494):
    """uniform_to_pmt_ap as afterpulse.py:172-249 expects: per element delaytime_cdf [n_ch, n] rising
    to the afterpulse probability, amplitude_cdf, bin sizes; 'Uniform' elements carry
    [t_low_bin, t_high_bin, probability] rows."""
    delay_bins, amp_bins = 1000, 100
    x = np.linspace(0, 1, delay_bins)
    base = 0.02 * (1 - np.exp(-4 * x)) / (1 - np.exp(-4.0))           # rises to 2 %
    scale = 0.8 + 0.4 * (np.arange(n_ch) % 7) / 6.0                   # per-PMT probability
    he = dict(delaytime_cdf=base[None, :] * scale[:, None],
              amplitude_cdf=np.tile(np.linspace(0, 1, amp_bins) ** 0.7, (n_ch, 1)),
              delaytime_bin_size=10.0, amplitude_bin_size=0.13)
    uni = dict(delaytime_cdf=np.tile(np.array([50.0, 500.0, 0.005]), (n_ch, 1)),
               amplitude_cdf=np.zeros((n_ch, 2)), delaytime_bin_size=1.0, amplitude_bin_size=1.0)
    return {'He': he, 'Uniform': uni}


class EleApHist:
    """Minimal multihist.Hist1d look-alike: `.n`, `.bin_centers`, `.histogram`, `.bin_edges` and
    `.get_random(size)` (pick a bin by weight, uniform inside the bin)."""

    def __init__(self, n=1e-3, t_max=7.3e5, n_bins=730, tau=2e5):
        self.n = n
        self.bin_edges = np.linspace(1e3 - 500, t_max + 500, n_bins + 1)
        self.bin_centers = 0.5 * (self.bin_edges[1:] + self.bin_edges[:-1])
        self.histogram = np.exp(-self.bin_centers / tau)
        self._p = self.histogram / self.histogram.sum()

    def get_random(self, size=10, rng=None):
        rnd = np.random if rng is None else rng
        bin_i = rnd.choice(np.arange(len(self.bin_centers)), size=size, p=self._p)
        return self.bin_centers[bin_i] + rnd.uniform(-0.5, 0.5, size=size) * np.diff(self.bin_edges)[bin_i]


def noise_sample(n_ch=494, length=1 << 16, seed=5):
    rng = np.random.default_rng(seed)
    return np.round(rng.normal(0, 2, (length, n_ch))).astype(np.float64)
