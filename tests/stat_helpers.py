"""Two-sample tests used for the stochastic stages (BASELINE.json: KS / chi-square, p > 0.01)."""
import numpy as np
from scipy import stats

P_MIN = 0.01


def ks_p(a, b):
    """Two-sample KS on (nearly) continuous data."""
    return stats.ks_2samp(np.asarray(a, float), np.asarray(b, float)).pvalue


def chi2_counts_p(a, b, min_expected=5):
    """Chi-square homogeneity test of two count vectors over the same categories."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    tot = a + b
    keep = tot * min(a.sum(), b.sum()) / (a.sum() + b.sum()) >= min_expected
    a, b = a[keep], b[keep]
    if len(a) < 2:
        return 1.0
    return stats.chi2_contingency(np.stack([a, b]))[1]


def discrete_p(a, b):
    """Chi-square on the pooled histogram of two integer samples (ties make KS conservative)."""
    a, b = np.asarray(a).astype(np.int64), np.asarray(b).astype(np.int64)
    lo, hi = min(a.min(), b.min()), max(a.max(), b.max())
    nb = int(min(hi - lo + 1, 60))
    edges = np.linspace(lo, hi + 1, nb + 1)
    return chi2_counts_p(np.histogram(a, edges)[0], np.histogram(b, edges)[0])


def mean_p(a, b):
    """Welch t-test on means (for per-call aggregates with few samples)."""
    return stats.ttest_ind(np.asarray(a, float), np.asarray(b, float), equal_var=False).pvalue
