"""Host-side scalar mirror of tools/metrics.py::calculate_distance for single
pairs (API compatibility); the all-pairs versions run on the GPU
(hsd_pairwise_aligned / hsd_pairwise_w1_merge)."""
import math

import numpy as np

SUPPORTED = ['l1', 'l2', 'kl', 'symmetric_kl', 'js', 'wasserstein_guass', 'wasserstein', 'hellinger']


def align_probablity_distribution(p, q, normalized=False):
    """Zero-pad to equal length and sort ascending (tools/metrics.py:18-36)."""
    n = max(len(p), len(q))
    p = np.sort(np.concatenate([np.asarray(p, dtype=float), np.zeros(n - len(p))]))
    q = np.sort(np.concatenate([np.asarray(q, dtype=float), np.zeros(n - len(q))]))
    if normalized:
        p = p / p.sum() if p.sum() > 0.0 else p
        q = q / q.sum() if q.sum() > 0.0 else q
    return p, q


def hellinger_distance(p, q):
    bc = sum(math.sqrt(max(a * b, 0)) for a, b in zip(p, q) if a >= 0 and b >= 0)
    if math.isclose(bc, 0.0, abs_tol=1e-6):
        bc = 0.0
    elif math.isclose(bc, 1.0, abs_tol=1e-6):
        bc = 1.0
    return math.sqrt(max(1.0 - bc, 0))


def calculate_distance(p, q, metric):
    """tools/metrics.py:151-192 for the two metrics the HSD path uses."""
    if not metric or not isinstance(metric, str):
        raise TypeError("Need to specify a metric.")
    metric = metric.lower()
    if metric not in SUPPORTED:
        raise NotImplementedError("{} metric is not implemented.".format(metric))
    p, q = align_probablity_distribution(list(p), list(q))
    if len(p) == 0 and len(q) == 0:
        return 0.0
    if metric == 'wasserstein':
        return float(np.mean(np.abs(p - q)))
    if metric == 'hellinger':
        return hellinger_distance(p, q)
    if metric == 'wasserstein_guass':
        u1, u2 = np.mean(p), np.mean(q)
        s1, s2 = np.mean(np.square(p - u1)), np.mean(np.square(q - u2))
        return float((u1 - u2) ** 2 + s1 + s2 - 2 * (s1 * s2) ** 0.5)
    raise NotImplementedError("{} is outside the HSD hot path; use the reference's tools/metrics.py".format(metric))
