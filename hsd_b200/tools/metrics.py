"""Host-side scalar mirror of tools/metrics.py::calculate_distance for single
pairs (API compatibility, every metric name and exception of the reference); the
all-pairs versions of the three metrics the HSD path reaches run on the GPU
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


def check_probablity_distribution(p, q):
    """tools/metrics.py:39-51: type, length and sum-to-one checks (same exceptions)."""
    if not (isinstance(p, (list, np.ndarray)) and isinstance(q, (list, np.ndarray))):
        raise TypeError("The probability distribution must be list or ndarray.")
    if len(p) != len(q):
        raise TypeError("Length of p({}) must be equal to length of q({})".format(len(p), len(q)))
    if not math.isclose(np.sum(p) - 1.0, 0.0, abs_tol=1e-4) or \
            not math.isclose(np.sum(q) - 1.0, 0.0, abs_tol=1e-4):
        raise ValueError("The sum of probability distribution must be 1.0.")


def KL_divergence(p, q, symmetric=False):
    """tools/metrics.py:75-90."""
    check_probablity_distribution(p, q)
    kl_pq = np.sum(p * np.log(p / q))
    if symmetric:
        return (kl_pq + np.sum(q * np.log(q / p))) / 2.0
    return kl_pq


def JS_divergence(p, q):
    """tools/metrics.py:93-106 as written: the mixture is m = p + q (not halved), so the inner
    KL calls fail the sum-to-one check exactly like the reference's do."""
    check_probablity_distribution(p, q)
    m = p + q
    return (KL_divergence(p, m, False) + KL_divergence(q, m, False)) / 2.0


def L_distance(p, q, order):
    """tools/metrics.py:109-117: sum |p - q| (order 1) or sum (p - q)^2 (order 2, no root)."""
    check_probablity_distribution(p, q)
    if order == 1:
        return np.sum(np.abs(p - q))
    elif order == 2:
        return np.sum(np.square(p - q))


def calculate_distance(p, q, metric):
    """tools/metrics.py:151-192, every metric name, same exceptions: TypeError without a metric,
    NotImplementedError for an unknown one, ValueError from the sum-to-one check of
    l1 / l2 / kl / symmetric_kl / js on un-normalised input (which is what HSD ring signals are)."""
    if not metric or not isinstance(metric, str):
        raise TypeError("Need to specify a metric.")
    metric = metric.lower()
    if metric not in SUPPORTED:
        raise NotImplementedError("{} metric is not implemented.".format(metric))
    p, q = align_probablity_distribution(list(p), list(q))
    if len(p) == 0 and len(q) == 0:
        return 0.0
    if metric == 'l1':
        return L_distance(p, q, order=1)
    if metric == 'l2':
        return L_distance(p, q, order=2)
    if metric == 'kl':
        return KL_divergence(p, q, symmetric=False)
    if metric == 'symmetric_kl':
        return KL_divergence(p, q, symmetric=True)
    if metric == 'js':
        return JS_divergence(p, q)
    if metric == 'wasserstein':
        return float(np.mean(np.abs(p - q)))
    if metric == 'hellinger':
        return hellinger_distance(p, q)
    u1, u2 = np.mean(p), np.mean(q)     # 'wasserstein_guass', tools/metrics.py:54-71
    s1, s2 = np.mean(np.square(p - u1)), np.mean(np.square(q - u2))
    return float((u1 - u2) ** 2 + s1 + s2 - 2 * (s1 * s2) ** 0.5)
