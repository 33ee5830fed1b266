"""CPU oracle for the HSD structural-distance hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``hsd_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and only as the checker or the timed baseline.

Every function restates one piece of the reference (Sngunfei/HSD, pure Python)
in plain numpy / scipy float64 and cites the reference ``file:line`` it follows.
Pinning status (see DESIGN.md §4 and tests/test_oracle_golden.py):

* rings, exact heat kernel, ring [sum, mean, var], serial distance loop, aligned
  distance: PINNED — against outputs of the unmodified reference run in the
  build container (tests/golden/*.npz, made by oracle/make_golden.py) and
  against the reference's own golden vector tests/robust_test/robust.csv.
* W1: PINNED against scipy.stats.wasserstein_distance (the reference's own
  dependency, installed here) and the known answer in tests/other_test/main.py:9-13.
* Chebyshev (pygsp) path: PARITY UNPINNED at the bit level — pygsp is an
  un-vendored, unpinned third-party dependency that is absent here; the
  restatement follows the published pygsp 0.5.1 algorithm and is pinned only
  mathematically (robust.csv to 2.2e-8, exact kernel to ~1e-15 for small alpha).
"""
from __future__ import annotations

import math
from typing import Dict, Hashable, Iterable, List, Sequence

import numpy as np

try:  # scipy is the reference's own W1 implementation (model/HSD.py:10)
    from scipy.stats import wasserstein_distance as _scipy_w1
except Exception:  # pragma: no cover - scipy is present in this image
    _scipy_w1 = None


# --------------------------------------------------------------------------
# graph helpers (plain adjacency lists; node order = order of first appearance,
# tools/util.py:11-24 / model/HSD.py:37)
# --------------------------------------------------------------------------
def adjacency_from_edges(n: int, edges: np.ndarray) -> List[np.ndarray]:
    """Undirected simple graph -> list of sorted neighbour arrays."""
    nbrs = [set() for _ in range(n)]
    for u, v in np.asarray(edges, dtype=np.int64):
        nbrs[int(u)].add(int(v))
        nbrs[int(v)].add(int(u))
    return [np.array(sorted(s), dtype=np.int64) for s in nbrs]


def adjacency_from_networkx(graph) -> (List[Hashable], List[np.ndarray]):
    nodes = list(graph.nodes())
    idx = {v: i for i, v in enumerate(nodes)}
    adj = [np.array(sorted(idx[w] for w in graph.neighbors(v)), dtype=np.int64) for v in nodes]
    return nodes, adj


# --------------------------------------------------------------------------
# rings — tools/hierarchy.py:25-38
# --------------------------------------------------------------------------
def rings_of(adj: Sequence[np.ndarray], node: int, max_hop: int) -> List[List[int]]:
    """Level-synchronous BFS; layer h = nodes at distance exactly h; empty
    layers kept (tools/hierarchy.py:25-38).  Layers are returned sorted because
    the reference's order is set-iteration order, i.e. unspecified."""
    layers = [[node]]
    cur = {node}
    visited = {node}
    for _ in range(1, max_hop + 1):
        nxt = set()
        for v in cur:
            for w in adj[v]:
                w = int(w)
                if w not in visited:
                    nxt.add(w)
                    visited.add(w)
        cur = nxt
        layers.append(sorted(nxt))
    return layers


def all_rings(adj: Sequence[np.ndarray], max_hop: int,
              sources: Iterable[int] | None = None) -> Dict[int, List[List[int]]]:
    """tools/hierarchy.py:16-22 for every (or the given) source."""
    src = range(len(adj)) if sources is None else sources
    return {int(s): rings_of(adj, int(s), max_hop) for s in src}


def ring_sizes(rings: Dict[int, List[List[int]]], hop: int) -> Dict[int, List[int]]:
    """model/HSD.py:87-94 (zero-pads to length ``hop`` only, as the reference does)."""
    out = {}
    for node, layers in rings.items():
        sizes = [len(l) for l in layers]
        if len(sizes) < hop:
            sizes = sizes + [0] * (hop - len(sizes))
        out[node] = sizes
    return out


# --------------------------------------------------------------------------
# W1 — scipy.stats._stats_py._cdf_distance(p=1) (scipy 1.18.1 :10098), the
# routine model/HSD.py:111 calls
# --------------------------------------------------------------------------
def w1_restated(u, v) -> float:
    """sum_k |F_u(x_k) - F_v(x_k)| (x_{k+1} - x_k) over the sorted union of values,
    F = right-continuous empirical CDF.  Raises ValueError on an empty input,
    like scipy does."""
    u = np.asarray(u, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    if u.size == 0 or v.size == 0:
        raise ValueError("Distribution can't be empty.")
    us, vs = np.sort(u), np.sort(v)
    allv = np.sort(np.concatenate([u, v]), kind="mergesort")
    deltas = np.diff(allv)
    cu = np.searchsorted(us, allv[:-1], "right") / u.size
    cv = np.searchsorted(vs, allv[:-1], "right") / v.size
    return float(np.sum(np.abs(cu - cv) * deltas))


def w1(u, v) -> float:
    """The reference's call (model/HSD.py:111): scipy when present."""
    if _scipy_w1 is not None:
        if len(u) == 0 or len(v) == 0:
            raise ValueError("Distribution can't be empty.")
        return float(_scipy_w1(u, v))
    return w1_restated(u, v)


# --------------------------------------------------------------------------
# aligned distance — tools/metrics.py:18-36, 117-138, 151-192
# --------------------------------------------------------------------------
def aligned_distance(p: Sequence[float], q: Sequence[float], metric: str = "wasserstein") -> float:
    if not metric or not isinstance(metric, str):
        raise TypeError("Need to specify a metric.")
    metric = metric.lower()
    if metric not in ("wasserstein", "hellinger", "wasserstein_guass"):
        raise NotImplementedError("{} metric is not implemented.".format(metric))
    length = max(len(p), len(q))
    p = np.sort(np.asarray(list(p) + [0.0] * (length - len(p)), dtype=np.float64))
    q = np.sort(np.asarray(list(q) + [0.0] * (length - len(q)), dtype=np.float64))
    if length == 0:
        return 0.0
    if metric == "wasserstein":
        return w1(p, q)
    if metric == "wasserstein_guass":   # tools/metrics.py:54-71
        u1, u2 = np.mean(p), np.mean(q)
        s1, s2 = np.mean(np.square(p - u1)), np.mean(np.square(q - u2))
        return float((u1 - u2) ** 2 + s1 + s2 - 2 * (s1 * s2) ** 0.5)
    bc = 0.0
    for px, qx in zip(p, q):
        if px < 0 or qx < 0:
            continue
        bc += math.sqrt(max(px * qx, 0))
    if math.isclose(bc, 0.0, abs_tol=1e-6):
        bc = 0.0
    elif math.isclose(bc, 1.0, abs_tol=1e-6):
        bc = 1.0
    return math.sqrt(max(1.0 - bc, 0))


# --------------------------------------------------------------------------
# wavelets — model/HSD.py:48-67 (exact) and the pygsp path it calls (approx)
# --------------------------------------------------------------------------
def laplacian_dense(adj: Sequence[np.ndarray]) -> np.ndarray:
    """Combinatorial L = D - A, unit weights (model/HSD.py:34)."""
    n = len(adj)
    L = np.zeros((n, n), dtype=np.float64)
    for i, nb in enumerate(adj):
        nb = nb[nb != i]
        L[i, nb] = -1.0
        L[i, i] = float(len(nb))
    return L


def threshold(w: np.ndarray, n: int, coeff: float = 1e-4) -> np.ndarray:
    """model/HSD.py:65-66: x if x > coeff/n else 0 (GraphWave.py:46 uses 1e-5)."""
    return np.where(w > coeff * 1.0 / n, w, 0.0)


def exact_wavelets(L: np.ndarray, scale: float, thr_coeff: float | None = 1e-4) -> np.ndarray:
    """model/HSD.py:61-66: U diag(exp(-s lambda)) U^T then threshold."""
    lam, U = np.linalg.eigh(L)
    w = np.dot(np.dot(U, np.diag(np.exp(-1 * scale * lam))), np.transpose(U))
    return w if thr_coeff is None else threshold(w, L.shape[0], thr_coeff)


def cheby_coeff(scale: float, lmax: float, order: int) -> np.ndarray:
    """pygsp.filters.approximations.compute_cheby_coeff for Heat(tau = scale*lmax)
    (call site model/HSD.py:52-53): kernel g(x) = exp(-tau x / lmax) = exp(-scale x),
    (order+1)-point Chebyshev-Gauss quadrature on [0, lmax], N = order + 1."""
    N = order + 1
    a1 = a2 = lmax / 2.0
    tau = scale * lmax
    tmpN = np.arange(N)
    num = np.cos(np.pi * (tmpN + 0.5) / N)
    g = np.exp(-tau * (a1 * num + a2) / lmax)
    c = np.empty(order + 1, dtype=np.float64)
    for o in range(order + 1):
        c[o] = 2.0 / N * np.dot(g, np.cos(np.pi * o * (tmpN + 0.5) / N))
    return c


def cheby_apply(L, coeff: np.ndarray, lmax: float, signal: np.ndarray) -> np.ndarray:
    """pygsp.filters.approximations.cheby_op (call site model/HSD.py:58):
    three-term recurrence on the shifted/scaled Laplacian, r = c0/2 T0 + sum c_k T_k.
    ``L`` dense or scipy.sparse; ``signal`` (N,) or (N, C)."""
    a1 = a2 = lmax / 2.0
    t_old = signal
    t_cur = (L @ signal - a2 * signal) / a1
    r = 0.5 * coeff[0] * t_old + coeff[1] * t_cur
    for k in range(2, len(coeff)):
        t_new = (2.0 / a1) * (L @ t_cur - a2 * t_cur) - t_old
        r = r + coeff[k] * t_new
        t_old, t_cur = t_cur, t_new
    return r


def estimate_lmax(L) -> float:
    """pygsp Graph.estimate_lmax: 1.01 * largest eigenvalue (ARPACK, tol 5e-3 in
    pygsp; here the dense exact value — pygsp's own estimate is only reproducible
    to ~2 digits, so both sides of every parity test take lmax as an input)."""
    lam = np.linalg.eigvalsh(np.asarray(L.todense() if hasattr(L, "todense") else L))
    return 1.01 * float(lam[-1])


def cheby_wavelets(L, scale: float, lmax: float, order: int = 50,
                   thr_coeff: float | None = 1e-4, columns: np.ndarray | None = None) -> np.ndarray:
    """model/HSD.py:49-66: response to every impulse, row i = response to impulse i.
    With ``columns`` given returns only those rows (shape (len(columns), N))."""
    n = L.shape[0]
    cols = np.arange(n) if columns is None else np.asarray(columns)
    E = np.zeros((n, len(cols)), dtype=np.float64)
    E[cols, np.arange(len(cols))] = 1.0
    R = cheby_apply(L, cheby_coeff(scale, lmax, order), lmax, E).T
    return R if thr_coeff is None else threshold(R, n, thr_coeff)


# --------------------------------------------------------------------------
# ring signals and distances — model/HSD.py:71-83, 98-114, 140-161
# --------------------------------------------------------------------------
def hierarchical_coefficients(wavelets: np.ndarray, rings: Dict[int, List[List[int]]]):
    """model/HSD.py:71-83: coeffs[i][h] = [Psi[i, j] for j in ring_h(i)]."""
    return {i: [[wavelets[i, j] for j in layer] for layer in layers] for i, layers in rings.items()}


def structural_distance_from_coeffs(coeffs, n: int, hop: int, pairs=None) -> np.ndarray:
    """model/HSD.py:100-114: D[i,j] = D[j,i] = sum_{h=0..hop} W1(coeffs_i[h], coeffs_j[h])."""
    D = np.zeros((n, n), dtype=np.float64)
    it = ((i, j) for i in range(n) for j in range(i + 1, n)) if pairs is None else pairs
    for i, j in it:
        d = 0.0
        for h in range(hop + 1):
            d += w1(coeffs[i][h], coeffs[j][h])
        D[i, j] = D[j, i] = d
    return D


def degree_signal_rows(adj: Sequence[np.ndarray], rings, hop: int):
    """Degree-valued ring signal (north-star formulation; the reference's loop
    model/HSD.py:98-114 with Psi[i, j] replaced by deg(j))."""
    deg = np.array([len(a) for a in adj], dtype=np.float64)
    return {i: [deg[np.asarray(layer, dtype=np.int64)] if len(layer) else np.zeros(0)
                for layer in layers] for i, layers in rings.items()}


def degree_distance_rows(adj: Sequence[np.ndarray], hop: int, rows: Sequence[int],
                         cols: Sequence[int] | None = None, empty: str = "raise") -> np.ndarray:
    """D[r, c] = sum_h W1(deg over ring_h(r), deg over ring_h(c)) for the given rows
    against ``cols`` (default all nodes).  ``empty='zero'`` treats an empty ring as
    the point mass at 0 (tools/metrics.py:18-36 zero padding); 'raise' propagates
    scipy's ValueError like model/HSD.py:111 would."""
    n = len(adj)
    cols = list(range(n)) if cols is None else list(cols)
    need = sorted(set(rows) | set(cols))
    rings = all_rings(adj, hop, need)
    sig = degree_signal_rows(adj, rings, hop)
    if empty == "zero":
        sig = {i: [s if len(s) else np.zeros(1) for s in layers] for i, layers in sig.items()}
    out = np.zeros((len(rows), len(cols)), dtype=np.float64)
    for a, r in enumerate(rows):
        for b, c in enumerate(cols):
            if r == c:
                continue
            d = 0.0
            for h in range(hop + 1):
                d += w1(sig[r][h], sig[c][h])
            out[a, b] = d
    return out


def worker_row(wavelets: np.ndarray, rings, n: int, hop: int, start: int, metric: str) -> np.ndarray:
    """model/HSD.py:140-161 exactly as written: hops 0..hop-1, and BOTH signals are
    read from row ``start`` of the wavelet matrix (the reference indexes q with
    startIndex too, :155)."""
    dists = np.zeros(n)
    layers = rings[start]
    for idx in range(start + 1, n):
        other = rings[idx]
        d = 0.0
        for h in range(hop):
            p = [wavelets[start, j] for j in layers[h]]
            q = [wavelets[start, j] for j in other[h]]
            d += aligned_distance(p, q, metric)
        dists[idx] = d
    return dists


# --------------------------------------------------------------------------
# MultiHSD ring statistics — model/multiscale_HSD.py:45-61, 64-73
# --------------------------------------------------------------------------
def ring_sum_mean(wavelets: np.ndarray, rings, node: int) -> List[float]:
    """get_triple: [sum, mean] per hop, [0, 0] for an empty ring."""
    out: List[float] = []
    for layer in rings[node]:
        vals = [wavelets[node, j] for j in layer]
        out.extend([float(np.sum(vals)), float(np.mean(vals))] if vals else [0.0, 0.0])
    return out


def ring_sum_mean_var(wavelets: np.ndarray, rings, node: int) -> List[float]:
    """The [sum, mean, var] triple robust.csv was written with (the CSV predates
    the current two-value get_triple; tests/robust_test/main.py:190-193)."""
    out: List[float] = []
    for layer in rings[node]:
        vals = [wavelets[node, j] for j in layer]
        out.extend([float(np.sum(vals)), float(np.mean(vals)), float(np.var(vals))] if vals
                   else [0.0, 0.0, 0.0])
    return out


def multiscale_scales(lmax: float, n_scales: int) -> np.ndarray:
    """model/multiscale_HSD.py:30."""
    return np.exp(np.linspace(np.log(0.01), np.log(lmax * 1.25), n_scales))
