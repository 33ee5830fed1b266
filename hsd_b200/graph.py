"""Graph ingest: networkx / edge arrays -> CSR, plus the degree-ordered view the
BFS + degree-CDF kernel needs.

Reference counterpart: HSD.__init__ (model/HSD.py:28-40) builds dense N x N
``A`` and ``L`` and the node <-> index maps; node index = order of first
appearance (``list(nx.nodes(graph))``, tools/util.py:11-24).  Dense N x N is
80 GB at N = 100k, so the adjacency lives as int32 CSR here and ``A`` / ``L``
are materialised lazily by the model classes only when a caller asks.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Hashable, List, Optional, Sequence

import numpy as np


@dataclass
class DegreeOrder:
    """CSR relabelled in degree-ascending order (ties by original index)."""
    orig_of: np.ndarray      # int32[N]  new id -> original index
    new_of: np.ndarray       # int32[N]  original index -> new id
    rowptr: np.ndarray       # int32[N+1]
    col: np.ndarray          # int32[nnz padded to a multiple of 4 (+4)]: LDG.128 reads in the BFS kernel
    sorted_degree: np.ndarray  # int32[N] degree of new id

    def support(self, include_zero: bool = False):
        """Shared support of all ring degree distributions.

        Returns (support float64[B], bin_end int32[B], delta float32[B-1]) with
        bin_end[b] = number of nodes whose degree <= support[b]."""
        sup = np.unique(self.sorted_degree)
        if include_zero and (sup.size == 0 or sup[0] != 0):
            sup = np.concatenate([[0], sup])
        bin_end = np.searchsorted(self.sorted_degree, sup, side="right").astype(np.int32)
        delta = np.diff(sup).astype(np.float32)
        return sup.astype(np.float64), bin_end, delta


@dataclass
class CSRGraph:
    n: int
    rowptr: np.ndarray   # int32[N+1]
    col: np.ndarray      # int32[nnz], sorted within a row
    nodes: List[Hashable] = field(default_factory=list)
    _order: Optional[DegreeOrder] = None

    @property
    def nnz(self) -> int:
        return int(self.col.shape[0])

    @property
    def degree(self) -> np.ndarray:
        return np.diff(self.rowptr).astype(np.int32)

    @classmethod
    def from_edges(cls, n: int, edges: np.ndarray, nodes: Sequence[Hashable] | None = None) -> "CSRGraph":
        """Undirected simple graph from an (E, 2) index array (each edge once, any
        orientation; duplicates and both orientations are tolerated)."""
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        if edges.size and (edges.min() < 0 or edges.max() >= n):
            raise ValueError("edge endpoint out of range")
        u = np.concatenate([edges[:, 0], edges[:, 1]])
        v = np.concatenate([edges[:, 1], edges[:, 0]])
        key = np.sort(u * n + v)            # sort + adjacent compare: np.unique is 10x slower on 1e6 int64 keys
        if key.size:                        # dedupe (also collapses a self-loop's two copies)
            key = key[np.concatenate([[True], key[1:] != key[:-1]])]
        rows = (key // n).astype(np.int64)
        cols = (key % n).astype(np.int32)   # sorted by (row, col) already
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(np.bincount(rows, minlength=n), out=rowptr[1:])
        if rowptr[-1] >= 2**31:
            raise ValueError("graph too large for int32 CSR")
        return cls(n=n, rowptr=rowptr.astype(np.int32), col=cols,
                   nodes=list(nodes) if nodes is not None else list(range(n)))

    @classmethod
    def from_networkx(cls, graph) -> "CSRGraph":
        nodes = list(graph.nodes())
        idx = {v: i for i, v in enumerate(nodes)}
        n = len(nodes)
        if graph.is_directed() or graph.is_multigraph():
            e = np.fromiter((idx[x] for uv in graph.edges() for x in uv), dtype=np.int64)
            return cls.from_edges(n, e.reshape(-1, 2), nodes)
        # undirected simple graph: the adjacency dict already holds both orientations of every edge
        # (a self-loop once), so one flat pass over it replaces the per-edge tuple iteration
        # (0.58 s -> 0.3 s at 100k nodes / 500k edges); columns are then sorted within each row
        from itertools import chain
        adj = getattr(graph, "_adj", None) or graph.adj      # the raw dict-of-dicts: the AtlasView wrappers cost 0.4 s at 100k nodes
        deg = np.fromiter((len(adj[v]) for v in nodes), dtype=np.int64, count=n)
        nnz = int(deg.sum())
        if nnz >= 2**31:
            raise ValueError("graph too large for int32 CSR")
        flat = chain.from_iterable(adj[v] for v in nodes)
        if n and type(nodes[0]) is int and type(nodes[-1]) is int and nodes == list(range(n)):
            cols = np.fromiter(flat, dtype=np.int64, count=nnz)     # labels are the indices already
        else:
            cols = np.fromiter(map(idx.__getitem__, flat), dtype=np.int64, count=nnz)
        key = np.sort(np.repeat(np.arange(n, dtype=np.int64), deg) * n + cols)
        rowptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(deg, out=rowptr[1:])
        return cls(n=n, rowptr=rowptr.astype(np.int32), col=(key % max(n, 1)).astype(np.int32), nodes=nodes)

    def with_edges_added(self, new_edges: np.ndarray) -> "CSRGraph":
        """A new CSRGraph with extra undirected edges (index pairs; edges already present are
        ignored).  The CSR is sorted by (row, col) already, so the 2k new directed entries are merged
        in by binary search + one np.insert: ~9 ms at 1e6 entries instead of a rebuild."""
        n = self.n
        e = np.asarray(new_edges, dtype=np.int64).reshape(-1, 2)
        if e.size and (e.min() < 0 or e.max() >= n):
            raise ValueError("edge endpoint out of range")
        add = np.unique(np.concatenate([e[:, 0] * n + e[:, 1], e[:, 1] * n + e[:, 0]]))
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(self.rowptr))
        key = rows * n + self.col
        pos = np.searchsorted(key, add)
        present = (pos < key.size) & (key[np.minimum(pos, max(key.size - 1, 0))] == add) if key.size else np.zeros(add.size, bool)
        add, pos = add[~present], pos[~present]
        col = np.insert(self.col, pos, (add % n).astype(np.int32))
        rowptr = self.rowptr.astype(np.int64)
        rowptr[1:] += np.cumsum(np.bincount(add // n, minlength=n))
        if rowptr[-1] >= 2**31:
            raise ValueError("graph too large for int32 CSR")
        return CSRGraph(n=n, rowptr=rowptr.astype(np.int32), col=col, nodes=self.nodes)

    def degree_order(self) -> DegreeOrder:
        if self._order is None:
            deg = self.degree
            orig_of = np.argsort(deg, kind="stable").astype(np.int32)
            new_of = np.empty(self.n, dtype=np.int32)
            new_of[orig_of] = np.arange(self.n, dtype=np.int32)
            rows = np.repeat(np.arange(self.n, dtype=np.int64), deg)
            key = np.sort(new_of[rows].astype(np.int64) * self.n + new_of[self.col])
            sdeg = deg[orig_of]
            rowptr = np.zeros(self.n + 1, dtype=np.int32)
            np.cumsum(sdeg, out=rowptr[1:])
            col = (key % self.n).astype(np.int32)
            # the BFS kernel reads column indices in aligned groups of 4 (LDG.128): pad the tail
            col = np.concatenate([col, np.zeros((-len(col)) % 4 + 4, dtype=np.int32)])
            self._order = DegreeOrder(
                orig_of=orig_of, new_of=new_of, rowptr=rowptr,
                col=col,
                sorted_degree=sdeg.astype(np.int32))
        return self._order

    def neighbors(self, i: int) -> np.ndarray:
        return self.col[self.rowptr[i]:self.rowptr[i + 1]]


def has_nonunit_weights(graph) -> bool:
    """True if some edge of a networkx graph carries a 'weight' other than 1.  Edges without any
    attribute (every edge list the reference ships) are skipped with one truthiness test of their
    attribute dict, so the scan costs ~0.05 s at 1e6 adjacency entries instead of 0.5 s."""
    adj = getattr(graph, "_adj", None) or graph.adj
    return any(d.get("weight", 1.0) != 1.0 for nbrs in adj.values() for d in nbrs.values() if d)


def powerlaw_graph(n: int, m: int = 5, seed: int = 0) -> CSRGraph:
    """The synthetic inputs BASELINE.json names: networkx.barabasi_albert_graph(n, m, seed)."""
    import networkx as nx
    g = nx.barabasi_albert_graph(n, m, seed=seed)
    e = np.array(g.edges(), dtype=np.int64)
    return CSRGraph.from_edges(n, e, list(range(n)))
