"""Drop-in for ``model/dynamic_HSD.py`` (class DynamicHSD, :10-55).

The reference's update method is an empty stub (``dynamic_add_node`` is ``pass``,
:23-24), so the behaviour is defined by BASELINE.json config 5: after edge /
node insertions, recompute only the rows (and mirrored columns) of the distance
matrix whose signatures can have changed; the result must equal a from-scratch
recompute on the updated graph.

Affected set (degree signal): a node's k-hop rings change only if it lies within
hop-1 of an endpoint of a new edge, and the degree of a ring member changes only
for the endpoints themselves, so every changed signature belongs to a node within
``hop`` hops (in the new graph) of an endpoint (``affected_nodes_device``: the OR of
the BFS kernel's ring bitmaps for the endpoints).  The update itself uses the exact
set instead: all signatures are rebuilt (the BFS kernel costs ~1 % of the pairwise
kernel) and compared bit for bit with the previous table; only rows whose signature
differs are recomputed.  When the set of distinct degrees (the shared support)
changes, every signature changes representation and the update is a full recompute.
"""
from __future__ import annotations

from typing import Iterable

import networkx as nx
import numpy as np
import torch

from .. import engine
from ..graph import CSRGraph
from ..tools import util
from .multiscale_HSD import MultiHSD
from ._device import on_model_device


class _UpdateWorkspace:
    """Grow-only device buffers of the incremental update, so that a timed update allocates nothing:
    two signature tables (current / previous, swapped per update), the K-major table
    [all nodes | one chunk of affected nodes] and the chunk's result block.  The signature length moves with
    the set of distinct degrees, so the buffers carry slack for 16 more distinct degrees per hop (a fresh
    cudaMalloc of one 400 MB table costs 30-85 ms on a B200 — more than the ring kernel it feeds;
    scripts/time_c5_update.py).  Affected rows are recomputed `rows` at a time: bounded memory whatever |A|."""

    def __init__(self, n: int, dev, hops: int, rows: int, k_used: int = 0):
        self.n, self.dev, self.hops, self.rows = n, dev, hops, rows
        self.n4 = engine.roundup(n, 4)
        self._sig = [None, None]
        self._flip = 0
        self._sigT = None
        self.block = torch.empty((rows, self.n4), dtype=torch.float32, device=dev)
        if k_used > 0:
            # everything up front, before the N x N matrix exists: a buffer allocated later would be carved out
            # of a cached free block — after `model._D = None` that is the matrix's own 40 GB block, and the next
            # matrix then costs a second 40 GB cudaMalloc (measured: 84.7 GB reserved instead of 44.7)
            self.next_signature_table(n, k_used)
            self.next_signature_table(n, k_used)
            self.k_major_table(k_used)

    def _capacity_k(self, k_used: int) -> int:
        return engine.roundup(k_used + 16 * max(self.hops, 1), engine.PAIR_KCHUNK)

    def next_signature_table(self, n: int, k_used: int) -> torch.Tensor:
        """float32 [n, roundup(k_used, 4)] view of the buffer that does NOT hold the previous table."""
        self._flip ^= 1
        ld = engine.roundup(k_used, 4)
        buf = self._sig[self._flip]
        if buf is None or buf.numel() < n * ld:
            buf = self._sig[self._flip] = torch.empty(n * self._capacity_k(k_used), dtype=torch.float32, device=self.dev)
        return buf[:n * ld].view(n, ld)

    def k_major_table(self, k_used: int) -> torch.Tensor:
        """Zeroed float32 [roundup(k_used, 16), n4 + rows] view (the pairwise kernel's TMA source)."""
        k_pad = engine.roundup(max(k_used, 1), engine.PAIR_KCHUNK)
        cols = self.n4 + self.rows
        if self._sigT is None or self._sigT.numel() < k_pad * cols:
            self._sigT = torch.empty(self._capacity_k(k_used) * cols, dtype=torch.float32, device=self.dev)
        t = self._sigT[:k_pad * cols].view(k_pad, cols)
        t.zero_()
        return t


class DynamicHSD(MultiHSD):

    UPDATE_ROWS = 4096       # affected rows recomputed per pairwise launch (bounds the update's workspace)

    def __init__(self, graph: nx.Graph, graphName: str, hop: int, n_scales: int, metric="euclidean",
                 signal="wavelet", device=None):
        super(DynamicHSD, self).__init__(graph, graphName, hop, n_scales, metric, signal=signal, device=device)
        self.embeddings = {}
        self._D = None           # device-resident distance matrix kept across updates
        self._pending = set()    # endpoints (node labels) of edges inserted since the last update
        self._sig_prev = None    # signature table / support of the matrix in self._D
        self._support_prev = None
        self.last_affected = None

    def init(self):
        super(DynamicHSD, self).init()

    def _device_graph(self, include_zero=False) -> engine.DeviceGraph:
        # the graph changes between steps: order it by degree on the device instead of on the host
        if self._dg is None or self._dg.include_zero != include_zero:
            self._dg = engine.DeviceGraph.upload_device_order(self.csr, include_zero=include_zero,
                                                              device=self._device())
        return self._dg

    # ---- graph edits ----
    def _refresh_graph(self, new_edges=None):
        """new_edges: the edges just inserted when the node set did not change — the CSR is then
        edited in place of a rebuild from networkx (1.8 s -> ~10 ms at 100k nodes)."""
        if new_edges is not None and self.n_node == self.graph.number_of_nodes():
            idx = np.array([(self.node2idx[u], self.node2idx[v]) for u, v in new_edges], dtype=np.int64).reshape(-1, 2)
            self.csr = self.csr.with_edges_added(idx)
        else:
            self.nodes = list(nx.nodes(self.graph))
            self.n_node = len(self.nodes)
            self.idx2node, self.node2idx = util.build_node_idx_map(self.graph)
            self.csr = CSRGraph.from_networkx(self.graph)
        self._A = self._L = None
        self._dg = self._dcsr = None
        self._ringset = self._ringset_src = None
        self._hierarchy = None
        self._hierarchy_lazy = True
        self.lmax = None

    def dynamic_add_edges(self, edges: Iterable):
        """Insert edges (pairs of node labels; unknown labels become new nodes)."""
        edges = [(u, v) for u, v in edges]
        n_before = self.graph.number_of_nodes()
        self.graph.add_edges_from(edges)
        for u, v in edges:
            self._pending.update((u, v))
        if self.graph.number_of_nodes() != n_before:
            self._D = None   # the matrix changes shape: next update is a full recompute
            self._refresh_graph()
        else:
            self._refresh_graph(new_edges=edges)

    def dynamic_add_node(self, newNode: str, edges: list):
        """model/dynamic_HSD.py:23-24 (a stub there): add ``newNode`` with the given edges."""
        self.graph.add_node(newNode)
        self.dynamic_add_edges([(e[0], e[1]) for e in edges])
        if newNode not in self._pending:
            self._pending.add(newNode)
            self._D = None
            self._refresh_graph()

    # ---- incremental distance (degree signal) ----
    @on_model_device
    def affected_nodes_device(self) -> torch.Tensor:
        """Original indices of nodes within ``hop`` hops of a pending endpoint (int64, sorted)."""
        dg = self._device_graph(include_zero=(self.empty == "zero"))
        ends = torch.tensor(sorted(self.node2idx[v] for v in self._pending), dtype=torch.int32, device=dg.rowptr.device)
        if ends.numel() == 0:
            return torch.zeros(0, dtype=torch.int64, device=dg.rowptr.device)
        ball = None
        for e0 in range(0, ends.numel(), 4096):   # bound the bitmap scratch: 4096 x (hop+1) x N/8 bytes
            _, _, bm, _ = engine.ring_signature_degree(dg, self.hop, rows=ends[e0:e0 + 4096], want_sig=False,
                                                       want_sizes=False, want_bitmaps=True)
            part = bm.view(-1, bm.shape[-1])
            red = part[0].clone()
            for chunk in torch.split(part[1:], 1 << 16):
                if chunk.numel():
                    red |= _or_reduce(chunk)
            ball = red if ball is None else (ball | red)
        bits = _unpack_bits(ball, dg.n)
        return torch.sort(dg.orig_of[bits].to(torch.int64)).values

    @on_model_device
    def structural_distance_update(self) -> torch.Tensor:
        """Distance matrix of the current graph (device, float32).  After insertions only the
        affected rows / columns are recomputed; equals a from-scratch recompute bit for bit
        because every entry is produced by the same kernel from the same signatures."""
        if self.signal != "degree":
            # wavelet signal: MultiHSD's distance is the sum over self.scales
            # (model/multiscale_HSD.py:101-119); self.scale is 0 for a MultiHSD (exp(-0 L) = I would
            # give an all-zero matrix).  Every Psi_s changes globally with the graph: full recompute.
            self.init_scales()
            self._D = self.structural_distance_multiscale_device()
            self._pending.clear()
            return self._D
        dg = self._device_graph(include_zero=(self.empty == "zero"))
        n, hops = dg.n, self.hop
        k_used = dg.k_used(hops)
        ws = self._update_workspace(n, k_used, dg.rowptr.device)
        sig = ws.next_signature_table(n, k_used)
        _, _, _, status = engine.ring_signature_degree(dg, hops, empty=self.empty, want_sizes=False, sig_out=sig)
        if self.empty == "raise" and int(status.item()) & 1:
            raise engine.EmptyRingError("Distribution can't be empty.")
        prev, prev_support = self._sig_prev, self._support_prev
        self._sig_prev, self._support_prev = sig, dg.support
        same_layout = (self._D is not None and self._D.shape[0] == n and prev is not None
                       and prev.shape == sig.shape and np.array_equal(prev_support, dg.support))
        n4 = engine.roundup(n, 4)
        sigT = ws.k_major_table(k_used)                 # [all nodes | one chunk of affected nodes]
        if same_layout:
            aff = torch.nonzero((sig != prev).any(dim=1), as_tuple=False).reshape(-1)   # exact changed set
            m = int(aff.numel())
        else:
            aff, m = torch.arange(n, device=sig.device), n
        self.last_affected = aff
        if m * 2 >= n:   # (also: new layout)  rectangular |A| x N costs more than the symmetric full matrix
            engine.signature_transpose(sig, k_used, sigT, 0)
            self._D = engine.pairwise_l1(sigT, n, symmetric=True, k_used=k_used,
                                         out=self._D if self._D is not None and self._D.shape[0] == n else None)
        elif m > 0:
            # rows = a chunk of the affected nodes (appended to the table behind all nodes), columns = all nodes;
            # hsd_scatter_symmetric writes the chunk's rows and mirrored columns into the resident matrix
            engine.signature_transpose(sig, k_used, sigT, 0)
            rows = ws.rows
            aff32 = aff.to(torch.int32)
            for c0 in range(0, m, rows):
                idx = aff[c0:c0 + rows]
                mq = int(idx.numel())
                engine.signature_transpose(sig, k_used, sigT, n4, src_rows=aff32[c0:c0 + rows].contiguous())
                blk = engine.pairwise_l1(sigT, n4 + rows, row0=n4, n_rows=mq, col0=0, n_cols=n, symmetric=False,
                                         k_used=k_used, out=ws.block[:mq, :n])
                engine.scatter_symmetric(blk, idx.contiguous(), self._D)
        self._pending.clear()
        return self._D

    def _update_workspace(self, n: int, k_used: int, dev) -> "_UpdateWorkspace":
        ws = getattr(self, "_ws", None)
        if ws is None or ws.n != n or ws.dev != dev:
            ws = self._ws = _UpdateWorkspace(n, dev, self.hop, min(self.UPDATE_ROWS, engine.roundup(max(n // 2, 4), 4)),
                                             k_used)
        return ws

    @on_model_device
    def structural_distance_update_sharded(self, rank: int, world: int, group=None, peer: bool = True) -> torch.Tensor:
        """structural_distance_update() on `world` GPUs (one process per GPU, every process applies
        the same insertions): returns this rank's row block [shard_rows(n, world, rank)] of the
        matrix, float32 on its device.  The plan (signature table, result block, peer mappings) is
        kept across updates; ShardedDegreeHSD.update deals the affected rows round-robin and stores
        them through peer memory.  A new node or a new distinct degree rebuilds the plan (full step)."""
        if self.signal != "degree":
            raise NotImplementedError("the sharded incremental update is defined for signal='degree'")
        from ..sharded import ShardedDegreeHSD
        dg = self._device_graph(include_zero=(self.empty == "zero"))
        plan = getattr(self, "_plan", None)
        reusable = (plan is not None and plan.world == world and plan.rank == rank and plan.dg.n == dg.n
                    and plan.hops == self.hop and np.array_equal(plan.dg.support, dg.support))
        if not reusable:
            # a changed support (a new distinct degree) changes the signature length, not the result block:
            # the new plan takes over the old one's symmetric-memory allocations where they still fit
            self._plan = plan = ShardedDegreeHSD(dg, self.hop, rank, world, group=group, empty=self.empty, peer=peer,
                                                 reuse=plan)
            blk = plan.step()
            self.last_affected = torch.arange(dg.n, device=dg.rowptr.device)
        else:
            blk, self.last_affected = plan.update(dg)
        plan.check()
        self._pending.clear()
        return blk

    # ---- exploratory helpers of the reference ----
    @on_model_device
    def explore_neighborhoods(self, node, maxHop=5) -> set:
        """model/dynamic_HSD.py:28-43: BFS layers of one node, stored in ``self.hierarchy[node]``;
        returns every node within ``self.hop`` hops.  (The reference discards the result of
        ``neighborhoods.union`` at :40, so its later layers are not deduplicated and it returns
        {node}; this implements the evident intent and matches tools/hierarchy.py:25-38.)"""
        dg = self._device_graph()
        src = torch.tensor([self.node2idx[node]], dtype=torch.int32, device=dg.rowptr.device)
        _, _, bm, _ = engine.ring_signature_degree(dg, self.hop, rows=src, want_sig=False, want_bitmaps=True)
        bmh = bm.cpu().numpy().view(np.uint32)
        bits = np.unpackbits(bmh.view(np.uint8), axis=-1, bitorder="little")[0, :, :dg.n]
        orig = dg.orig_of.cpu().numpy()
        layers = [[self.nodes[j] for j in np.sort(orig[np.nonzero(bits[h])[0]])] for h in range(self.hop + 1)]
        hier = self.hierarchy
        if hier is None:
            hier = {}
        hier[node] = layers
        self._hierarchy = hier
        return set(v for layer in layers for v in layer)

    def convert_neighborhoods_to_subgraph(self, neighbors) -> nx.Graph:
        """model/dynamic_HSD.py:46-55: induced subgraph on ``neighbors`` (edges only)."""
        neighbors = set(neighbors)
        sub = nx.Graph()
        sub.add_edges_from((u, v) for u, v in nx.edges(self.graph, neighbors) if u in neighbors and v in neighbors)
        return sub


def _or_reduce(rows: torch.Tensor) -> torch.Tensor:
    """Bitwise OR over dim 0 of an int32 [m, words] tensor (tree reduction with torch ops)."""
    while rows.shape[0] > 1:
        m = rows.shape[0]
        half = m // 2
        merged = rows[:half] | rows[half:2 * half]
        rows = torch.cat([merged, rows[2 * half:]]) if m % 2 else merged
    return rows[0]


def _unpack_bits(words: torch.Tensor, n: int) -> torch.Tensor:
    """Indices of set bits (bit ids < n) of an int32 bitmap."""
    shifts = torch.arange(32, device=words.device, dtype=torch.int32)
    bits = ((words[:, None] >> shifts[None, :]) & 1).reshape(-1)[:n]
    return torch.nonzero(bits, as_tuple=False).reshape(-1)
