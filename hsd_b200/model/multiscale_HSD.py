"""Drop-in for ``model/multiscale_HSD.py`` (class MultiHSD, :15-119).

embed / parallel_embed run the Chebyshev SpMM kernel on blocks of impulse columns
for up to 8 scales at a time and reduce each block straight into the per-node
[sum, mean] ring statistics (hsd_cheb_spmm -> hsd_ring_reduce); the dense N x N
wavelet matrix the reference ships through a multiprocessing pipe per scale
(:79-81) is never materialised.
"""
from __future__ import annotations

from collections import defaultdict

import networkx as nx
import numpy as np
import torch

from .. import engine, wavelets as _wav
from .._lib import check, lib
from .HSD import HSD
from ._device import on_model_device


def _ring_reduce(psiT, n_scales, n, n_cols, rings, sizes, hops, col0, emb):
    """hsd_ring_reduce with a workspace sized for ~1000 CTAs of partial sums."""
    per = n_cols * n_scales * (hops + 1)
    scratch = torch.empty(per * max(1, min(256, (1 << 24) // max(per, 1))), dtype=torch.float64, device=emb.device)
    check(lib.hsd_ring_reduce(engine._ptr(psiT), n_scales, n, n_cols, engine._ptr(rings.bitmaps),
                              engine._ptr(sizes), engine._ptr(rings.orig_of), hops, col0,
                              engine._ptr(emb), engine._ptr(scratch), scratch.numel(), engine._stream()))


class MultiHSD(HSD):

    def __init__(self, graph: nx.Graph, graphName: str, hop: int, n_scales: int, metric="euclidean",
                 signal="wavelet", device=None):
        super(MultiHSD, self).__init__(graph, graphName, 0, hop, metric, signal=signal, device=device)
        self.n_scales = n_scales
        self.scales = None
        self.embeddings = None
        self.init()

    @on_model_device
    def init(self):
        """model/multiscale_HSD.py:26-31: scales from the (estimated) largest Laplacian
        eigenvalue; rings from a .layers file when one is configured, else the BFS kernel."""
        if self.signal == "degree":
            self.scales = np.zeros(1)   # the degree signal has no scale axis
        else:
            if self.lmax is None:
                self.lmax = _wav.estimate_lmax(self.csr, device=self._device())
            self.scales = np.exp(np.linspace(np.log(0.01), np.log(self.lmax * 1.25), self.n_scales))
        super(MultiHSD, self).init()

    # ---- batched device path ----
    @on_model_device
    def embed_device(self, approx=True, stat="triple", col_range=None) -> torch.Tensor:
        """emb[N, n_scales, hop+1, 2] float64 = [sum, mean] of Psi_s[i, ring_h(i)].
        col_range=(begin, end) restricts the work to those nodes' impulse columns (rows of emb
        outside the range stay zero) — the unit of multi-GPU sharding (embed_device_sharded)."""
        n, hops = self.n_node, self.hop
        cbeg, cend = (0, n) if col_range is None else col_range
        rings = self._rings()
        scales = [float(s) for s in self.scales]
        dev = self._device()
        emb = torch.zeros((n, len(scales), hops + 1, 2), dtype=torch.float64, device=dev)
        sizes = rings.sizes.contiguous()
        thr = self.THRESHOLD_COEFF * 1.0 / n
        if approx:
            if self._weighted:
                raise NotImplementedError("the Chebyshev kernel takes unit edge weights; use approx=False")
            csr = self._device_csr()
            if self.lmax is None:
                self.lmax = _wav.estimate_lmax(self.csr, device=self._device())
            for s0 in range(0, len(scales), 8):
                sc = scales[s0:s0 + 8]
                coeffs = np.stack([_wav.cheby_coefficients(s, self.lmax, self.CHEB_ORDER) for s in sc])
                cb = _wav.column_block(n, len(sc))
                work = torch.empty((3, n, cb), dtype=torch.float64, device=dev)
                out = torch.empty((len(sc), n, cb), dtype=torch.float64, device=dev)
                part = torch.zeros((n, len(sc), hops + 1, 2), dtype=torch.float64, device=dev)
                for c0 in range(cbeg, cend, cb):
                    c = min(cb, cend - c0)
                    w = work.reshape(-1)[:3 * n * c].view(3, n, c)
                    o = out.reshape(-1)[:len(sc) * n * c].view(len(sc), n, c)
                    _wav.cheb_wavelet_block(csr, self.lmax, coeffs, c0, c, thr, w, o)
                    _ring_reduce(o, len(sc), n, c, rings, sizes, hops, c0, part)
                emb[:, s0:s0 + len(sc)] = part
        else:
            L = torch.as_tensor(np.asarray(self.L, dtype=np.float64), device=dev)
            eig = torch.linalg.eigh(L)
            if col_range is not None:
                raise ValueError("col_range applies to the Chebyshev path only")
            for si, s in enumerate(scales):
                psi = _wav.exact_wavelets_dense(L, s, self.THRESHOLD_COEFF, eig)
                part = torch.zeros((n, 1, hops + 1, 2), dtype=torch.float64, device=dev)
                psiT = psi.t().contiguous().view(1, n, n)   # psiT[0, v, c] = Psi[c, v]
                _ring_reduce(psiT, 1, n, n, rings, sizes, hops, 0, part)
                emb[:, si:si + 1] = part
        return emb

    @on_model_device
    def embed_device_sharded(self, rank: int, world: int, group=None) -> torch.Tensor:
        """Multi-GPU embed (SURVEY.md §8 e): every rank owns whole Chebyshev recurrences for a
        contiguous block of impulse columns (graph and rings replicated), computes the ring
        statistics of its own nodes, and ONE all-gather of the [N, S, H+1, 2] table replicates
        the result.  The reference's counterpart is Pool-over-scales with a dense N x N matrix
        per scale shipped through a pipe (model/multiscale_HSD.py:79-81)."""
        import torch.distributed as dist
        n = self.n_node
        per = ((n + world - 1) // world + 1) // 2 * 2          # even: the SpMM works on column pairs
        beg, end = min(rank * per, n), min((rank + 1) * per, n)
        emb = self.embed_device(approx=True, col_range=(beg, end)) if end > beg else None
        table = torch.zeros((world * per,) + ((len(self.scales), self.hop + 1, 2)), dtype=torch.float64,
                            device=self._device())
        if emb is not None:
            table[beg:end] = emb[beg:end]
        if world > 1:
            dist.all_gather_into_tensor(table, table[rank * per:(rank + 1) * per].clone(), group=group)
        return table[:n]

    def _emb_to_dict(self, emb: torch.Tensor, stat: str) -> dict:
        e = emb.cpu().numpy()
        flat = e.reshape(self.n_node, -1) if stat == "triple" else e[..., 0].reshape(self.n_node, -1)
        out = defaultdict(list)
        for i, node in enumerate(self.nodes):
            out[node] = flat[i].tolist()
        return out

    # embed nodes into vectors using multi-scale wavelets (model/multiscale_HSD.py:35-42)
    def embed(self, approx=True) -> dict:
        return self._emb_to_dict(self.embed_device(approx=approx), "triple")

    def _ring_stats_of_row(self, wavelets, node):
        """[sum, mean] per hop of one caller-supplied wavelet row, through hsd_ring_reduce."""
        i = self.node2idx[node]
        rings = self._rings()
        dev = self._device()
        row = torch.as_tensor(np.ascontiguousarray(np.asarray(wavelets, dtype=np.float64)[i]), device=dev)
        part = torch.zeros((self.n_node, 1, self.hop + 1, 2), dtype=torch.float64, device=dev)
        _ring_reduce(row.view(1, self.n_node, 1), 1, self.n_node, 1, rings, rings.sizes.contiguous(), self.hop, i, part)
        return part[i, 0].cpu().numpy()

    @on_model_device
    def get_triple(self, wavelets: np.ndarray, node: str) -> list:
        """model/multiscale_HSD.py:45-61: [sum, mean] per hop, [0, 0] for an empty ring."""
        return self._ring_stats_of_row(wavelets, node).reshape(-1).tolist()

    @on_model_device
    def get_layer_sum(self, wavelets: np.ndarray, node: str) -> list:
        """model/multiscale_HSD.py:64-73: ring sums only."""
        return self._ring_stats_of_row(wavelets, node)[:, 0].tolist()

    def parallel_embed(self, n_workers=None) -> dict:
        """model/multiscale_HSD.py:76-98 (n_workers is accepted and ignored)."""
        self.embeddings = self.embed()
        return self.embeddings

    @on_model_device
    def init_scales(self):
        """(Re)compute lmax and the scale grid (model/multiscale_HSD.py:27-30) if the graph changed."""
        if self.signal != "degree" and self.lmax is None:
            self.lmax = _wav.estimate_lmax(self.csr, device=self._device())
            self.scales = np.exp(np.linspace(np.log(0.01), np.log(self.lmax * 1.25), self.n_scales))

    @on_model_device
    def structural_distance_multiscale_device(self) -> torch.Tensor:
        """Device-resident float64 sum over self.scales of the Chebyshev-wavelet HSD matrix."""
        total = None
        for scale in self.scales:
            D = HSD.structural_distance_device(self, float(scale), approx=True)
            total = D.to(torch.float64) if total is None else total + D
        return total

    def parallel_calculate_structural_distance(self, n_workers: int = None):
        """model/multiscale_HSD.py:101-119: sum over scales of
        HSD.calculate_structural_distance(scale, approx=True)."""
        with torch.cuda.device(self._device()):
            return self.structural_distance_multiscale_device().cpu().numpy()
