"""Drop-in for ``model/GraphWave.py`` (class GraphWave, :10-69) — the component
adjacent to the hot path (SURVEY.md §8 f rank 1).  Wavelets come from the same
device paths as HSD (Chebyshev SpMM kernel, order 40, threshold 1e-5/N here);
the characteristic-function sampling is the fused sincos reduction kernel
hsd_characteristic_function."""
from __future__ import annotations

import networkx as nx
import numpy as np
import torch

from .. import engine, wavelets as _wav
from ..graph import CSRGraph
from ..tools import util
from ._device import on_model_device


class GraphWave(object):

    CHEB_ORDER = 40          # model/GraphWave.py:34
    THRESHOLD_COEFF = 1e-5   # model/GraphWave.py:46

    def __init__(self, graph: nx.Graph, device=None):
        self.graph = graph
        self.nodes = list(nx.nodes(graph))
        self.idx2node, self.node2idx = util.build_node_idx_map(graph)
        self.csr = CSRGraph.from_networkx(graph)
        self.device = torch.device(device) if device is not None else engine.require_cuda()
        self.adjacent = nx.adjacency_matrix(graph).todense()
        self.laplacian = nx.laplacian_matrix(graph).todense()
        L = torch.as_tensor(np.asarray(self.laplacian, dtype=np.float64), device=self.device)
        lam, U = torch.linalg.eigh(L)      # eager, like model/GraphWave.py:23
        self._L, self._eig = L, (lam, U)
        self.eigenvalues, self.eigenvectors = lam.cpu().numpy(), U.cpu().numpy()
        self.wavelets = None
        self.lmax = None

    def _device(self):
        return self.device

    @on_model_device
    def calculate_wavelets(self, scale, approx=True) -> np.ndarray:
        if approx:
            if self.lmax is None:
                self.lmax = _wav.estimate_lmax(self.csr, device=self.device)
            psi = _wav.cheb_wavelets_dense(_wav.DeviceCSR(self.csr, self.device), float(scale), self.lmax,
                                           self.CHEB_ORDER, self.THRESHOLD_COEFF)
        else:
            assert getattr(self, "eigenvalues", None) is not None, "GraphWave eigenvalues is None!"
            psi = _wav.exact_wavelets_dense(self._L, float(scale), self.THRESHOLD_COEFF, self._eig)
        self._psi = psi
        self.wavelets = psi.cpu().numpy()
        return self.wavelets

    @on_model_device
    def calculate_characteristic_value(self, X: np.ndarray, sample_points):
        """model/GraphWave.py:53-59: [Re, Im] of mean(exp(i t X)) per sample point."""
        x = torch.as_tensor(np.asarray(X, dtype=np.float64), device=self.device).reshape(1, -1)
        return self._characteristic(x, sample_points)[0].cpu().numpy()

    def _characteristic(self, rows: torch.Tensor, sample_points) -> torch.Tensor:
        from .._lib import check, lib
        t = torch.as_tensor(np.asarray(sample_points, dtype=np.float64), device=self.device).contiguous()
        rows = rows.contiguous()
        out = torch.empty((rows.shape[0], 2 * t.numel()), dtype=torch.float64, device=self.device)
        for r0 in range(0, rows.shape[0], 32768):
            blk = rows[r0:r0 + 32768]
            check(lib.hsd_characteristic_function(engine._ptr(blk), blk.stride(0), blk.shape[0], blk.shape[1],
                                                  engine._ptr(t), t.numel(), engine._ptr(out[r0:r0 + 32768]),
                                                  engine._stream()))
        return out

    @on_model_device
    def embed(self, sample_points):
        assert self.wavelets is not None, "GraphWave wavelets is None!"
        psi = torch.as_tensor(np.asarray(self.wavelets, dtype=np.float64), device=self.device)
        emb = self._characteristic(psi, sample_points).cpu().numpy()
        return {node: emb[i] for i, node in enumerate(self.nodes)}


def recommend_scale_range(eignvalues):
    return util.recommend_scale_range(list(eignvalues))


def scale_boundary(e1, eN, eta=0.85, gamma=0.95):
    return util.scale_boundary(e1, eN, eta, gamma)
