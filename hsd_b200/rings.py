"""Ring sets ("hierarchy" in the reference) on the device.

The reference keeps ``hierarchy[node] = [[node], ring_1, ..., ring_H]`` as Python
lists of node labels (tools/hierarchy.py:16-38).  Here a ring set is a bitmap
tensor int32[N, H+1, ceil(N/32)] plus the ring sizes; the dict form is produced
only when a caller asks for it (and accepted when a caller assigns one, as
tests/robust_test/main.py:179 does).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Hashable, List, Optional, Sequence

import numpy as np
import torch

from . import engine


# metrics of tools/metrics.py::calculate_distance available on the GPU (the other names the
# reference lists — l1, l2, kl, symmetric_kl, js — raise ValueError there on un-normalised ring
# signals, tools/metrics.py:39-51, and are not part of the HSD path)
METRIC_IDS = {"wasserstein": 0, "hellinger": 1, "wasserstein_guass": 2}


@dataclass
class RingSet:
    bitmaps: torch.Tensor            # int32[N, H+1, words]; row = source's original index
    sizes: torch.Tensor              # int32[N, H+1]
    hops: int
    n: int
    orig_of: Optional[torch.Tensor]  # int32[N]: bit id -> original index (None: bits are original ids)
    bit_of: Optional[torch.Tensor] = None  # int32[N]: original index -> bit id (None: identity)

    @classmethod
    def bfs(cls, dg: engine.DeviceGraph, hops: int) -> "RingSet":
        """k-hop rings of every node by the BFS kernel (tools/hierarchy.py:16-38)."""
        _, sizes, bitmaps, _ = engine.ring_signature_degree(dg, hops, want_sig=False,
                                                            want_bitmaps=True)
        return cls(bitmaps=bitmaps, sizes=sizes, hops=hops, n=dg.n, orig_of=dg.orig_of, bit_of=dg.new_of)

    @classmethod
    def from_hierarchy(cls, hierarchy: Dict[Hashable, Sequence[Sequence[Hashable]]],
                       nodes: Sequence[Hashable], node2idx: Dict[Hashable, int], hops: int,
                       device) -> "RingSet":
        """Ingest a reference-style dict.  Layers beyond ``hops`` are ignored, missing
        layers are empty, and the '' sentinel the file reader produces
        (tools/hierarchy.py:87-94) is skipped like model/multiscale_HSD.py:52 does."""
        n = len(nodes)
        words = (n + 31) // 32
        bits = np.zeros((n, hops + 1, words * 32), dtype=np.uint8)
        sizes = np.zeros((n, hops + 1), dtype=np.int32)
        for i, v in enumerate(nodes):
            layers = hierarchy[v]
            for h in range(min(hops + 1, len(layers))):
                members = [node2idx[w] for w in layers[h] if not (isinstance(w, str) and w == "")]
                if members:
                    bits[i, h, members] = 1
                sizes[i, h] = len(set(members))
        packed = np.packbits(bits, axis=-1, bitorder="little").view(np.uint32).view(np.int32)
        return cls(bitmaps=torch.from_numpy(np.ascontiguousarray(packed)).to(device),
                   sizes=torch.from_numpy(sizes).to(device), hops=hops, n=n, orig_of=None)

    def members_host(self) -> List[List[np.ndarray]]:
        """rings[i][h] = sorted original indices (host)."""
        bm = self.bitmaps.cpu().numpy().view(np.uint32)
        bits = np.unpackbits(bm.view(np.uint8), axis=-1, bitorder="little")[..., :self.n]
        orig = None if self.orig_of is None else self.orig_of.cpu().numpy()
        out = []
        for i in range(self.n):
            layers = []
            for h in range(self.hops + 1):
                idx = np.nonzero(bits[i, h])[0]
                layers.append(np.sort(orig[idx]) if orig is not None else idx)
            out.append(layers)
        return out

    def to_hierarchy(self, nodes: Sequence[Hashable]) -> Dict[Hashable, List[List[Hashable]]]:
        """The reference's dict form (empty rings kept as [], tools/hierarchy.py:37)."""
        mem = self.members_host()
        return {nodes[i]: [[nodes[j] for j in layer] for layer in mem[i]] for i in range(self.n)}


# --------------------------------------------------------------------------
# value-mode distances (wavelet-valued ring signals)
# --------------------------------------------------------------------------
def sorted_ring_values(psi: torch.Tensor, rings: RingSet):
    """K2v: gather Psi[i, ring_h(i)] ascending (model/HSD.py:71-83 + scipy's argsort).
    Returns (vals float64[total], offsets int64[N*(H+1)+1])."""
    from ._lib import check, lib
    n, hops = rings.n, rings.hops
    if psi.dtype != torch.float64 or psi.shape != (n, n):
        raise ValueError("psi must be a float64 (N, N) CUDA tensor")
    sizes = rings.sizes.contiguous()
    offsets = torch.zeros(n * (hops + 1) + 1, dtype=torch.int64, device=psi.device)
    torch.cumsum(sizes.reshape(-1).to(torch.int64), 0, out=offsets[1:])
    total = int(offsets[-1].item())
    max_ring = int(sizes.max().item())
    vals = torch.empty(max(total, 1), dtype=torch.float64, device=psi.device)
    check(lib.hsd_ring_signature_values(
        engine._ptr(psi), psi.stride(0), engine._ptr(rings.bitmaps), engine._ptr(sizes),
        engine._ptr(offsets), engine._ptr(rings.orig_of), n, hops, n, max_ring,
        engine._ptr(vals), engine._stream()))
    return vals, offsets


def value_distance(psi: torch.Tensor, rings: RingSet, hop_begin: int = 0,
                   hop_end: Optional[int] = None, mode: str = "w1",
                   metric: str = "wasserstein") -> torch.Tensor:
    """D[i, j] = sum_{h in [hop_begin, hop_end)} dist(Psi[i, ring_h(i)], Psi[j, ring_h(j)]), float64.

    mode 'w1'      exact ragged Wasserstein-1 (model/HSD.py:98-114);
    mode 'aligned' zero-pad + sort (tools/metrics.py:151-192), metric wasserstein | hellinger."""
    from ._lib import check, lib
    n, hops = rings.n, rings.hops
    hop_end = hops + 1 if hop_end is None else hop_end
    vals, offsets = sorted_ring_values(psi, rings)
    sizes = rings.sizes.contiguous()
    D = torch.zeros((n, n), dtype=torch.float64, device=psi.device)
    if mode == "w1":
        status = torch.zeros(1, dtype=torch.int32, device=psi.device)
        for r0 in range(0, n, 32768):
            check(lib.hsd_pairwise_w1_merge(engine._ptr(vals), engine._ptr(offsets), engine._ptr(sizes),
                                            n, hops, hop_begin, hop_end, r0, min(32768, n - r0),
                                            engine._ptr(D), D.stride(0), engine._ptr(status),
                                            engine._stream()))
        if int(status.item()) & 1:
            raise engine.EmptyRingError("Distribution can't be empty.")
    elif mode == "aligned":
        m = METRIC_IDS.get(str(metric).lower())
        if m is None:
            raise NotImplementedError("{} metric is not implemented.".format(metric))
        for r0 in range(0, n, 32768):
            check(lib.hsd_pairwise_aligned(engine._ptr(vals), engine._ptr(offsets), engine._ptr(sizes),
                                           n, hops, hop_begin, hop_end, m, r0, min(32768, n - r0),
                                           engine._ptr(D), D.stride(0), engine._stream()))
    else:
        raise ValueError("mode must be 'w1' or 'aligned'")
    return D


def worker_distance(psi: torch.Tensor, rings: RingSet, hop_end: int, metric: str = "wasserstein") -> torch.Tensor:
    """The reference's row-parallel variant exactly as written (model/HSD.py:118-161):
    D[i, j] = D[j, i] = sum_{h < hop_end} aligned(Psi[i, ring_h(i)], Psi[i, ring_h(j)]), i < j."""
    from ._lib import check, lib
    n, hops = rings.n, rings.hops
    m = METRIC_IDS.get(str(metric).lower())
    if m is None:
        if not metric or not isinstance(metric, str):
            raise TypeError("Need to specify a metric.")
        raise NotImplementedError("{} metric is not implemented.".format(metric))
    sv, order = torch.sort(psi, dim=1, stable=True)
    order = order.to(torch.int32).contiguous()
    sv = sv.contiguous()
    sizes = rings.sizes.contiguous()
    U = torch.zeros((n, n), dtype=torch.float64, device=psi.device)
    for r0 in range(0, n, 32768):
        check(lib.hsd_pairwise_worker(engine._ptr(sv), engine._ptr(order), engine._ptr(rings.bitmaps),
                                      engine._ptr(sizes), engine._ptr(rings.bit_of), n, hops, hop_end, m,
                                      r0, min(32768, n - r0), engine._ptr(U), U.stride(0), engine._stream()))
    return U + U.t()   # mirror the strict upper triangle (model/HSD.py:132-134)
