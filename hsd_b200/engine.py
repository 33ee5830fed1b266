"""Device pipeline: CSR -> rings -> signatures -> pairwise distances.

torch is used for device memory, streams and (in sharded.py) torch.distributed;
every arithmetic step is a hand-written kernel reached through the C-ABI in
include/hsd_b200.h.  There is no CPU path: all functions here require CUDA
tensors and raise otherwise.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

from ._lib import check, lib
from .graph import CSRGraph

PAIR_TILE = 128
PAIR_KCHUNK = 16


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("hsd_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("hsd_b200 kernels need contiguous tensors")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("hsd_b200 requires a CUDA device (B200 / sm_100a); no CPU fallback exists")
    return torch.device("cuda", torch.cuda.current_device())


def roundup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# --------------------------------------------------------------------------
# device-resident graph
# --------------------------------------------------------------------------
@dataclass
class DeviceGraph:
    """CSR in degree-ascending order on the device + the shared degree support."""
    n: int
    rowptr: torch.Tensor
    col: torch.Tensor
    orig_of: torch.Tensor      # int32[N] new id -> original index
    new_of: torch.Tensor       # int32[N]
    bin_end: torch.Tensor      # int32[B]
    delta: torch.Tensor        # float32[max(B-1,1)]
    n_bins: int
    support: np.ndarray        # host float64[B]
    include_zero: bool

    @classmethod
    def upload(cls, g: CSRGraph, include_zero: bool = False, device=None,
               non_blocking: bool = True) -> "DeviceGraph":
        dev = device or require_cuda()
        o = g.degree_order()
        sup, bin_end, delta = o.support(include_zero)
        if delta.size == 0:
            delta = np.zeros(1, dtype=np.float32)

        def up(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=non_blocking)

        return cls(n=g.n, rowptr=up(o.rowptr), col=up(o.col) if o.col.size else torch.zeros(1, dtype=torch.int32, device=dev),
                   orig_of=up(o.orig_of), new_of=up(o.new_of),
                   bin_end=up(bin_end), delta=up(delta), n_bins=int(sup.size), support=sup,
                   include_zero=include_zero)

    @classmethod
    def upload_device_order(cls, g: CSRGraph, include_zero: bool = False, device=None) -> "DeviceGraph":
        """Same result as upload(), but the degree order (stable argsort by degree, relabelled and
        re-sorted CSR, support tables) is computed on the device with torch sort / gather ops from the
        CSR in original order — 44 ms of host numpy at 100k nodes become ~2 ms, which matters when
        the graph changes between steps (DynamicHSD).  One small D2H (the distinct degrees) remains:
        the host needs the support to size the signature."""
        dev = device or require_cuda()
        n = g.n
        rowptr = torch.from_numpy(g.rowptr).to(dev)
        col = torch.from_numpy(g.col).to(dev) if g.col.size else torch.zeros(0, dtype=torch.int32, device=dev)
        deg = (rowptr[1:] - rowptr[:-1])
        orig_of = torch.argsort(deg, stable=True)                       # int64
        new_of = torch.empty(n, dtype=torch.int64, device=dev)
        new_of[orig_of] = torch.arange(n, dtype=torch.int64, device=dev)
        rows = torch.repeat_interleave(torch.arange(n, dtype=torch.int64, device=dev), deg.long())
        key = torch.sort(new_of[rows] * n + new_of[col.long()]).values
        sdeg = deg[orig_of]
        rowptr2 = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        rowptr2[1:] = torch.cumsum(sdeg, 0)
        # the BFS kernel reads column indices in aligned groups of 4 (LDG.128): pad the tail
        col2 = torch.zeros(key.numel() + (-key.numel()) % 4 + 4, dtype=torch.int32, device=dev)
        col2[:key.numel()] = (key % n).int()
        sup_d = torch.unique(sdeg)                                      # ascending
        if include_zero and (sup_d.numel() == 0 or int(sup_d[0]) != 0):
            sup_d = torch.cat([torch.zeros(1, dtype=sup_d.dtype, device=dev), sup_d])
        bin_end = torch.searchsorted(sdeg.contiguous(), sup_d, right=True).int()
        delta = (sup_d[1:] - sup_d[:-1]).float()
        if delta.numel() == 0:
            delta = torch.zeros(1, dtype=torch.float32, device=dev)
        sup = sup_d.cpu().numpy().astype(np.float64)
        return cls(n=n, rowptr=rowptr2, col=col2 if col2.numel() else torch.zeros(1, dtype=torch.int32, device=dev),
                   orig_of=orig_of.int(), new_of=new_of.int(), bin_end=bin_end.contiguous(), delta=delta.contiguous(),
                   n_bins=int(sup.size), support=sup, include_zero=include_zero)

    @property
    def n_words(self) -> int:
        return (self.n + 31) // 32

    @property
    def nnz(self) -> int:
        """CSR entries (rowptr[n]); the col array is padded beyond it for 16-byte reads."""
        if getattr(self, "_nnz", None) is None:
            self._nnz = int(self.rowptr[-1].item())
        return self._nnz

    def k_used(self, hops: int) -> int:
        """Signature length: hop 0 is one scalar (the source degree), every later hop
        is the delta-scaled CDF over the B-1 gaps of the shared support."""
        return 1 + hops * (self.n_bins - 1)


# --------------------------------------------------------------------------
# K1/K2
# --------------------------------------------------------------------------
_BFS_WS = {}   # device index -> int32 workspace tensor (grow-only), registered with the library


def ensure_bfs_workspace(n_nodes: int, device) -> None:
    """Graphs above ~400k nodes keep the BFS bitmaps in a global workspace (hsd_bfs_workspace_words);
    allocate / grow it and register it with the library before a BFS launch.  No-op otherwise."""
    words = int(lib.hsd_bfs_workspace_words(n_nodes))
    if not words:
        return
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _BFS_WS.get(key)
    if ws is None or ws.numel() < words:
        ws = torch.empty(words, dtype=torch.int32, device=dev)
        _BFS_WS[key] = ws
    check(lib.hsd_bfs_set_workspace(_ptr(ws), ws.numel()))


_DENSE_WS = {}   # device index -> int32 workspace tensor (grow-only) of the dense ring variant


def ring_algorithm(n_nodes: int, n_src: int, hops: int, device, distinct: bool = True) -> str:
    """'dense' (hsd_ring_signature_degree_dense: bitmap dynamic programming over all nodes) or 'frontier'
    (hsd_ring_signature_degree: one frontier-expansion BFS per source).  Dense pays O(E N / 32) per level
    for ALL nodes at the intermediate levels, so it is chosen when at least a fifth (a tenth from 64k nodes)
    of the nodes are sources (HSD_RING_DENSE_MIN_FRAC), hops >= 2 and its two N x N-bit tables fit
    in a quarter of the free device memory.  HSD_RING_ALGO=dense|frontier overrides
    (dense still needs hops >= 1)."""
    import os
    force = os.environ.get("HSD_RING_ALGO", "")
    if hops < 1 or (n_nodes + 31) // 32 * 8 > 200 * 1024:
        return "frontier"
    if not distinct:          # the dense variant emits one output row per node: sources must be distinct
        return "frontier"
    if force in ("dense", "frontier"):
        return force
    # measured break-even per rank of a W-way split (scripts/time_shard_bfs.py, profiles/r2_rings_notes.md):
    # 20k nodes: dense wins up to W = 4 (0.177 vs 0.186 ms) and loses at W = 8 (0.170 vs 0.132: latency-bound);
    # 100k nodes: dense still wins at W = 8 (4.9 vs 5.3 ms)
    min_frac = float(os.environ.get("HSD_RING_DENSE_MIN_FRAC", "0.2" if n_nodes < 65536 else "0.1"))
    if hops < 2 or n_src < min_frac * n_nodes or n_nodes < 512:
        return "frontier"
    words = int(lib.hsd_ring_dense_workspace_words(n_nodes))
    dev = torch.device(device)
    ws = _DENSE_WS.get(dev.index if dev.index is not None else torch.cuda.current_device())
    if ws is not None and ws.numel() >= words:
        return "dense"
    free, _ = torch.cuda.mem_get_info(dev)
    return "dense" if words * 4 <= free // 4 else "frontier"


def dense_ring_workspace(n_nodes: int, device) -> torch.Tensor:
    words = int(lib.hsd_ring_dense_workspace_words(n_nodes))
    dev = torch.device(device)
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    ws = _DENSE_WS.get(key)
    if ws is None or ws.numel() < words:
        _DENSE_WS[key] = None
        ws = torch.empty(words, dtype=torch.int32, device=dev)
        _DENSE_WS[key] = ws
    return ws


def launch_ring_signature(dg_rowptr, dg_col, n, nnz, src, out_rows, n_src, hops, bin_end, delta, n_bins,
                          sig, ld, sizes, bitmaps, empty_as_zero, status, device, peers=None, threads=0,
                          stream=None, distinct: bool = True) -> str:
    """One call site for the K1/K2 entry points: picks the dense or the frontier variant (same outputs,
    bit for bit) and launches it on `stream` (default: the current stream).  Returns the variant used."""
    stream = _stream() if stream is None else stream
    algo = ring_algorithm(n, n_src, hops, device, distinct)
    n_peers = int(peers.numel()) if peers is not None else 0
    if algo == "dense":
        ws = dense_ring_workspace(n, device)
        check(lib.hsd_ring_signature_degree_dense(
            _ptr(dg_rowptr), _ptr(dg_col), n, _ptr(src), _ptr(out_rows), n_src, hops, _ptr(bin_end), _ptr(delta),
            n_bins, _ptr(sig), ld, _ptr(peers) if n_peers else None, n_peers, _ptr(sizes), _ptr(bitmaps),
            empty_as_zero, _ptr(status), _ptr(ws), ws.numel(), nnz, stream))
        return algo
    ensure_bfs_workspace(n, device)
    if n_peers:
        check(lib.hsd_ring_signature_degree_allgather(
            _ptr(dg_rowptr), _ptr(dg_col), n, _ptr(src), _ptr(out_rows), n_src, hops, _ptr(bin_end), _ptr(delta),
            n_bins, _ptr(sig), ld, _ptr(peers), n_peers, _ptr(sizes), empty_as_zero, _ptr(status), threads, stream))
    else:
        check(lib.hsd_ring_signature_degree(
            _ptr(dg_rowptr), _ptr(dg_col), n, _ptr(src), _ptr(out_rows), n_src, hops, _ptr(bin_end), _ptr(delta),
            n_bins, _ptr(sig), ld, _ptr(sizes), _ptr(bitmaps), empty_as_zero, _ptr(status), threads, stream))
    return algo


# ---- column-split dense variant (one process per GPU; SURVEY §8 e) ----
def ring_cols_range(n_nodes: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the 16-byte bitmap pieces rank `rank` of `world` owns."""
    n4 = ((n_nodes + 31) // 32 + 3) // 4
    per = (n4 + world - 1) // world
    q0 = min(rank * per, n4)
    return q0, min(q0 + per, n4)


def ring_counts_cols(dg: DeviceGraph, hops: int, rank: int, world: int,
                     counts: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Partial prefix counts int32[N, hops*(B-1) + hops] of this rank's bitmap columns for ALL nodes
    (hsd_ring_counts_dense_cols); rows are degree-order ids.  Sum over the ranks (integer all-reduce),
    then signature_from_counts."""
    dev = dg.rowptr.device
    q0, q1 = ring_cols_range(dg.n, rank, world)
    ld_c = hops * (dg.n_bins - 1) + hops
    if counts is None:
        counts = torch.empty((dg.n, ld_c), dtype=torch.int32, device=dev)
    words = int(lib.hsd_ring_cols_workspace_words(dg.n, q0, q1))
    key = ("cols", dev.index if dev.index is not None else torch.cuda.current_device())
    ws = _DENSE_WS.get(key)
    if ws is None or ws.numel() < max(words, 4):
        _DENSE_WS[key] = None
        ws = torch.empty(max(words, 4), dtype=torch.int32, device=dev)
        _DENSE_WS[key] = ws
    check(lib.hsd_ring_counts_dense_cols(_ptr(dg.rowptr), _ptr(dg.col), dg.n, dg.nnz, hops, _ptr(dg.bin_end),
                                         dg.n_bins, q0, q1, _ptr(counts), counts.stride(0), _ptr(ws), ws.numel(),
                                         _stream()))
    return counts


def signature_from_counts(dg: DeviceGraph, hops: int, counts: torch.Tensor, src: torch.Tensor,
                          out_rows: torch.Tensor, sig: Optional[torch.Tensor], sizes: Optional[torch.Tensor],
                          empty: str, status: torch.Tensor) -> None:
    """Summed counts -> signature rows / ring sizes / empty-ring flag (hsd_ring_signature_from_counts)."""
    check(lib.hsd_ring_signature_from_counts(
        _ptr(dg.rowptr), _ptr(counts), counts.stride(0), _ptr(src), _ptr(out_rows), int(src.numel()), hops,
        _ptr(dg.delta), dg.n_bins, _ptr(sig), sig.stride(0) if sig is not None else 0, _ptr(sizes),
        1 if empty == "zero" else 0, _ptr(status), _stream()))


def ring_signature_degree(dg: DeviceGraph, hops: int, rows: Optional[torch.Tensor] = None,
                          want_sig: bool = True, want_sizes: bool = True,
                          want_bitmaps: bool = False, empty: str = "raise",
                          sig_ld: Optional[int] = None, sig_out: Optional[torch.Tensor] = None):
    """Run the BFS + degree-CDF kernel for the given sources.

    rows: int32 CUDA tensor of ORIGINAL node indices (default: all nodes, in
    original order).  Output row r belongs to rows[r].  Returns
    (sig float32[n, sig_ld] | None, ring_sizes int32[n, hops+1] | None,
     bitmaps uint32-as-int32[n, hops+1, n_words] | None, status int32[1]).
    Bitmaps are indexed by degree-order id (map with dg.orig_of).
    sig_out: a caller-owned float32 CUDA tensor [n, ld >= k_used] (contiguous) that receives the
    signatures instead of a fresh allocation (DynamicHSD keeps two and swaps them per update)."""
    if empty not in ("raise", "zero"):
        raise ValueError("empty must be 'raise' or 'zero'")
    if empty == "zero" and want_sig and not dg.include_zero:
        raise ValueError("empty='zero' needs a DeviceGraph uploaded with include_zero=True")
    dev = dg.rowptr.device
    if rows is None:
        src = dg.new_of
        n_src = dg.n
    else:
        rows = rows.to(device=dev, dtype=torch.int64)
        src = dg.new_of[rows].contiguous()
        n_src = int(rows.numel())
    out_rows = torch.arange(n_src, dtype=torch.int32, device=dev)
    k_used = dg.k_used(hops)
    ld = int(sig_ld) if sig_ld is not None else roundup(k_used, 4)
    if sig_out is not None:
        if (not want_sig or sig_out.dtype != torch.float32 or sig_out.dim() != 2 or sig_out.shape[0] != n_src
                or sig_out.shape[1] < k_used or sig_out.shape[1] % 4 or not sig_out.is_contiguous()):
            raise ValueError("sig_out must be a contiguous float32 [n_sources, ld >= k_used, ld % 4 == 0] tensor")
        sig, ld = sig_out, int(sig_out.shape[1])
    else:
        sig = torch.empty((n_src, ld), dtype=torch.float32, device=dev) if want_sig else None
    if sig is not None and ld > k_used:
        sig[:, k_used:].zero_()
    sizes = torch.empty((n_src, hops + 1), dtype=torch.int32, device=dev) if want_sizes else None
    bitmaps = (torch.empty((n_src, hops + 1, dg.n_words), dtype=torch.int32, device=dev)
               if want_bitmaps else None)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    distinct = rows is None or int(torch.unique(rows).numel()) == n_src
    launch_ring_signature(dg.rowptr, dg.col, dg.n, dg.nnz, src, out_rows, n_src, hops, dg.bin_end, dg.delta,
                          dg.n_bins, sig, ld, sizes, bitmaps, 1 if empty == "zero" else 0, status, dev,
                          distinct=distinct)
    return sig, sizes, bitmaps, status


# --------------------------------------------------------------------------
# layout + K3
# --------------------------------------------------------------------------
def alloc_signature_table(k_used: int, n: int, device) -> torch.Tensor:
    """Zeroed K-major table float32[k_pad][n_pad] for the pairwise kernel."""
    k_pad = roundup(max(k_used, 1), PAIR_KCHUNK)
    n_pad = roundup(n, 4)
    return torch.zeros((k_pad, n_pad), dtype=torch.float32, device=device)


def signature_transpose(sig: torch.Tensor, k_used: int, sigT: torch.Tensor, col0: int = 0,
                        src_rows: Optional[torch.Tensor] = None) -> None:
    """sigT[k, col0 + r] = sig[src_rows[r] if src_rows is given else r, k]."""
    n_rows = sig.shape[0] if src_rows is None else int(src_rows.numel())
    check(lib.hsd_signature_transpose(_ptr(sig), sig.stride(0), n_rows, k_used,
                                      _ptr(sigT), sigT.stride(0), col0, _ptr(src_rows), _stream()))


def scatter_symmetric(blk: torch.Tensor, idx: torch.Tensor, D: torch.Tensor, mirror: bool = True) -> None:
    """D[idx[a], :n] = blk[a] and (mirror) D[:n, idx[a]] = blk[a] for the m recomputed rows of an
    incremental update; idx int64 ascending node ids, blk float32[m, n] row-major."""
    m, n = blk.shape
    if idx.dtype != torch.int64 or idx.numel() != m or blk.stride(1) != 1 or D.stride(1) != 1:
        raise ValueError("idx must be int64[m]; blk and D row-major")
    if D.shape[1] < n or (mirror and D.shape[0] < n):
        raise ValueError("D too small")
    check(lib.hsd_scatter_symmetric(blk.data_ptr(), blk.stride(0), m, n, _ptr(idx.contiguous()), D.data_ptr(),
                                    D.stride(0), 1 if mirror else 0, _stream()))


def pairwise_l1(sigT: torch.Tensor, n: int, row0: int = 0, n_rows: Optional[int] = None,
                col0: int = 0, n_cols: Optional[int] = None, symmetric: Optional[bool] = None,
                out: Optional[torch.Tensor] = None, k_used: Optional[int] = None) -> torch.Tensor:
    """out[i-row0, j-col0] = sum_k |sigT[k, i] - sigT[k, j]| (float32).  k_used = number of
    signature rows that hold data (default: all rows of the table)."""
    k_used = sigT.shape[0] if k_used is None else k_used
    if roundup(k_used, PAIR_KCHUNK) > sigT.shape[0]:
        raise ValueError("signature table has fewer rows than k_used rounded up to the K chunk")
    n_rows = n - row0 if n_rows is None else n_rows
    n_cols = n - col0 if n_cols is None else n_cols
    if symmetric is None:
        symmetric = (row0 == col0 and n_rows == n_cols)
    if out is None:
        out = torch.empty((n_rows, n_cols), dtype=torch.float32, device=sigT.device)
    if out.shape[0] < n_rows or out.shape[1] < n_cols or out.stride(1) != 1:
        raise ValueError("out too small or not row-major")
    check(lib.hsd_pairwise_l1(_ptr(sigT), k_used, sigT.stride(0), row0, n_rows, col0, n_cols,
                              1 if symmetric else 0, out.data_ptr(), out.stride(0), _stream()))
    return out


# --------------------------------------------------------------------------
# whole degree-mode path on one device
# --------------------------------------------------------------------------
class EmptyRingError(ValueError):
    """Raised where the reference's scipy call would raise
    ValueError("Distribution can't be empty.") (model/HSD.py:111)."""


def degree_distance_device(dg: DeviceGraph, hops: int, empty: str = "raise",
                           row0: int = 0, n_rows: Optional[int] = None,
                           out: Optional[torch.Tensor] = None,
                           check_status: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Full degree-mode HSD on the current device: returns (D float32[n_rows, N], ring_sizes).

    D[i, j] = sum_{h=0..hops} W1(deg over ring_h(i), deg over ring_h(j)) — the
    loop of model/HSD.py:98-114 with a degree-valued ring signal."""
    n = dg.n
    k_used = dg.k_used(hops)
    sig, sizes, _, status = ring_signature_degree(dg, hops, empty=empty)
    sigT = alloc_signature_table(k_used, n, sig.device)
    signature_transpose(sig, k_used, sigT, 0)
    n_rows = n - row0 if n_rows is None else n_rows
    full = (row0 == 0 and n_rows == n)
    D = pairwise_l1(sigT, n, row0, n_rows, 0, n, symmetric=full, out=out, k_used=k_used)
    if check_status and empty == "raise" and int(status.item()) & 1:
        raise EmptyRingError("Distribution can't be empty.")
    return D, sizes


def topk_rows(D: torch.Tensor, k: int, self_col0: int = 0, n_cols: Optional[int] = None,
              col_mask: Optional[torch.Tensor] = None):
    """k nearest neighbours of every row of a float32 distance block (row r is node self_col0 + r).
    Returns (idx int32[n_rows, k], dist float32[n_rows, k]) ordered by (distance, column).
    col_mask: optional int32 bitmap over columns (bit j set = column j may be a neighbour)."""
    if D.dtype != torch.float32 or D.stride(1) != 1:
        raise ValueError("D must be a row-major float32 CUDA tensor")
    n_rows = D.shape[0]
    n_cols = D.shape[1] if n_cols is None else n_cols
    idx = torch.empty((n_rows, k), dtype=torch.int32, device=D.device)
    val = torch.empty((n_rows, k), dtype=torch.float32, device=D.device)
    check(lib.hsd_topk_rows(D.data_ptr(), D.stride(0), n_rows, n_cols, k, self_col0, _ptr(col_mask),
                            _ptr(idx), _ptr(val), _stream()))
    return idx, val


def fp32_issue_peak(iters: int = 20000, reps: int = 3) -> float:
    """Measured FP32 CUDA-core issue rate in lane-ops/s (FADD sub + |.|-accumulate mix)."""
    dev = require_cuda()
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    ops = ctypes.c_int64(0)
    best = 0.0
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.hsd_fp32_peak_probe(_ptr(sink), iters, ctypes.byref(ops), _stream()))
        e1.record()
        e1.synchronize()
        best = max(best, ops.value / (e0.elapsed_time(e1) * 1e-3))
    return best


# --------------------------------------------------------------------------
# host-buffer pipeline (the e2e path: host CSR in, host matrix out)
# --------------------------------------------------------------------------
def _pin(a: np.ndarray) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.pin_memory() if torch.cuda.is_available() else t


class HostDegreePipeline:
    """Degree-mode HSD from HOST buffers to a HOST result, copies included.

    Inputs (pinned): the degree-ordered CSR and support tables of a CSRGraph.
    Output: float32 rows [row0, row0+n_rows) x N written into a pinned host tensor.
    The pairwise kernel runs panel by panel (symmetric trapezoids on one GPU, plain row
    blocks when the rows are a shard); each finished panel is copied device->host on a
    second stream while the next one computes, so the PCIe transfer of the result — the
    dominant cost at N = 20k (1.6 GB) — overlaps the kernels."""

    def __init__(self, g: CSRGraph, hops: int, empty: str = "raise", device=None,
                 row0: int = 0, n_rows: Optional[int] = None, n_chunks: int = 8,
                 host_mirror: Optional[bool] = None, host_threads: Optional[int] = None):
        """host_mirror (full matrices only; opt-in, HSD_E2E_HOST_MIRROR=1): ship only the upper
        trapezoid of every row panel over PCIe (0.86-0.95 GB instead of 1.6 GB at N = 20k) and let
        `host_threads` CPU threads mirror it into the rows below (hsd_mirror_upper_to_lower_host)
        while later panels are still in flight.  Off by default: on the B200 boxes of this pool
        (16 vCPUs, ~60 GB/s of host memory traffic) the mirror alone takes 26 ms, longer than
        the 15 ms of PCIe time it saves (measured 32.7 ms/step against 30.2 ms for the full copy,
        gpurun_out/r2_e2e_a.log -> profiles/r2_e2e_notes.md); it pays on hosts whose cores can move
        > 110 GB/s."""
        self.dev = device or require_cuda()
        self.g, self.hops, self.empty = g, hops, empty
        o = g.degree_order()
        sup, bin_end, delta = o.support(include_zero=(empty == "zero"))
        if delta.size == 0:
            delta = np.zeros(1, dtype=np.float32)
        self.n = g.n
        self.nnz = g.nnz
        self.n_bins = int(sup.size)
        self.host = {k: _pin(v) for k, v in dict(rowptr=o.rowptr, col=o.col if o.col.size else np.zeros(1, np.int32),
                                                 orig_of=o.orig_of, new_of=o.new_of, bin_end=bin_end,
                                                 delta=delta).items()}
        self.dev_in = {k: torch.empty_like(v, device=self.dev) for k, v in self.host.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in self.host.values())
        self.k_used = 1 + hops * (self.n_bins - 1)
        self.row0 = row0
        self.n_rows = self.n - row0 if n_rows is None else n_rows
        self.full = (row0 == 0 and self.n_rows == self.n)
        self.sig = torch.zeros((self.n, roundup(self.k_used, 4)), dtype=torch.float32, device=self.dev)
        self.sigT = alloc_signature_table(self.k_used, self.n, self.dev)
        self.out_rows_idx = torch.arange(self.n, dtype=torch.int32, device=self.dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.D = torch.empty((self.n_rows, self.n), dtype=torch.float32, device=self.dev)
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.panels = self._plan_panels(n_chunks)
        self.launches_per_step = 2 + len(self.panels)
        import os
        if host_mirror is None:
            host_mirror = os.environ.get("HSD_E2E_HOST_MIRROR", "0") == "1"
        self.host_mirror = bool(host_mirror) and self.full
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.host_threads = max(1, min(int(host_threads or os.environ.get("HSD_E2E_HOST_THREADS", 0) or avail), 64))
        if self.host_mirror:     # each row panel [p0, p0+pr) ships columns [p0, N) only
            self.d2h_bytes = sum(pr * (self.n - p0) * 4 for p0, pr in self.panels)
        else:
            self.d2h_bytes = self.n_rows * self.n * 4

    def _plan_panels(self, n_chunks: int):
        """Row panels (multiples of the 128-row tile).  Symmetric mode: tile row I costs
        (nt - I) tiles, so equal-work panels start narrow — the first copy starts early."""
        nt = (self.n_rows + PAIR_TILE - 1) // PAIR_TILE
        n_chunks = max(1, min(n_chunks, nt))
        if self.full:
            ntc = (self.n + PAIR_TILE - 1) // PAIR_TILE
            work = np.cumsum([ntc - i for i in range(nt)])
        else:
            work = np.cumsum(np.ones(nt))
        cuts = [0]
        for c in range(1, n_chunks):
            t = int(np.searchsorted(work, work[-1] * c / n_chunks)) + 1
            if t > cuts[-1] and t < nt:
                cuts.append(t)
        cuts.append(nt)
        return [(a * PAIR_TILE, min(b * PAIR_TILE, self.n_rows) - a * PAIR_TILE) for a, b in zip(cuts[:-1], cuts[1:])]

    def run(self, out: torch.Tensor) -> torch.Tensor:
        """out: pinned float32 host tensor [n_rows, N]. Returns it (synchronised)."""
        if out.shape != (self.n_rows, self.n) or out.dtype != torch.float32:
            raise ValueError("out must be float32 of shape (n_rows, N)")
        cur = torch.cuda.current_stream(self.dev)
        for k, v in self.host.items():
            self.dev_in[k].copy_(v, non_blocking=True)
        d = self.dev_in
        self.status.zero_()
        arrived = []
        if out.stride(1) != 1:
            raise ValueError("out must be row-major")
        launch_ring_signature(d["rowptr"], d["col"], self.n, self.nnz, d["new_of"], self.out_rows_idx, self.n,
                              self.hops, d["bin_end"], d["delta"], self.n_bins, self.sig, self.sig.stride(0),
                              None, None, 1 if self.empty == "zero" else 0, self.status, self.dev)
        signature_transpose(self.sig, self.k_used, self.sigT, 0)
        for (p0, pr) in self.panels:
            if self.full:
                # trapezoid: rows [p0, p0+pr) x cols [p0, N), mirrored into rows below
                view = self.D[p0:, p0:]
                check(lib.hsd_pairwise_l1(_ptr(self.sigT), self.k_used, self.sigT.stride(0),
                                          p0, pr, p0, self.n - p0, 1, view.data_ptr(), self.D.stride(0), _stream()))
            else:
                view = self.D[p0:]
                check(lib.hsd_pairwise_l1(_ptr(self.sigT), self.k_used, self.sigT.stride(0),
                                          self.row0 + p0, pr, 0, self.n, 0, view.data_ptr(), self.D.stride(0), _stream()))
            ev = torch.cuda.Event()
            ev.record(cur)
            self.copy_stream.wait_event(ev)
            if self.host_mirror:
                # upper trapezoid of the panel only: rows [p0, p0+pr) x columns [p0, N)
                check(lib.hsd_copy2d_to_host(out[p0:, p0:].data_ptr(), out.stride(0) * 4,
                                             self.D[p0:, p0:].data_ptr(), self.D.stride(0) * 4,
                                             (self.n - p0) * 4, pr, self.copy_stream.cuda_stream))
                done = torch.cuda.Event()
                done.record(self.copy_stream)
                arrived.append((done, p0, pr))
            else:
                with torch.cuda.stream(self.copy_stream):
                    out[p0:p0 + pr].copy_(self.D[p0:p0 + pr], non_blocking=True)
        for done, p0, pr in arrived:
            # the host cores fill D[j][i] = D[i][j] below the panel while later panels are still crossing PCIe
            done.synchronize()
            check(lib.hsd_mirror_upper_to_lower_host(out.data_ptr(), out.stride(0), self.n, p0, p0 + pr,
                                                     self.host_threads))
        self.copy_stream.synchronize()
        if self.empty == "raise" and int(self.status.item()) & 1:
            raise EmptyRingError("Distribution can't be empty.")
        return out
