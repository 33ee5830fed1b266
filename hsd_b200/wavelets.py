"""Heat-kernel wavelets Psi_s = exp(-s L) on the device.

Reference: HSD.calculate_wavelets (model/HSD.py:48-67) and
GraphWave.calculate_wavelets (model/GraphWave.py:29-49): either pygsp's
Chebyshev approximation applied to one unit impulse at a time, or a dense
``eigh``; then the threshold ``x if x > coeff/N else 0``.

Here the Chebyshev path is one CSR SpMM kernel over a block of impulse columns
(hsd_cheb_spmm); the exact path takes the eigendecomposition from the vendor solver
through torch (cuSOLVER, FP64 — the step *before* the hot path, LAPACK in the
reference) and forms U diag(exp(-s lambda)) U^T with the threshold in one fused
kernel (hsd_exact_wavelets).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import engine
from ._lib import check, lib
from .graph import CSRGraph


def cheby_coefficients(scale: float, lmax: float, order: int) -> np.ndarray:
    """Chebyshev coefficients of g(x) = exp(-scale * x) on [0, lmax] by
    (order+1)-point Chebyshev-Gauss quadrature — what pygsp's compute_cheby_coeff
    returns for Heat(tau = scale * lmax) (call site model/HSD.py:52-53)."""
    n = order + 1
    theta = np.pi * (np.arange(n) + 0.5) / n
    g = np.exp(-scale * (lmax / 2.0) * (np.cos(theta) + 1.0))
    k = np.arange(order + 1)[:, None]
    return (2.0 / n) * (np.cos(k * theta[None, :]) @ g)


def estimate_lmax(g: CSRGraph, device=None, tol: float = 1e-6, max_iter: int = 2000) -> float:
    """1.01 x the largest Laplacian eigenvalue — what pygsp's Graph.estimate_lmax returns
    (ARPACK there, tol 5e-3; call sites model/HSD.py:51, model/multiscale_HSD.py:28) — by
    power iteration on the device (hsd_laplacian_spmv + a Rayleigh quotient), FP64, fixed start
    vector, stopped when the eigenvalue estimate moves by less than `tol` relative.  pygsp's own
    value is only reproducible to ~2 digits, so parity tests pass lmax explicitly to both sides."""
    dev = device or engine.require_cuda()
    csr = DeviceCSR(g, dev)
    n = g.n
    if n == 1:
        return 0.0
    # deterministic, not orthogonal to the top eigenvector in practice: degree-weighted alternating signs
    x = torch.from_numpy((np.diff(g.rowptr).astype(np.float64) + 1.0) * np.where(np.arange(n) % 2, -1.0, 1.0)).to(dev)
    x /= torch.linalg.vector_norm(x)
    y = torch.empty_like(x)
    lam_prev = 0.0
    for it in range(max_iter):
        check(lib.hsd_laplacian_spmv(engine._ptr(csr.rowptr), engine._ptr(csr.col), n, engine._ptr(x),
                                     engine._ptr(y), engine._stream()))
        if it % 8 == 7:           # the Rayleigh quotient needs a host read: poll every 8 iterations
            lam = float(torch.dot(x, y))
            if abs(lam - lam_prev) <= tol * abs(lam):
                lam_prev = lam
                break
            lam_prev = lam
        nrm = torch.linalg.vector_norm(y)
        x, y = y / nrm, x
    return 1.01 * lam_prev


def estimate_lmax_arpack(g: CSRGraph) -> float:
    """The pygsp recipe itself (host ARPACK: k=1, tol=5e-3, ncv=min(N,10)) with a fixed-seed start
    vector; kept for comparison, not used by the models."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    n = g.n
    rows = np.repeat(np.arange(n), np.diff(g.rowptr))
    keep = rows != g.col
    A = sp.csr_matrix((np.ones(int(keep.sum())), (rows[keep], g.col[keep])), shape=(n, n))
    L = sp.diags(np.asarray(A.sum(1)).ravel()) - A
    if n < 12:
        lam = float(np.linalg.eigvalsh(L.toarray())[-1])
    else:
        lam = float(spla.eigsh(L.tocsc(), k=1, tol=5e-3, ncv=min(n, 10), v0=np.random.default_rng(0).standard_normal(n),
                               return_eigenvectors=False)[0])
    return 1.01 * lam


class DeviceCSR:
    """CSR in ORIGINAL node order on the device (the Laplacian of the SpMM)."""

    def __init__(self, g: CSRGraph, device=None):
        dev = device or engine.require_cuda()
        self.n = g.n
        self.rowptr = torch.from_numpy(g.rowptr).to(dev)
        self.col = torch.from_numpy(g.col).to(dev) if g.col.size else torch.zeros(1, dtype=torch.int32, device=dev)


def column_block(n: int, n_scales: int, budget_bytes: float = 8e9, l2_bytes: float = 105e6) -> int:
    """Impulse columns per SpMM pass.  T_{k-1} (N x C doubles) is re-read once per neighbour, so the
    block is sized to about the L2 (measured at N = 50k, microseconds per column: C = 128 -> 25.6,
    192 -> 22.5, 256 -> 20.9, 384 -> 24.5, 1024 -> 25.3), within a memory budget for the 3 + S work
    planes.  105e6 / (8 N) gives 256 at N = 50k; blocks of >= 256 columns also get one node per CTA."""
    c = min(int(l2_bytes / (8 * n)), int(budget_bytes / ((3 + n_scales) * n * 8)))
    c = max(64, c // 64 * 64)
    return n if c >= n else c


def cheb_wavelet_block(csr: DeviceCSR, lmax: float, coeffs: np.ndarray, col0: int, n_cols: int,
                       threshold: float, work: Optional[torch.Tensor] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[s, v, c] = Psi_{s}[col0 + c, v] for a block of impulse columns (float64).
    coeffs: host float64[n_scales, order+1].  The kernel works on column pairs: an odd block
    (only ever the last one of an odd-sized graph) is run one column wider into a scratch
    buffer and copied back."""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
    n_scales, order1 = coeffs.shape
    dev = csr.rowptr.device
    if n_cols % 2:
        wide = cheb_wavelet_block(csr, lmax, coeffs, col0, n_cols + 1, threshold)
        if out is None:
            return wide[:, :, :n_cols].contiguous()
        out.copy_(wide[:, :, :n_cols])
        return out
    if work is None:
        work = torch.empty((3, csr.n, n_cols), dtype=torch.float64, device=dev)
    if out is None:
        out = torch.empty((n_scales, csr.n, n_cols), dtype=torch.float64, device=dev)
    check(lib.hsd_cheb_spmm(engine._ptr(csr.rowptr), engine._ptr(csr.col), csr.n, float(lmax),
                            coeffs.ctypes.data, n_scales, order1 - 1, col0, n_cols, float(threshold),
                            engine._ptr(work), engine._ptr(out), engine._stream()))
    return out


def cheb_wavelets_dense(csr: DeviceCSR, scale: float, lmax: float, order: int,
                        thr_coeff: Optional[float]) -> torch.Tensor:
    """Full N x N Psi (row i = response to impulse i, model/HSD.py:59) — reference-sized graphs."""
    n = csr.n
    thr = -np.inf if thr_coeff is None else thr_coeff * 1.0 / n
    coeffs = cheby_coefficients(scale, lmax, order)[None, :]
    psi = torch.empty((n, n), dtype=torch.float64, device=csr.rowptr.device)
    cb = column_block(n, 1)
    for c0 in range(0, n, cb):
        c = min(cb, n - c0)
        blk = cheb_wavelet_block(csr, lmax, coeffs, c0, c, thr)
        psi[c0:c0 + c, :] = blk[0].t()
    return psi


def exact_wavelets_dense(L: torch.Tensor, scale: float, thr_coeff: Optional[float],
                         eig=None) -> torch.Tensor:
    """U diag(exp(-s lambda)) U^T then threshold (model/HSD.py:61-66), FP64 on device: ONE fused
    kernel (hsd_exact_wavelets: scaling, product, threshold, symmetric mirror) — the un-thresholded
    N x N product is never materialised.  Only the eigendecomposition is a vendor call (cuSOLVER
    through torch.linalg.eigh), as numpy's LAPACK eigh is in the reference."""
    lam, U = eig if eig is not None else torch.linalg.eigh(L)
    n = L.shape[0]
    U = U.contiguous()
    lam = lam.contiguous()
    psi = torch.empty((n, n), dtype=torch.float64, device=U.device)
    thr = 0.0 if thr_coeff is None else thr_coeff * 1.0 / n
    check(lib.hsd_exact_wavelets(engine._ptr(U), U.stride(0), engine._ptr(lam), n, float(scale), float(thr),
                                 0 if thr_coeff is None else 1, engine._ptr(psi), psi.stride(0), engine._stream()))
    return psi
