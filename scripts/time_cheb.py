"""Chebyshev SpMM block timing (dev aid; ncu target)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from hsd_b200 import wavelets as wv
from hsd_b200.graph import powerlaw_graph
n, order, S, C = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = powerlaw_graph(n, 5, seed=0)
lmax = wv.estimate_lmax(g)
csr = wv.DeviceCSR(g)
scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / lmax), S))
coeffs = np.stack([wv.cheby_coefficients(s, lmax, order) for s in scales])
work = torch.empty((3, n, C), dtype=torch.float64, device="cuda"); out = torch.empty((S, n, C), dtype=torch.float64, device="cuda")
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); wv.cheb_wavelet_block(csr, lmax, coeffs, 0, C, 1e-4 / n, work, out); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    byts = order * (8.0 * g.nnz + 4 * (n + 1) + (3 + 2 * S) * 8.0 * n * C)
    print(f"cheb n={n} order={order} S={S} C={C}: {ms:.3f} ms, {ms/C*1e3:.1f} us/col, alg {byts/ms/1e6:.0f} GB/s")
print(f"checksum {float(out.sum()):.12e} abs {float(out.abs().sum()):.12e}")
