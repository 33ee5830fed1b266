"""Quick stage timing of the degree-mode path at a given size (dev aid)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
print("n", n, "bins", dg.n_bins, "k_used", dg.k_used(hops), "nnz", g.nnz)

def ev():
    return torch.cuda.Event(enable_timing=True)

peak = engine.fp32_issue_peak()
print("fp32 issue peak %.2f T lane-ops/s" % (peak / 1e12))
out = torch.empty((n, n), dtype=torch.float32, device="cuda")
for it in range(4):
    e = [ev() for _ in range(5)]
    e[0].record()
    sig, sizes, _, status = engine.ring_signature_degree(dg, hops)
    e[1].record()
    sigT = engine.alloc_signature_table(dg.k_used(hops), n, sig.device)
    engine.signature_transpose(sig, dg.k_used(hops), sigT)
    e[2].record()
    engine.pairwise_l1(sigT, n, symmetric=True, out=out, k_used=dg.k_used(hops))
    e[3].record()
    torch.cuda.synchronize()
    t = [e[i].elapsed_time(e[i + 1]) for i in range(3)]
    pairs = n * (n - 1) / 2
    flops = 2 * pairs * dg.k_used(hops)
    print("iter %d: bfs+sig %.3f ms, transpose %.3f ms, pairwise %.3f ms | pairs/s %.3e | pairwise %.2f TFLOP/s (%.1f%% of measured issue peak)"
          % (it, t[0], t[1], t[2], pairs / (sum(t) * 1e-3), flops / (t[2] * 1e-3) / 1e12,
             100 * flops / (t[2] * 1e-3) / peak))
print("ring size means", sizes.float().mean(0).tolist(), "status", status.item())
print("D checksum", out.double().sum().item(), "D[0,1]", out[0, 1].item())
