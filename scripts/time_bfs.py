"""BFS + signature kernel alone on a subset of sources (dev aid; ncu target)."""
import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
n, hops, n_src = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
rows = torch.arange(0, n, max(1, n // n_src), dtype=torch.int32, device="cuda")[:n_src]
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sig, sizes, _, st = engine.ring_signature_degree(dg, hops, rows=rows)
    e1.record(); torch.cuda.synchronize()
    print(f"bfs {rows.numel()} sources of n={n}, {hops} hops: {e0.elapsed_time(e1):.3f} ms")
