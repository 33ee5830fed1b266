"""k-NN top-k kernel alone on the resident degree-mode matrix (dev aid; ncu target)."""
import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
n, hops, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 20
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
D, _ = engine.degree_distance_device(dg, hops)
for it in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); idx, val = engine.topk_rows(D, k); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"topk n={n} k={k}: {ms:.3f} ms, matrix read at {4.0 * n * n / ms / 1e6:.0f} GB/s")
