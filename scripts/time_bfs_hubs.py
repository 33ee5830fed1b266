import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
n, hops = 20000, 3
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
def t(rows, reps=5):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); engine.ring_signature_degree(dg, hops, rows=rows); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
for rows in ([0], [5], [19999], list(range(8)), list(range(148)), list(range(0, 20000, 8)), list(range(1, 20000, 8))):
    r = torch.tensor(rows, dtype=torch.int32, device="cuda")
    print(len(rows), "sources starting", rows[0], "deg", int(g.degree[rows[0]]), ": %.3f ms" % t(r))
