"""BFS phase of one emulated rank of an 8-way split (latency regime) — dev aid."""
import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
from hsd_b200.sharded import ShardedDegreeHSD
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = powerlaw_graph(n, 5, seed=0); dg = engine.DeviceGraph.upload(g)
for world in (8, 4, 2, 1):
    p = ShardedDegreeHSD(dg, hops, 0, world)
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); p.signatures(); b.record(); torch.cuda.synchronize()
    print(f"world {world}: rank-0 ring/signature phase {a.elapsed_time(b):.3f} ms, variant {engine.ring_algorithm(dg.n, p.n_src, hops, 'cuda')}, hub split: {p.hub_split is not None and int(p.hub_split[0].numel())}")
