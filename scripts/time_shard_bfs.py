"""BFS phase of one emulated rank of an 8-way split (latency regime) — dev aid."""
import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
from hsd_b200.sharded import ShardedDegreeHSD
g = powerlaw_graph(20000, 5, seed=0); dg = engine.DeviceGraph.upload(g)
for world in (8, 4, 2):
    p = ShardedDegreeHSD(dg, 3, 0, world)
    for it in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); p.signatures(); b.record(); torch.cuda.synchronize()
    print(f"world {world}: rank-0 BFS phase {a.elapsed_time(b):.3f} ms, hub split: {p.hub_split is not None and int(p.hub_split[0].numel())}")
