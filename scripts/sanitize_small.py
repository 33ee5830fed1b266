"""Small end-to-end pass of every kernel for compute-sanitizer memcheck (dev aid)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch, networkx as nx
from hsd_b200 import engine, rings
from hsd_b200.graph import powerlaw_graph
from model import HSD, MultiHSD
from model.GraphWave import GraphWave

g = powerlaw_graph(301, 3, seed=0)
dg = engine.DeviceGraph.upload(g)
D, sizes = engine.degree_distance_device(dg, 3)
idx, val = engine.topk_rows(D, 7)
pipe = engine.HostDegreePipeline(g, 3, n_chunks=3)
out = torch.empty((g.n, g.n), dtype=torch.float32).pin_memory(); pipe.run(out)
G = nx.barabasi_albert_graph(120, 3, seed=1)
m = HSD(G, "s", 0, 2, "wasserstein")
Dw = m.calculate_structural_distance(0.7, approx=True)
m.wavelets = m.calculate_wavelets(0.7, approx=False)
for metric in ("wasserstein", "hellinger", "wasserstein_guass"):
    m.metric = metric; m.distMat = None; m.parallel_calculate_HSD()
mm = MultiHSD(G, "s", 2, 3); e = mm.embed()
gw = GraphWave(G); gw.calculate_wavelets(1.0, approx=True); gw.embed(np.linspace(0, 10, 7))
torch.cuda.synchronize()
print("sanitize pass ok", float(D.sum()), float(Dw.sum()))
