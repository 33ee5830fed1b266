import sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
g = powerlaw_graph(2500, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
sig, sizes, _, st = engine.ring_signature_degree(dg, 3)
k = dg.k_used(3)
sigT = engine.alloc_signature_table(k, g.n, sig.device)
engine.signature_transpose(sig, k, sigT)
torch.cuda.synchronize(); print("sig ok")
for (r0, nr, sym) in [(0, 2500, True), (0, 1250, False), (1280, 1220, False), (1250, 1250, False), (1252, 1248, False)]:
    try:
        D = engine.pairwise_l1(sigT, g.n, r0, nr, 0 if not sym else r0, g.n if not sym else nr, symmetric=sym)
        torch.cuda.synchronize()
        print(r0, nr, sym, "ok", float(D.double().sum()))
    except Exception as e:
        print(r0, nr, sym, "FAIL", str(e)[:80]); break
