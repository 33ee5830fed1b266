"""Value-mode (wavelet ring signal) timing on the reference's mid-size graphs (dev aid):
stages of HSD.structural_distance_device and MultiHSD.parallel_calculate_structural_distance."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch, networkx as nx
from hsd_b200 import rings as R
from model import HSD, MultiHSD

def graph(name):
    z = np.load("tests/golden/graphs_mid.npz")
    nodes = [str(v) for v in z[f"{name}_nodes"]]
    g = nx.Graph(); g.add_nodes_from(nodes); g.add_edges_from((nodes[u], nodes[v]) for u, v in z[f"{name}_edges"])
    return g

def ev(): return torch.cuda.Event(enable_timing=True)

for name in (sys.argv[1:] or ["cora_lcc", "facebook"]):
    g = graph(name)
    m = HSD(g, name, 1.0, 3, "wasserstein")
    m.lmax = None
    for rep in range(2):
        e = [ev() for _ in range(5)]
        e[0].record(); psi = m._wavelets_device(1.0, approx=True)
        e[1].record(); rs = m._rings()
        e[2].record(); vals, offs = R.sorted_ring_values(psi, rs)
        e[3].record(); D = R.value_distance(psi, rs, 0, 4, mode="w1")
        e[4].record(); torch.cuda.synchronize()
        t = [e[i].elapsed_time(e[i + 1]) for i in range(4)]
    print(f"{name}: n={m.n_node} max ring {rs.sizes.max(0).values.tolist()} total ring members {int(rs.sizes.sum())}")
    print(f"   cheb wavelets {t[0]:.2f} ms | rings {t[1]:.2f} | gather+sort {t[2]:.2f} | value_distance (gather+sort+W1 merge) {t[3]:.2f} ms | checksum {float(D.sum()):.9f}")
    mm = MultiHSD(g, name, 3, 4)
    mm.parallel_calculate_structural_distance()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    Dm = mm.parallel_calculate_structural_distance()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"   MultiHSD.parallel_calculate_structural_distance, 4 scales: {dt*1e3:.1f} ms, checksum {float(Dm.sum()):.9f}")
