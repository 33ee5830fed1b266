"""Mid-size graphs the reference ships (cora 2 708 nodes, facebook 5 908) as fixtures for the value-mode
path, plus outputs of the UNMODIFIED reference on cora's largest connected component (cora itself has 78
components: nodes of the small ones have empty hop-3 rings and the reference's scipy call raises
"Distribution can't be empty." — stored as `cora_raises`).  TEST INFRASTRUCTURE ONLY (see make_golden.py
for the import shim).  The reference's full `calculate_structural_distance` on cora is 14.7 M scipy calls
(~22 min), so the golden holds its exact wavelets' ring coefficients reduced to SAMPLED pairs: the reference
runs `calculate_wavelets(1.0, approx=False)` and `get_hierarchical_coeffcients` unmodified, and the loop body
of model/HSD.py:108-112 (scipy.stats.wasserstein_distance per hop) is applied to 400 sampled pairs.

    python oracle/make_golden_mid.py        # writes tests/golden/graphs_mid.npz, reference_cora.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, graph_arrays, load_reference, quiet, read_graph, rings_to_arrays  # noqa: E402


def main():
    model, tools, hierarchy, metrics = load_reference()
    from scipy.stats import wasserstein_distance
    blob = {}
    graphs = {}
    import networkx as nx
    for name in ["cora", "facebook"]:
        g = read_graph(f"{REF}/data/graph/{name}.edgelist")
        graphs[name] = g
        nodes, edges = graph_arrays(g)
        blob[f"{name}_nodes"], blob[f"{name}_edges"] = nodes, edges
        print(name, len(nodes), len(edges))
    np.savez_compressed(os.path.join(OUT, "graphs_mid.npz"), **blob)

    hop, scale = 3, 1.0
    full = graphs["cora"]
    with quiet():
        mf = model.HSD(full, "cora", scale, hop, "wasserstein")
        mf.hierarchy = hierarchy.get_hierarchical_representation(full, hop)
    sizes_full, _ = rings_to_arrays(mf.hierarchy, mf.nodes, mf.node2idx, hop)
    blob_raises = bool((sizes_full == 0).any())      # model/HSD.py:111 raises on the first such pair
    lcc = max(nx.connected_components(full), key=len)
    g = nx.Graph()
    g.add_nodes_from(v for v in full.nodes() if v in lcc)           # the reference's order of first appearance
    g.add_edges_from((u, v) for u, v in full.edges() if u in lcc)
    nodes_l, edges_l = graph_arrays(g)
    blob["cora_lcc_nodes"], blob["cora_lcc_edges"] = nodes_l, edges_l
    np.savez_compressed(os.path.join(OUT, "graphs_mid.npz"), **blob)
    with quiet():
        m = model.HSD(g, "cora_lcc", scale, hop, "wasserstein")
        m.hierarchy = hierarchy.get_hierarchical_representation(g, hop)
        wav = m.calculate_wavelets(scale, approx=False)
        coeffs = m.get_hierarchical_coeffcients(wav)
    nodes = m.nodes
    n = len(nodes)
    sizes, _ = rings_to_arrays(m.hierarchy, nodes, m.node2idx, hop)
    rng = np.random.default_rng(0)
    pairs = sorted({tuple(sorted(p)) for p in rng.integers(0, n, size=(400, 2)).tolist() if p[0] != p[1]})
    d = []
    for i, j in pairs:   # model/HSD.py:108-112
        c1, c2 = coeffs[nodes[i]], coeffs[nodes[j]]
        d.append(sum(wasserstein_distance(c1[h], c2[h]) for h in range(hop + 1)))
    np.savez_compressed(os.path.join(OUT, "reference_cora.npz"), hop=hop, scale=scale, ring_sizes=sizes,
                        cora_raises=blob_raises, cora_ring_sizes=sizes_full,
                        pairs=np.array(pairs, dtype=np.int32), dist=np.array(d, dtype=np.float64),
                        wavelet_rows=wav[[0, 1, 1000, n - 1]], wavelet_row_ids=np.array([0, 1, 1000, n - 1]))
    print("cora: max ring", sizes.max(axis=0), "sampled pairs", len(pairs), "dist range", min(d), max(d))


if __name__ == "__main__":
    main()
