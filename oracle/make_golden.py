"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only
at /root/reference) in the build container.  TEST INFRASTRUCTURE ONLY.

The reference cannot be imported as shipped (pygsp is not
installable here, numpy.float is gone), so it is loaded under a two-line
pre-import shim: an empty `pygsp` module and `numpy.float = float`.  Nothing in
the reference is modified or copied; only its *outputs* (float64 arrays) and the
small graph/label data files it ships are stored, so the GPU box — which has no
/root/reference — can check parity against them.

    python oracle/make_golden.py            # writes tests/golden/

Run time ~1 min (europe full matrix through the reference's scipy loop is ~20 s).
"""
from __future__ import annotations

import io
import os
import sys
import types
import contextlib

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    """Import the reference's `model` / `tools` packages under the shim."""
    sys.modules.setdefault("pygsp", types.ModuleType("pygsp"))
    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001 - the reference uses the removed alias
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import model  # noqa: E402  (reference package)
    import tools  # noqa: E402
    from tools import hierarchy, metrics  # noqa: E402
    return model, tools, hierarchy, metrics


def read_graph(path, nodetype=str):
    import networkx as nx
    return nx.read_edgelist(path, create_using=nx.Graph, nodetype=nodetype, edgetype=float,
                            data=[("weight", float)])


def graph_arrays(graph):
    nodes = list(graph.nodes())
    idx = {v: i for i, v in enumerate(nodes)}
    edges = np.array([[idx[u], idx[v]] for u, v in graph.edges()], dtype=np.int32)
    return np.array([str(v) for v in nodes]), edges


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def rings_to_arrays(hier, nodes, node2idx, hop):
    """dict node -> layers  =>  (sizes[N, hop+1], flat sorted member indices)."""
    sizes = np.zeros((len(nodes), hop + 1), dtype=np.int32)
    flat = []
    for i, v in enumerate(nodes):
        for h, layer in enumerate(hier[v]):
            sizes[i, h] = len(layer)
            flat.extend(sorted(node2idx[w] for w in layer))
    return sizes, np.array(flat, dtype=np.int32)


def main():
    os.makedirs(OUT, exist_ok=True)
    model, tools, hierarchy, metrics = load_reference()
    from scipy.stats import wasserstein_distance

    # ---------------- graphs + labels shipped by the reference ----------------
    graphs = {}
    blob = {}
    for name in ["karate", "mkarate", "barbell", "tree", "europe", "usa"]:
        g = read_graph(f"{REF}/data/graph/{name}.edgelist")
        graphs[name] = g
        nodes, edges = graph_arrays(g)
        blob[f"{name}_nodes"] = nodes
        blob[f"{name}_edges"] = edges
    g = read_graph(f"{REF}/tests/robust_test/graph.edgelist", nodetype=int)
    g.add_edge(5, 6)  # tests/robust_test/main.py:47-52 (get_variated_graphs)[0]
    graphs["robust"] = g
    blob["robust_nodes"], blob["robust_edges"] = graph_arrays(g)
    g = read_graph(f"{REF}/tests/HeatKernel_test/line.edgelist")
    graphs["line"] = g
    blob["line_nodes"], blob["line_edges"] = graph_arrays(g)
    for name in ["mkarate", "barbell", "tree"]:
        lab = {}
        with open(f"{REF}/data/label/{name}.label") as f:
            for line in f:
                parts = line.split()
                if len(parts) >= 2:
                    lab[parts[0]] = parts[1]
        blob[f"{name}_labels"] = np.array([lab.get(str(v), "") for v in blob[f"{name}_nodes"]])
    np.savez_compressed(os.path.join(OUT, "graphs.npz"), **blob)

    # ---------------- robust.csv golden vector (tests/robust_test/robust.csv) --
    rows = np.loadtxt(f"{REF}/tests/robust_test/robust.csv", delimiter=",")
    np.savez_compressed(os.path.join(OUT, "robust_csv.npz"), node=rows[:, 0].astype(np.int32),
                        values=rows[:, 1:])

    # ---------------- reference runs -------------------------------------------
    def run_full(name, hop, scale):
        g = graphs[name]
        with quiet():
            m = model.HSD(g, name, 0, hop, "wasserstein")
            m.hierarchy = hierarchy.get_hierarchical_representation(g, hop)
            W = m.calculate_wavelets(scale, approx=False)
            D = m.calculate_structural_distance(scale, approx=False)
        sizes, flat = rings_to_arrays(m.hierarchy, m.nodes, m.node2idx, hop)
        hd = m.get_nodes_hierarchical_degree()
        return m, W, D, sizes, flat, np.array([hd[v] for v in m.nodes], dtype=np.int32)

    out = {}
    for name, hop, scale in [("karate", 3, 1.0), ("barbell", 2, 0.5), ("mkarate", 3, 1.0),
                             ("europe", 3, 1.0)]:
        m, W, D, sizes, flat, hd = run_full(name, hop, scale)
        out[f"{name}_hop"] = hop
        out[f"{name}_scale"] = scale
        out[f"{name}_D"] = D if name != "europe" else D[np.triu_indices(D.shape[0], 1)]
        out[f"{name}_ring_sizes"] = sizes
        out[f"{name}_ring_flat"] = flat
        out[f"{name}_hier_degree"] = hd
        if name != "europe":
            out[f"{name}_wavelets"] = W
        else:
            out["europe_checksum"] = D.sum()
        print(name, "D sum", D.sum(), "D[0,1]", D[0, 1])

        if name == "karate":
            # model/HSD.py:140-161 as written (needs self.wavelets, F7)
            m.wavelets = W
            for metric in ["wasserstein", "hellinger", "wasserstein_guass"]:
                m.metric = metric
                rows_ = np.stack([m._calculate_worker(i) for i in range(m.n_node)])
                out[f"karate_worker_{metric}"] = rows_
            # model/multiscale_HSD.py:45-73 without the pygsp-dependent __init__
            mm = object.__new__(model.MultiHSD)
            mm.__dict__.update(m.__dict__)
            out["karate_triple"] = np.array([mm.get_triple(W, v) for v in m.nodes])
            out["karate_layer_sum"] = np.array([mm.get_layer_sum(W, v) for v in m.nodes])

    # usa: reference wavelets + reference coefficient gather, scipy on a row sample
    g = graphs["usa"]
    hop, scale = 3, 1.0
    with quiet():
        m = model.HSD(g, "usa", 0, hop, "wasserstein")
        m.hierarchy = hierarchy.get_hierarchical_representation(g, hop)
        W = m.calculate_wavelets(scale, approx=False)
        coeffs = m.get_hierarchical_coeffcients(W)
    sample = np.array([0, 1, 7, 100, 333, 600, 901, 1189], dtype=np.int32)
    Drows = np.zeros((len(sample), m.n_node))
    for a, i in enumerate(sample):
        ci = coeffs[m.nodes[i]]
        for j in range(m.n_node):
            if j == i:
                continue
            cj = coeffs[m.nodes[j]]
            Drows[a, j] = sum(wasserstein_distance(ci[h], cj[h]) for h in range(hop + 1))
    sizes, _ = rings_to_arrays(m.hierarchy, m.nodes, m.node2idx, hop)
    out.update(usa_hop=hop, usa_scale=scale, usa_rows=sample, usa_Drows=Drows, usa_ring_sizes=sizes)
    print("usa sample rows sum", Drows.sum(), "D[0,1]", Drows[0, 1])

    # reference metric known answers (tests/other_test/main.py:9-13 and the aligned variant)
    out["kat_w1"] = np.array([wasserstein_distance([1, 2, 3, 4, 5], [15]),
                              metrics.calculate_distance([1, 2, 3, 4, 5], [15], "wasserstein")])
    rng = np.random.default_rng(0)
    P = [list(rng.random(int(n))) for n in rng.integers(0, 9, size=24)]
    out["aligned_cases_len"] = np.array([len(p) for p in P], dtype=np.int32)
    out["aligned_cases_flat"] = np.array([x for p in P for x in p])
    for metric in ["wasserstein", "hellinger", "wasserstein_guass"]:
        out[f"aligned_{metric}"] = np.array([[metrics.calculate_distance(list(p), list(q), metric)
                                              for q in P] for p in P])
    np.savez_compressed(os.path.join(OUT, "reference_runs.npz"), **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
