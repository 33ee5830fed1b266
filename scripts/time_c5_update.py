#!/usr/bin/env python
"""Where the time of a DynamicHSD update goes (BASELINE config 5): every engine call of
structural_distance_update() is bracketed by a device synchronize and timed on the host.

    python scripts/time_c5_update.py [N=100000] [HOP=4] [EDGES=5000]

Prints one JSON line: per-stage ms of (a) a from-scratch call, (b) the update after EDGES insertions."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import networkx as nx
    import torch
    from hsd_b200 import engine
    from model import DynamicHSD
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    hop = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    k_ins = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    stages = {}

    def timed(name, fn):
        def wrap(*a, **kw):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **kw)
            torch.cuda.synchronize()
            stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
            return r
        return wrap

    engine.DeviceGraph.upload_device_order = classmethod(
        timed("upload_device_order", engine.DeviceGraph.upload_device_order.__func__))
    for name in ("ring_signature_degree", "alloc_signature_table", "signature_transpose", "pairwise_l1",
                 "scatter_symmetric"):
        setattr(engine, name, timed(name, getattr(engine, name)))

    m = DynamicHSD(nx.barabasi_albert_graph(n, 5, seed=0), "ba", hop, 1, "wasserstein", signal="degree")
    out = {"n": n, "hop": hop, "inserted_edges": k_ins}

    def run(label):
        stages.clear()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.structural_distance_update()
        torch.cuda.synchronize()
        total = (time.perf_counter() - t0) * 1e3
        out[label] = {"total_ms": round(total, 2), "n_bins": int(m._dg.n_bins),
                      "affected": int(m.last_affected.numel()),
                      "mem_reserved_gb": round(torch.cuda.memory_reserved() / 1e9, 2),
                      **{k: round(v, 2) for k, v in stages.items()}}

    run("first")
    m._D = None
    m._sig_prev = None
    run("from_scratch")
    rng = np.random.default_rng(7)
    while True:
        u, v = (int(x) for x in rng.integers(0, n, 2))
        if u != v and not m.graph.has_edge(u, v):
            break
    m.dynamic_add_edges([(u, v)])
    run("warmup_1_edge")
    rng = np.random.default_rng(1)
    edges = set()
    while len(edges) < k_ins:
        u, v = (int(x) for x in rng.integers(0, n, 2))
        if u != v and not m.graph.has_edge(u, v):
            edges.add((min(u, v), max(u, v)))
    m.dynamic_add_edges(sorted(edges))
    run("update")
    run("repeat_no_change")      # same graph again: no signature differs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
