"""Timing of the other BASELINE.json configs (C1, C3, C4, C5) on one GPU -> JSON lines.
Dev/measurement aid; bench.py stays on C2 as the contract says."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch, networkx as nx
from hsd_b200 import engine, wavelets as wv
from hsd_b200.graph import CSRGraph, powerlaw_graph

def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        a, b = ev(), ev(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best, r

which = sys.argv[1:] or ["c1", "c3", "c4", "c5"]
peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0}

if "c1" in which:
    g = np.load("tests/golden/graphs.npz")
    for name in ["europe", "usa"]:
        nodes = [str(v) for v in g[f"{name}_nodes"]]
        G = nx.Graph(); G.add_nodes_from(nodes); G.add_edges_from((nodes[u], nodes[v]) for u, v in g[f"{name}_edges"])
        from model import HSD
        m = HSD(G, name, 0, 3, "wasserstein")
        m.structural_distance_device(1.0, approx=False)      # warm-up (cuSOLVER init)
        t0 = time.perf_counter(); D = m.calculate_structural_distance(1.0, approx=False); t = time.perf_counter() - t0
        n = m.n_node
        print(json.dumps({"config": "C1", "graph": name, "n": n, "hops": 3, "signal": "wavelet (exact, FP64)",
                          "wall_s": t, "pairs_per_s": n * (n - 1) / 2 / t, "checksum": float(D.sum()),
                          "reference_cpu_s": {"europe": 18.1, "usa": 256.2}[name]}))

if "c3" in which:
    g = powerlaw_graph(100000, 5, seed=0); dg = engine.DeviceGraph.upload(g); n = g.n
    out = torch.empty((n, n), dtype=torch.float32, device="cuda")
    k = dg.k_used(4)
    t_b, (sig, sizes, _, st) = timed(lambda: engine.ring_signature_degree(dg, 4))
    sigT = engine.alloc_signature_table(k, n, sig.device); engine.signature_transpose(sig, k, sigT)
    t_p, _ = timed(lambda: engine.pairwise_l1(sigT, n, symmetric=True, out=out, k_used=k))
    peak = engine.fp32_issue_peak()
    pairs = n * (n - 1) / 2
    print(json.dumps({"config": "C3 on 1 GPU", "n": n, "hops": 4, "K": k, "bfs_ms": t_b, "pairwise_ms": t_p,
                      "pairs_per_s": pairs / ((t_b + t_p) * 1e-3), "pairwise_tflops": 2 * pairs * k / (t_p * 1e-3) / 1e12,
                      "fp32_peak_tflops": peak / 1e12, "frac": 2 * pairs * k / (t_p * 1e-3) / peak}))
    del out, sigT, sig

if "c4" in which:
    n, order, S = 50000, 30, 4
    g = powerlaw_graph(n, 5, seed=0)
    lmax = wv.estimate_lmax(g)
    csr = wv.DeviceCSR(g)
    scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / lmax), S))
    coeffs = np.stack([wv.cheby_coefficients(s, lmax, order) for s in scales])
    for C in [256, 1024, 4096]:
        work = torch.empty((3, n, C), dtype=torch.float64, device="cuda"); outb = torch.empty((S, n, C), dtype=torch.float64, device="cuda")
        t, _ = timed(lambda: wv.cheb_wavelet_block(csr, lmax, coeffs, 0, C, 1e-4 / n, work, outb))
        byts = order * (8.0 * g.nnz + 4 * (n + 1) + (3 + 2 * S) * 8.0 * n * C)
        print(json.dumps({"config": "C4 cheb_spmm block", "n": n, "order": order, "scales": S, "cols": C, "ms": t,
                          "alg_GBs": byts / (t * 1e-3) / 1e9, "hbm_peak": peaks["hbm_gbs"], "frac": byts / (t * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "full_job_s_extrapolated": t * 1e-3 * n / C}))
        del work, outb
    # whole MultiHSD.embed (rings + SpMM + ring reduce) on a 20k-node graph for an end-to-end number
    G = nx.barabasi_albert_graph(20000, 5, seed=0)
    from model import MultiHSD
    m = MultiHSD(G, "ba20k", 3, 4)
    m.scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / m.lmax), 4)); m.CHEB_ORDER = 30
    t0 = time.perf_counter(); emb = m.embed_device(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    print(json.dumps({"config": "C4-like MultiHSD.embed_device", "n": 20000, "hop": 3, "order": 30, "scales": 4, "wall_s": t,
                      "checksum": float(emb.sum())}))

if "c5" in which:
    n, hop = 100000, 4
    G = nx.barabasi_albert_graph(n, 5, seed=0)
    from model import DynamicHSD
    m = DynamicHSD(G, "ba100k", hop, 1, "wasserstein", signal="degree")
    t0 = time.perf_counter(); m.structural_distance_update(); torch.cuda.synchronize(); t_full = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    for k_ins in [50, 5000]:
        edges = set()
        while len(edges) < k_ins:
            u, v = (int(x) for x in rng.integers(0, n, 2))
            if u != v and not m.graph.has_edge(u, v): edges.add((min(u, v), max(u, v)))
        m.dynamic_add_edges(sorted(edges))
        t0 = time.perf_counter(); D = m.structural_distance_update(); torch.cuda.synchronize(); t_inc = time.perf_counter() - t0
        print(json.dumps({"config": "C5", "n": n, "hop": hop, "inserted_edges": k_ins, "affected_rows": int(m.last_affected.numel()),
                          "full_s": t_full, "incremental_s": t_inc}))

if "c4full" in which:
    n, order, S, hop = 50000, 30, 4, 3
    G = nx.barabasi_albert_graph(n, 5, seed=0)
    from model import MultiHSD
    t0 = time.perf_counter(); m = MultiHSD(G, "ba50k", hop, S); t_init = time.perf_counter() - t0
    m.scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / m.lmax), S)); m.CHEB_ORDER = order
    m._rings(); torch.cuda.synchronize()
    t0 = time.perf_counter(); emb = m.embed_device(); torch.cuda.synchronize(); t = time.perf_counter() - t0
    byts = order * (8.0 * m.csr.nnz + 4 * (n + 1) + (3 + 2 * S) * 8.0 * n * n)
    print(json.dumps({"config": "C4 full: MultiHSD.embed_device", "n": n, "hop": hop, "order": order, "scales": S,
                      "init_s": t_init, "embed_s": t, "alg_GBs_incl_ring_reduce": byts / t / 1e9, "hbm_peak": peaks["hbm_gbs"],
                      "checksum": float(emb.sum()), "emb_shape": list(emb.shape)}))

if "c5h2" in which:
    n, hop = 100000, 2
    G = nx.barabasi_albert_graph(n, 5, seed=0)
    from model import DynamicHSD
    m = DynamicHSD(G, "ba100k", hop, 1, "wasserstein", signal="degree")
    m.structural_distance_update(); torch.cuda.synchronize()          # warm-up (lazy inits)
    m._D = None
    t0 = time.perf_counter(); m.structural_distance_update(); torch.cuda.synchronize(); t_full = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    for k_ins in [5, 50, 1000]:
        edges = set()
        while len(edges) < k_ins:
            u, v = (int(x) for x in rng.integers(0, n, 2))
            if u != v and not m.graph.has_edge(u, v): edges.add((min(u, v), max(u, v)))
        t0 = time.perf_counter(); m.dynamic_add_edges(sorted(edges)); t_edit = time.perf_counter() - t0
        t0 = time.perf_counter(); D = m.structural_distance_update(); torch.cuda.synchronize(); t_inc = time.perf_counter() - t0
        print(json.dumps({"config": "C5 variant hop=2", "n": n, "hop": hop, "inserted_edges": k_ins,
                          "affected_rows": int(m.last_affected.numel()), "full_s": t_full,
                          "host_edit_s": t_edit, "incremental_s": t_inc}))

if "topk" in which:
    for n, hops in [(20000, 3), (100000, 4)]:
        g = powerlaw_graph(n, 5, seed=0); dg = engine.DeviceGraph.upload(g)
        D, _ = engine.degree_distance_device(dg, hops)
        t, (idx, val) = timed(lambda: engine.topk_rows(D, 20))
        byts = 4.0 * n * n
        print(json.dumps({"config": "k-NN top-20 on the resident matrix", "n": n, "ms": t, "matrix_GB": byts / 1e9,
                          "matrix_reads_GBs": byts / (t * 1e-3) / 1e9, "hbm_peak": peaks["hbm_gbs"],
                          "d2h_bytes_instead_of_matrix": int(idx.numel() * 8)}))
        del D
