"""BASELINE.json's headline size (C2: BA 20 000 nodes, 3 hops, full N x N) through properties
that do not need the oracle to finish 2e8 pairs: symmetry, zero diagonal, the triangle
inequality (a sum of W1 metrics is a metric), agreement of the three product paths
(device, host pipeline, emulated sharding), a row-sum checksum, and a sampled comparison
against the oracle (scipy W1 over the reference's BFS rings)."""
import numpy as np
import pytest

from oracle import hsd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import powerlaw_graph
    g = powerlaw_graph(20000, 5, seed=0)
    dg = engine.DeviceGraph.upload(g)
    D, sizes = engine.degree_distance_device(dg, 3)
    torch.cuda.synchronize()
    return g, dg, D, sizes


def test_c2_structure(c2):
    import torch
    g, dg, D, sizes = c2
    n = g.n
    assert D.shape == (n, n) and D.dtype == torch.float32
    assert torch.equal(D, D.t())
    assert torch.all(torch.diagonal(D) == 0)
    assert torch.all(D >= 0) and torch.isfinite(D).all()
    # ring sizes: hop 0 is the node, hop 1 its degree
    assert torch.all(sizes[:, 0] == 1)
    assert np.array_equal(sizes[:, 1].cpu().numpy(), g.degree)
    assert dg.n_bins == 146 and dg.k_used(3) == 436           # B = 146 at C2


def test_c2_triangle_inequality_sampled(c2):
    import torch
    g, dg, D, _ = c2
    rng = np.random.default_rng(0)
    idx = torch.from_numpy(rng.integers(0, g.n, size=(200000, 3))).cuda()
    a, b, c = idx[:, 0], idx[:, 1], idx[:, 2]
    lhs = D[a, c]
    rhs = D[a, b] + D[b, c]
    assert torch.all(lhs <= rhs * (1 + 1e-5) + 1e-4)


def test_c2_sampled_pairs_match_oracle(c2):
    g, dg, D, sizes = c2
    rng = np.random.default_rng(1)
    rows = [0, 7, 19999]                       # a hub, an early node, the last node
    cols = sorted(set(rng.integers(0, g.n, size=150).tolist()) | {1, 2, 3})
    adj = [g.neighbors(i).astype(np.int64) for i in range(g.n)]
    ref = O.degree_distance_rows(adj, 3, rows, cols)
    got = D[rows][:, cols].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6 * ref.max())
    rings = O.all_rings(adj, 3, rows)
    for r in rows:
        assert sizes[r].tolist() == [len(l) for l in rings[r]]


def test_c2_paths_agree_bitwise(c2):
    import torch
    from hsd_b200 import engine
    from hsd_b200.sharded import ShardedDegreeHSD, shard_rows
    g, dg, D, _ = c2
    n = g.n
    checksum = D.double().sum(1)
    # host pipeline (pinned host buffers, panel-streamed D2H)
    pipe = engine.HostDegreePipeline(g, 3)
    out = torch.empty((n, n), dtype=torch.float32).pin_memory()
    pipe.run(out)
    assert torch.equal(out.cuda().double().sum(1), checksum)
    assert torch.equal(out[12345].cuda(), D[12345])
    del out
    # 8 emulated ranks with peer-memory mirroring
    world = 8
    per = shard_rows(n, world, 0)[2]
    blocks = [torch.zeros((per, engine.roundup(n, 4)), dtype=torch.float32, device="cuda") for _ in range(world)]
    plans = [ShardedDegreeHSD(dg, 3, r, world, peer=True, peer_blocks=blocks) for r in range(world)]
    for p in plans:
        p.signatures()
    for p in plans:
        for q in plans:
            sl = slice(q.rank * q.per, q.rank * q.per + q.n_src)
            p.sig_all[sl] = q.sig_all[sl]
    for p in plans:
        p.distances()
    torch.cuda.synchronize()
    got = torch.cat([b[:, :n] for b in blocks], 0)[:n]
    assert torch.equal(got, D)


def test_c3_sampled_pairs_match_oracle():
    """The north-star target size (BA 100 000 nodes, 4 hops; SURVEY C3: B = 248, K = 989): three
    rows of the device result against scipy W1 over the reference's BFS rings on sampled columns,
    plus the structural properties, without materialising more than the row block needed."""
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import powerlaw_graph
    g = powerlaw_graph(100000, 5, seed=0)
    dg = engine.DeviceGraph.upload(g)
    assert dg.n_bins == 248 and dg.k_used(4) == 989
    sig, sizes, _, status = engine.ring_signature_degree(dg, 4)
    assert int(status.item()) == 0
    k = dg.k_used(4)
    sigT = engine.alloc_signature_table(k, g.n, sig.device)
    engine.signature_transpose(sig, k, sigT)
    # rows 0..127 (one tile row, includes the biggest hubs) and the last 128 rows, against all columns
    top = engine.pairwise_l1(sigT, g.n, 0, 128, 0, g.n, symmetric=False, k_used=k)
    bot = engine.pairwise_l1(sigT, g.n, g.n - 128, 128, 0, g.n, symmetric=False, k_used=k)
    torch.cuda.synchronize()
    assert torch.equal(top[:, g.n - 128:], bot[:, :128].t())            # symmetry across the two blocks
    assert torch.all(torch.diagonal(top[:, :128]) == 0)
    rng = np.random.default_rng(2)
    cols = sorted(set(rng.integers(0, g.n, size=40).tolist()) | {0, 1, 99999})
    rows = [0, 5, 99999]
    adj = [g.neighbors(i).astype(np.int64) for i in range(g.n)]
    ref = O.degree_distance_rows(adj, 4, rows, cols)
    got = np.stack([top[0].cpu().numpy()[cols], top[5].cpu().numpy()[cols], bot[127].cpu().numpy()[cols]]).astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6 * ref.max())
    rings = O.all_rings(adj, 4, rows)
    for r in rows:
        assert sizes[r].tolist() == [len(l) for l in rings[r]]


def test_bfs_above_the_shared_memory_limit():
    """450 000 nodes: the four N-bit bitmaps of a source (225 KB) no longer fit shared memory, so
    the BFS kernel runs its global-workspace variant (persistent CTAs).  Ring sizes and degree
    signatures of sampled sources against the oracle."""
    import torch
    from hsd_b200 import engine
    from hsd_b200._lib import lib
    from hsd_b200.graph import powerlaw_graph
    n, hops = 450000, 2
    assert lib.hsd_bfs_workspace_words(n) > 0 and lib.hsd_bfs_workspace_words(100000) == 0
    g = powerlaw_graph(n, 3, seed=0)
    dg = engine.DeviceGraph.upload(g)
    rows = torch.tensor([0, 1, 17, 1000, 123456, n - 1] + list(range(5000, 5600)), dtype=torch.int32, device="cuda")
    sig, sizes, _, status = engine.ring_signature_degree(dg, hops, rows=rows)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    adj = [g.neighbors(i).astype(np.int64) for i in (0, 1, 17, 1000, 123456, n - 1)]
    deg = g.degree.astype(np.float64)
    sup = dg.support
    for a, src in enumerate((0, 1, 17, 1000, 123456, n - 1)):
        # hop 1 = neighbours, hop 2 = neighbours of neighbours not seen before (tools/hierarchy.py:25-38)
        ring1 = set(adj[a].tolist()) - {src}
        ring2 = set()
        for v in ring1:
            ring2.update(g.neighbors(v).tolist())
        ring2 -= ring1 | {src}
        assert sizes[a].tolist() == [1, len(ring1), len(ring2)]
        row = sig[a].cpu().numpy().astype(np.float64)
        assert row[0] == deg[src]
        for h, ring in enumerate((ring1, ring2)):
            d = deg[np.fromiter(ring, dtype=np.int64)]
            cdf = np.searchsorted(np.sort(d), sup[:-1], side="right") / len(d)
            want = cdf * np.diff(sup)
            got = row[1 + h * (len(sup) - 1):1 + (h + 1) * (len(sup) - 1)]
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-9)


# ----------------------------------------------------------------------------------------------
# BASELINE config 4 at size: 50 000 nodes, Chebyshev order 30, 4 scales (SURVEY §8 d "C4")
# ----------------------------------------------------------------------------------------------
def _sparse_laplacian(g):
    import scipy.sparse as sp
    rows = np.repeat(np.arange(g.n), np.diff(g.rowptr))
    A = sp.csr_matrix((np.ones(g.col.size), (rows, g.col)), shape=(g.n, g.n))
    return (sp.diags(np.asarray(A.sum(1)).ravel()) - A).tocsr()


@pytest.mark.parametrize("scale_rule", ["accurate", "reference"])
def test_c4_chebyshev_order30_4scales_at_50k_nodes(scale_rule):
    """hsd_cheb_spmm + hsd_ring_reduce at N = 50 000 / order 30 / 4 scales against the scipy.sparse
    float64 recurrence (oracle.cheby_apply = the pygsp restatement; lmax passed to both sides):
    64 sampled impulse columns of Psi_s and the [sum, mean] ring statistics of those nodes
    (model/multiscale_HSD.py:45-61).  'accurate': scales in [0.01, 40/lmax] (order 30 is a good
    approximation there); 'reference': MultiHSD.init's own formula up to 1.25*lmax — parity is
    against the SAME polynomial either way.  Tolerance: rtol 1e-5 plus the threshold
    floor 1e-4/N per ring member (a value within rounding of the threshold may flip)."""
    import networkx as nx
    import torch
    from hsd_b200 import wavelets as wv
    from hsd_b200.graph import powerlaw_graph
    from model import MultiHSD
    n, order, S, hop = 50000, 30, 4, 3
    g = powerlaw_graph(n, 5, seed=0)
    L = _sparse_laplacian(g)
    lmax = wv.estimate_lmax(g)
    assert 1.0 < lmax / (1.01 * g.degree.max()) <= 2.0       # d_max + 1 <= lambda_max <= 2 d_max
    if scale_rule == "accurate":
        scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / lmax), S))
    else:
        scales = O.multiscale_scales(lmax, S)
    thr = 1e-4 / n
    blocks = [(0, 16), (12344, 16), (31000, 16), (n - 16, 16)]     # ids 0..15 are the hubs of a BA graph
    cols = np.concatenate([np.arange(c0, c0 + c) for c0, c in blocks])
    csr = wv.DeviceCSR(g)
    coeffs = np.stack([wv.cheby_coefficients(float(s), lmax, order) for s in scales])
    for s_i, s in enumerate(scales):
        np.testing.assert_allclose(coeffs[s_i], O.cheby_coeff(float(s), lmax, order), rtol=1e-12, atol=1e-14)   # quadrature round-off on the ~1e-15 tail
    # oracle: un-thresholded responses of the 64 impulses, all scales
    E = np.zeros((n, cols.size))
    E[cols, np.arange(cols.size)] = 1.0
    ref = np.stack([O.cheby_apply(L, O.cheby_coeff(float(s), lmax, order), lmax, E).T for s in scales])   # [S, 64, N]
    got_raw, got_thr = [], []
    for c0, c in blocks:
        got_raw.append(wv.cheb_wavelet_block(csr, lmax, coeffs, c0, c, -np.inf).cpu().numpy())     # [S, N, c]
        got_thr.append(wv.cheb_wavelet_block(csr, lmax, coeffs, c0, c, thr).cpu().numpy())
    got_raw = np.concatenate(got_raw, axis=2).transpose(0, 2, 1)
    got_thr = np.concatenate(got_thr, axis=2).transpose(0, 2, 1)
    scale_mag = np.abs(ref).max(axis=(1, 2), keepdims=True)
    np.testing.assert_allclose(got_raw, ref, rtol=1e-5, atol=1e-12 * float(scale_mag.max()))
    ref_thr = np.where(ref > thr, ref, 0.0)
    flips = (got_thr != 0) != (ref_thr != 0)
    assert flips.sum() <= 4 and np.all(np.abs(ref[flips] - thr) < 1e-9 * thr + 1e-18)
    np.testing.assert_allclose(np.where(flips, ref_thr, got_thr), ref_thr, rtol=1e-5, atol=1e-12 * float(scale_mag.max()))
    # ring statistics of the sampled nodes through the model class (rings by the BFS kernel)
    m = MultiHSD(nx.barabasi_albert_graph(n, 5, seed=0), "ba50k", hop, S)
    assert np.array_equal(m.csr.col, g.col)
    m.lmax, m.scales, m.CHEB_ORDER = lmax, scales, order
    adj = [g.neighbors(i).astype(np.int64) for i in range(n)]
    for c0, c in blocks[:2] + blocks[3:]:
        emb = m.embed_device(col_range=(c0, c0 + c))[c0:c0 + c].cpu().numpy()          # [c, S, hop+1, 2]
        for a in range(0, c, 5):
            node = c0 + a
            rings = O.rings_of(adj, node, hop)
            row = ref_thr[:, int(np.nonzero(cols == node)[0][0]), :]                    # [S, N]
            for s_i in range(S):
                for h, ring in enumerate(rings):
                    vals = row[s_i, np.asarray(ring, dtype=np.int64)] if len(ring) else np.zeros(0)
                    want = [vals.sum(), vals.mean()] if len(ring) else [0.0, 0.0]
                    floor = thr * max(len(ring), 1)
                    np.testing.assert_allclose(emb[a, s_i, h], want, rtol=1e-5, atol=floor * 1e-3 + 1e-300)


# ----------------------------------------------------------------------------------------------
# BASELINE config 5 at size: 100 000 nodes, 1 % edge insertions (SURVEY §8 d "C5")
# ----------------------------------------------------------------------------------------------
def _insertions(G, k, rng):
    n = G.number_of_nodes()
    edges = set()
    while len(edges) < k:
        u, v = (int(x) for x in rng.integers(0, n, 2))
        if u != v and not G.has_edge(u, v):
            edges.add((min(u, v), max(u, v)))
    return sorted(edges)


@pytest.mark.parametrize("hop,batches", [(4, (5000,)), (2, (5, 5000))])
def test_c5_incremental_update_at_100k_nodes(hop, batches):
    """DynamicHSD on the C3 graph: after inserting edges drawn with numpy.random.default_rng(1)
    (5 000 = 1 % of the edges; hop 2 also a 5-edge batch, which takes the genuinely incremental
    path) the updated device matrix is BIT-equal to a from-scratch matrix of the edited graph, and
    three of its rows match scipy W1 over the reference's BFS rings on sampled columns."""
    import networkx as nx
    import torch
    from model import DynamicHSD, HSD
    n = 100000
    G = nx.barabasi_albert_graph(n, 5, seed=0)
    m = DynamicHSD(G, "ba100k", hop, 1, "wasserstein", signal="degree")
    m.structural_distance_update()
    rng = np.random.default_rng(1)
    for k_ins in batches:
        m.dynamic_add_edges(_insertions(m.graph, k_ins, rng))
        D = m.structural_distance_update()
        torch.cuda.synchronize()
        n_aff = int(m.last_affected.numel())
        assert 0 < n_aff <= n
        if k_ins == 5:
            assert n_aff < n // 2                      # really incremental: rows x all + mirrored scatter
        fresh = HSD(m.graph, "fresh", 0, hop, "wasserstein", signal="degree").structural_distance_device()
        torch.cuda.synchronize()
        for r0 in range(0, n, 20000):                  # blockwise: no 40 GB boolean temporary
            assert torch.equal(D[r0:r0 + 20000], fresh[r0:r0 + 20000])
        del fresh
    csr = m.csr
    adj = [csr.neighbors(i).astype(np.int64) for i in range(n)]
    aff = m.last_affected.cpu().numpy()
    rows = [int(aff[0]), int(aff[len(aff) // 2]), int(aff[-1])]
    cols = sorted(set(np.random.default_rng(3).integers(0, n, size=24).tolist()) | {0, n - 1})
    ref = O.degree_distance_rows(adj, hop, rows, cols)
    got = D[rows][:, cols].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6 * ref.max())
