"""Kernel-level parity through the C-ABI on synthetic tables and random graphs: edge sizes
(N not a multiple of the tile, tiny K, rectangular and trapezoid ranges), ragged / empty
rings, disconnected graphs, self-loops."""
import numpy as np
import pytest

from oracle import hsd_oracle as O

pytestmark = pytest.mark.gpu


def _l1_ref(T, rows, cols):
    # integer-valued floats: every partial sum is exact in fp32, so equality is bit-exact
    return np.abs(T[:, rows][:, :, None] - T[:, cols][:, None, :]).sum(0)


@pytest.mark.parametrize("n,k", [(1, 1), (5, 3), (127, 16), (128, 17), (129, 33), (300, 100), (1000, 7)])
def test_pairwise_l1_symmetric_exact(n, k):
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n * 1000 + k)
    T = rng.integers(0, 50, size=(k, n)).astype(np.float32)
    sigT = engine.alloc_signature_table(k, n, "cuda")
    sigT[:k, :n] = torch.from_numpy(T).cuda()
    D = engine.pairwise_l1(sigT, n, symmetric=True, k_used=k).cpu().numpy()
    ref = _l1_ref(T.astype(np.float64), np.arange(n), np.arange(n))
    assert np.array_equal(D, ref)


@pytest.mark.parametrize("n,k,r0,nr,c0,nc", [(500, 40, 0, 500, 0, 500), (500, 40, 128, 200, 0, 500),
                                            (777, 9, 4, 129, 256, 300), (260, 5, 256, 4, 0, 260)])
def test_pairwise_l1_rectangles_exact(n, k, r0, nr, c0, nc):
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(7)
    T = rng.integers(-20, 20, size=(k, n)).astype(np.float32)
    sigT = engine.alloc_signature_table(k, n, "cuda")
    sigT[:k, :n] = torch.from_numpy(T).cuda()
    D = engine.pairwise_l1(sigT, n, r0, nr, c0, nc, symmetric=False).cpu().numpy()
    assert np.array_equal(D, _l1_ref(T.astype(np.float64), np.arange(r0, r0 + nr), np.arange(c0, c0 + nc)))


def test_pairwise_l1_trapezoid_panels_compose_the_matrix():
    import torch
    from hsd_b200 import engine
    from hsd_b200._lib import check, lib
    n, k = 700, 21
    rng = np.random.default_rng(3)
    T = rng.integers(0, 9, size=(k, n)).astype(np.float32)
    sigT = engine.alloc_signature_table(k, n, "cuda")
    sigT[:k, :n] = torch.from_numpy(T).cuda()
    D = torch.full((n, n), -1.0, dtype=torch.float32, device="cuda")
    for p0, pr in [(0, 128), (128, 256), (384, 316)]:
        view = D[p0:, p0:]
        check(lib.hsd_pairwise_l1(sigT.data_ptr(), k, sigT.stride(0), p0, pr, p0, n - p0, 1,
                                  view.data_ptr(), D.stride(0), torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(D.cpu().numpy(), _l1_ref(T.astype(np.float64), np.arange(n), np.arange(n)))


def test_signature_transpose_with_row_gather():
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(0)
    sig = torch.from_numpy(rng.random((300, 44)).astype(np.float32)).cuda()
    perm = torch.from_numpy(rng.permutation(300).astype(np.int32)).cuda()
    sigT = engine.alloc_signature_table(41, 300, "cuda")
    engine.signature_transpose(sig, 41, sigT, 0, src_rows=perm)
    assert torch.equal(sigT[:41, :300], sig[perm.long(), :41].t())
    assert torch.all(sigT[41:] == 0)


def _random_graph(rng, n, m, self_loops=0):
    from hsd_b200.graph import CSRGraph
    e = rng.integers(0, n, size=(m, 2))
    if self_loops:
        s = rng.integers(0, n, size=self_loops)
        e = np.concatenate([e, np.stack([s, s], 1)])
    return CSRGraph.from_edges(n, e)


@pytest.mark.parametrize("n,m,hops,loops", [(40, 30, 3, 0), (200, 150, 4, 5), (513, 2000, 2, 0), (64, 0, 2, 0), (90, 300, 6, 3)])
def test_bfs_rings_on_random_graphs(n, m, hops, loops):
    """Disconnected components, isolated nodes, self-loops, rings that run empty."""
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n + m)
    g = _random_graph(rng, n, m, loops)
    dg = engine.DeviceGraph.upload(g, include_zero=True)
    sig, sizes, bitmaps, status = engine.ring_signature_degree(dg, hops, want_bitmaps=True, empty="zero")
    bm = bitmaps.cpu().numpy().view(np.uint32)
    bits = np.unpackbits(bm.view(np.uint8), axis=-1, bitorder="little")[..., :n]
    orig = g.degree_order().orig_of
    adj = [g.neighbors(i).astype(np.int64) for i in range(n)]
    ref = O.all_rings(adj, hops)
    for i in range(n):
        got = [sorted(orig[np.nonzero(bits[i, h])[0]].tolist()) for h in range(hops + 1)]
        assert got == ref[i]
        assert sizes[i].tolist() == [len(l) for l in ref[i]]
    # signatures -> distances, empty rings as the point mass at 0
    k = dg.k_used(hops)
    sigT = engine.alloc_signature_table(k, n, sig.device)
    engine.signature_transpose(sig, k, sigT)
    D = engine.pairwise_l1(sigT, n, symmetric=True).cpu().numpy().astype(np.float64)
    rows = list(range(0, n, max(1, n // 12)))
    want = O.degree_distance_rows(adj, hops, rows, empty="zero")
    np.testing.assert_allclose(D[rows], want, rtol=1e-5, atol=1e-6 * max(want.max(), 1.0))


def test_value_mode_kernels_on_ragged_signals():
    """hsd_ring_signature_values + hsd_pairwise_w1_merge / _aligned against scipy on random
    signals (negative values, ties, rings of very different sizes)."""
    import torch
    from hsd_b200 import engine, rings
    rng = np.random.default_rng(11)
    g = _random_graph(rng, 60, 140)
    n = g.n
    dg = engine.DeviceGraph.upload(g)
    rs = rings.RingSet.bfs(dg, 2)
    psi = np.round(rng.standard_normal((n, n)), 1)          # ties on purpose
    psi_d = torch.from_numpy(psi).cuda()
    adj = [g.neighbors(i).astype(np.int64) for i in range(n)]
    ref_rings = O.all_rings(adj, 2)
    vals, offs = rings.sorted_ring_values(psi_d, rs)
    vals, offs = vals.cpu().numpy(), offs.cpu().numpy()
    for i in [0, 13, 59]:
        for h in range(3):
            want = np.sort(psi[i, ref_rings[i][h]])
            got = vals[offs[i * 3 + h]:offs[i * 3 + h + 1]]
            assert np.array_equal(got, want)
    empties = any(len(ref_rings[i][h]) == 0 for i in range(n) for h in range(3))
    Da = rings.value_distance(psi_d, rs, mode="aligned", metric="wasserstein").cpu().numpy()
    Dh = rings.value_distance(psi_d, rs, mode="aligned", metric="hellinger").cpu().numpy()
    Dg = rings.value_distance(psi_d, rs, mode="aligned", metric="wasserstein_guass").cpu().numpy()
    for i, j in [(0, 1), (5, 40), (22, 59)]:
        wa = sum(O.aligned_distance(list(psi[i, ref_rings[i][h]]), list(psi[j, ref_rings[j][h]]), "wasserstein") for h in range(3))
        wh = sum(O.aligned_distance(list(psi[i, ref_rings[i][h]]), list(psi[j, ref_rings[j][h]]), "hellinger") for h in range(3))
        assert Da[i, j] == pytest.approx(wa, rel=1e-12, abs=1e-14) and Da[j, i] == Da[i, j]
        assert Dh[i, j] == pytest.approx(wh, rel=1e-12, abs=1e-14)
        wg = sum(O.aligned_distance(list(psi[i, ref_rings[i][h]]), list(psi[j, ref_rings[j][h]]), "wasserstein_guass") for h in range(3))
        assert Dg[i, j] == pytest.approx(wg, rel=1e-10, abs=1e-12)
    if empties:
        with pytest.raises(ValueError):
            rings.value_distance(psi_d, rs, mode="w1")
    else:
        Dw = rings.value_distance(psi_d, rs, mode="w1").cpu().numpy()
        for i, j in [(0, 1), (5, 40), (22, 59)]:
            ww = sum(O.w1(psi[i, ref_rings[i][h]], psi[j, ref_rings[j][h]]) for h in range(3))
            assert Dw[i, j] == pytest.approx(ww, rel=1e-12, abs=1e-14)


@pytest.mark.parametrize("n_rows,n_cols,k", [(50, 50, 5), (300, 300, 20), (64, 1000, 64), (10, 7, 10), (200, 4097, 1)])
def test_topk_rows_matches_stable_argsort(n_rows, n_cols, k):
    """hsd_topk_rows == the k first entries of a stable argsort by (distance, column), self excluded;
    heavy ties on purpose (quantised values)."""
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n_rows * 7 + k)
    D = np.round(rng.random((n_rows, n_cols)) * 20).astype(np.float32) / 4.0     # many ties, >= 0
    idx, val = engine.topk_rows(torch.from_numpy(D).cuda(), k)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for r in range(n_rows):
        cand = [j for j in range(n_cols) if j != r]
        order = sorted(cand, key=lambda j: (D[r, j], j))[:k]
        want_i = order + [-1] * (k - len(order))
        assert idx[r].tolist() == want_i
        assert np.array_equal(val[r, :len(order)], D[r, order])
        assert np.all(np.isinf(val[r, len(order):]))


def _topk_expect(D, r, k, n_cols, self_col):
    order = sorted((j for j in range(n_cols) if j != self_col), key=lambda j: (D[r, j], j))[:k]
    return order + [-1] * (k - len(order))


@pytest.mark.parametrize("n_rows,n_alloc,n_cols,k,levels", [
    (40, 6000, 6000, 10, 2),      # ~2000 columns tie at the smallest value: > TK_CAND candidates -> radix select
    (40, 6000, 5998, 64, 3),      # same through the 16-byte path with a ragged tail
    (50, 1024, 1022, 7, 1000),    # 16-byte loads + 2 tail columns, few ties
    (33, 1003, 1003, 20, 1000),   # leading dimension not a multiple of 4: scalar loads
    (20, 20000, 20000, 20, 1 << 20),  # C2 row length, essentially no ties
    (6, 50000, 50000, 20, 1 << 20),   # long rows: 4-deep register buffers, the row is read once
    (4, 50000, 50000, 10, 5000),      # long rows with ~10-fold ties at every value
    (4, 50004, 50001, 64, 40),        # long rows, > 1024 ties at the bound: second pass, then radix select
    (3, 49999, 49999, 5, 1 << 20),    # long rows through the scalar-load path
])
def test_topk_rows_two_pass_and_fallback(n_rows, n_alloc, n_cols, k, levels):
    """The two-pass kernel (per-thread minima -> bound -> candidate list) and its radix fallback give
    the (distance, column)-ordered k smallest of every row; rows are nodes 5.. so the excluded
    self column is not the row index."""
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n_alloc + k)
    D = (np.floor(rng.random((n_rows, n_alloc)) * levels) / 8.0).astype(np.float32)
    idx, val = engine.topk_rows(torch.from_numpy(D).cuda(), k, self_col0=5, n_cols=n_cols)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    for r in range(n_rows):
        want = _topk_expect(D, r, k, n_cols, 5 + r)
        assert idx[r].tolist() == want
        assert np.array_equal(val[r], D[r, want])


def test_topk_rows_with_column_mask_and_model_api(golden_graphs):
    import torch
    from conftest import nx_graph
    from hsd_b200 import engine
    from model import HSD
    g = nx_graph(golden_graphs, "europe")
    m = HSD(g, "europe", 0, 3, "wasserstein", signal="degree")
    D = m.calculate_structural_distance(0.0).astype(np.float32)
    idx, val = m.nearest_neighbors(10)
    for r in [0, 100, 398]:
        order = sorted((j for j in range(m.n_node) if j != r), key=lambda j: (D[r, j], j))[:10]
        assert idx[r].tolist() == order and np.array_equal(val[r], D[r, order])
    # neighbours restricted to a "training fold" (even columns)
    allowed = np.zeros(((m.n_node + 31) // 32) * 32, dtype=np.uint8)
    allowed[0:m.n_node:2] = 1
    mask = torch.from_numpy(np.packbits(allowed, bitorder="little").view(np.int32).copy()).cuda()
    Dd = torch.from_numpy(D).cuda()
    idx2, _ = engine.topk_rows(Dd, 5, col_mask=mask)
    idx2 = idx2.cpu().numpy()
    for r in [1, 50, 397]:
        order = sorted((j for j in range(0, m.n_node, 2) if j != r), key=lambda j: (D[r, j], j))[:5]
        assert idx2[r].tolist() == order


@pytest.mark.parametrize("n,m,ld,mirror", [(70, 9, 72, True), (257, 33, 257, True), (64, 64, 64, True), (100, 1, 104, False)])
def test_scatter_symmetric_rows_and_mirrored_columns(n, m, ld, mirror):
    """hsd_scatter_symmetric: D[idx[a], c] = blk[a, c] and (mirror) D[c, idx[a]] = blk[a, c]; nothing
    else is touched (blk rows are rows of a symmetric matrix, as in the incremental update)."""
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n + m)
    S = rng.random((n, n)).astype(np.float32)
    S = S + S.T                                   # symmetric: doubly-affected entries agree
    idx = np.sort(rng.choice(n, size=m, replace=False)).astype(np.int64)
    buf = torch.full((n, ld), -7.0, dtype=torch.float32, device="cuda")
    D = buf[:, :n]
    blk_buf = torch.zeros((m, n + 3), dtype=torch.float32, device="cuda")
    blk_buf[:, :n] = torch.from_numpy(S[idx]).cuda()
    engine.scatter_symmetric(blk_buf[:, :n], torch.from_numpy(idx).cuda(), D, mirror=mirror)
    want = np.full((n, n), -7.0, dtype=np.float32)
    want[idx, :] = S[idx]
    if mirror:
        want[:, idx] = S[idx].T
    assert np.array_equal(D.cpu().numpy(), want)
    assert bool((buf[:, n:] == -7.0).all())


@pytest.mark.parametrize("name,k", [("europe", 10), ("barbell", 5)])
def test_topk_distances_match_sklearn_precomputed_knn(golden_runs, name, k):
    """The reference's consumer of the matrix is sklearn's precomputed-metric KNN
    (tools/evaluate.py:61-69).  On the REFERENCE's own distance matrix (golden fixture) the k neighbour
    distances of every node from hsd_topk_rows equal sklearn's kneighbors (self removed); indices are
    compared where the k-th and (k+1)-th distances differ (sklearn's tie order is unspecified)."""
    import torch
    from sklearn.neighbors import NearestNeighbors
    from hsd_b200 import engine
    if name == "europe":
        n = 399
        D = np.zeros((n, n))
        D[np.triu_indices(n, 1)] = golden_runs["europe_D"]
        D = D + D.T
    else:
        D = np.asarray(golden_runs[f"{name}_D"], dtype=np.float64)
        n = D.shape[0]
    D32 = D.astype(np.float32)
    D64 = D32.astype(np.float64)                       # both sides see the same float32-representable values
    idx, val = engine.topk_rows(torch.from_numpy(D32).cuda(), k)
    idx, val = idx.cpu().numpy(), val.cpu().numpy().astype(np.float64)
    dist, ind = NearestNeighbors(n_neighbors=k + 2, metric="precomputed").fit(D64).kneighbors(D64)
    for i in range(n):
        # remove ONE entry for the node itself (distance 0; other nodes may be at distance 0 too)
        d = dist[i].tolist()
        d.remove(0.0)
        assert val[i].tolist() == d[:k]
        if d[k - 1] < d[k]:                              # no tie across the cut: the neighbour SETS agree
            want = [j for j in ind[i].tolist() if j != i][:k]
            if len(want) == k and i in ind[i]:
                assert sorted(idx[i].tolist()) == sorted(want)


@pytest.mark.parametrize("n_cols,levels", [(40000, 1 << 20), (36001, 50)])
def test_topk_long_rows_with_column_mask(n_cols, levels):
    """Long rows (4-deep register buffers) together with a column bitmap: only allowed columns may be
    neighbours; the tie-heavy case goes through the second pass and the radix select."""
    import torch
    from hsd_b200 import engine
    rng = np.random.default_rng(n_cols)
    n_rows, k = 3, 12
    D = (np.floor(rng.random((n_rows, n_cols)) * levels) / 8.0).astype(np.float32)
    allowed = np.zeros(((n_cols + 31) // 32) * 32, dtype=np.uint8)
    allowed[:n_cols] = rng.random(n_cols) < 0.3
    allowed[[0, 1, 2]] = 1                      # the rows' own columns are allowed but still excluded
    mask = torch.from_numpy(np.packbits(allowed, bitorder="little").view(np.int32).copy()).cuda()
    idx, val = engine.topk_rows(torch.from_numpy(D).cuda(), k, col_mask=mask)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    cols = np.nonzero(allowed[:n_cols])[0]
    for r in range(n_rows):
        order = sorted((int(j) for j in cols if j != r), key=lambda j: (D[r, j], j))[:k]
        assert idx[r].tolist() == order
        assert np.array_equal(val[r], D[r, order])


@pytest.mark.parametrize("n", [1, 63, 130, 399])
def test_exact_wavelets_fused_kernel(n):
    """hsd_exact_wavelets = U diag(exp(-s lambda)) U^T with the threshold of model/HSD.py:65 fused,
    against the reference's own two np.dot products (model/HSD.py:61-63) on the host."""
    import torch
    from hsd_b200 import wavelets as wv
    rng = np.random.default_rng(n)
    A = (rng.random((n, n)) < 0.1).astype(np.float64)
    A = np.triu(A, 1)
    A = A + A.T
    L = np.diag(A.sum(1)) - A
    lam, U = np.linalg.eigh(L)
    for scale, coeff in ((0.7, 1e-4), (3.0, None)):
        ref = np.dot(np.dot(U, np.diag(np.exp(-1 * scale * lam))), np.transpose(U))
        eig = (torch.from_numpy(lam).cuda(), torch.from_numpy(U).cuda())
        got = wv.exact_wavelets_dense(torch.from_numpy(L).cuda(), scale, coeff, eig).cpu().numpy()
        assert np.array_equal(got, got.T)
        if coeff is not None:
            thr = coeff / n
            keep = np.abs(ref - thr) > 1e-12          # entries within rounding of the threshold may flip
            ref = np.where(ref > thr, ref, 0.0)
            np.testing.assert_allclose(got[keep], ref[keep], rtol=1e-11, atol=1e-14)
        else:
            np.testing.assert_allclose(got, ref, rtol=1e-11, atol=1e-14)
