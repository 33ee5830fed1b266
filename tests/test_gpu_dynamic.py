"""Config 5: incremental update after edge insertions == from-scratch recompute."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ba(n, seed=0):
    import networkx as nx
    return nx.barabasi_albert_graph(n, 5, seed=seed)


def _insertions(g, k, seed=1):
    rng = np.random.default_rng(seed)
    n = g.number_of_nodes()
    out = set()
    while len(out) < k:
        u, v = (int(x) for x in rng.integers(0, n, 2))
        if u != v and not g.has_edge(u, v) and (min(u, v), max(u, v)) not in out:
            out.add((min(u, v), max(u, v)))
    return sorted(out)


@pytest.mark.parametrize("n,k,hop,update_rows", [(3000, 3, 2, None), (3000, 30, 3, None), (3000, 3, 2, 64)])
def test_incremental_equals_from_scratch(n, k, hop, update_rows):
    import torch
    from model import DynamicHSD, HSD
    g = _ba(n)
    m = DynamicHSD(g.copy(), "ba", hop, 1, "wasserstein", signal="degree")
    if update_rows:
        m.UPDATE_ROWS = update_rows       # several chunks of affected rows through the fixed-size workspace
    D0 = m.structural_distance_update().clone()
    new_edges = _insertions(g, k)
    m.dynamic_add_edges(new_edges)
    D1 = m.structural_distance_update()
    g2 = g.copy()
    g2.add_edges_from(new_edges)
    fresh = HSD(g2, "ba2", 0, hop, "wasserstein", signal="degree").structural_distance_device()
    assert torch.equal(D1, fresh)                      # bit-exact: same kernel, same signatures
    aff = m.last_affected.cpu().numpy()
    # every entry that changed lies in an affected row or column
    changed = (D0 != D1).cpu().numpy()
    mask = np.zeros(n, dtype=bool)
    mask[aff] = True
    assert not changed[~mask][:, ~mask].any()
    assert 0 < len(aff) <= n
    if update_rows:
        assert update_rows < len(aff) < n // 2           # the chunked incremental path, more than one chunk
    # a second update without an edit changes nothing and recomputes nothing
    assert torch.equal(m.structural_distance_update(), fresh) and m.last_affected.numel() == 0
    # the exact changed set is inside the analytic bound: nodes within `hop` hops of an endpoint
    m._pending.update(x for e in new_edges for x in e)
    ball = set(m.affected_nodes_device().cpu().numpy().tolist())
    m._pending.clear()
    assert set(aff.tolist()) <= ball


def test_add_node_grows_matrix():
    from model import DynamicHSD, HSD
    import torch
    g = _ba(500)
    m = DynamicHSD(g.copy(), "ba", 2, 1, "wasserstein", signal="degree")
    m.structural_distance_update()
    m.dynamic_add_node(500, [(500, 0), (500, 17), (500, 250)])
    D = m.structural_distance_update()
    assert D.shape == (501, 501)
    g2 = g.copy()
    g2.add_edges_from([(500, 0), (500, 17), (500, 250)])
    fresh = HSD(g2, "x", 0, 2, "wasserstein", signal="degree").structural_distance_device()
    assert torch.equal(D, fresh)


def test_explore_neighborhoods_and_subgraph():
    from model import DynamicHSD
    from oracle import hsd_oracle as O
    g = _ba(300)
    m = DynamicHSD(g, "ba", 3, 1, "wasserstein", signal="degree")
    seen = m.explore_neighborhoods(7)
    adj = [np.array(sorted(g.neighbors(v))) for v in range(300)]
    ref = O.rings_of(adj, 7, 3)
    assert m.hierarchy[7] == ref
    assert seen == set(x for layer in ref for x in layer)
    sub = m.convert_neighborhoods_to_subgraph(set(ref[0] + ref[1]))
    assert set(sub.nodes()) <= set(ref[0] + ref[1])
    assert all(g.has_edge(u, v) for u, v in sub.edges())


def _low_degree_insertions(g, k, seed=3):
    """Edges between nodes of degree <= 6: the set of distinct degrees (the shared support) stays
    the same, so the update takes the incremental path instead of a full recompute."""
    deg = dict(g.degree())
    low = [v for v in g if deg[v] <= 6]
    rng = np.random.default_rng(seed)
    out = set()
    while len(out) < k:
        u, v = (int(x) for x in rng.choice(low, 2, replace=False))
        if not g.has_edge(u, v):
            out.add((min(u, v), max(u, v)))
    return sorted(out)


@pytest.mark.parametrize("world,peer,mode", [(1, False, None), (2, False, None), (3, False, None),
                                             (2, True, "rows"), (4, True, "rows"), (3, True, "mirror")])
def test_sharded_incremental_update_equals_from_scratch(world, peer, mode):
    """Config 5 across ranks (SURVEY §8 e "Dynamic"), the ranks emulated in one process: after the
    insertions every rank's row block equals the rows of a from-scratch single-GPU matrix bit for
    bit, in the collective-free fallback (own rows x affected columns + affected own rows) and in
    both peer modes (affected rows dealt round-robin and stored whole into the owners' blocks, their
    columns either recomputed locally — "rows" — or stored mirrored — "mirror")."""
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import CSRGraph
    from hsd_b200.sharded import ShardedDegreeHSD, shard_rows
    n, hop = 3000, 2
    g0 = _ba(n)
    edges = _low_degree_insertions(g0, 4)
    g1 = g0.copy()
    g1.add_edges_from(edges)
    dg0 = engine.DeviceGraph.upload(CSRGraph.from_networkx(g0))
    dg1 = engine.DeviceGraph.upload(CSRGraph.from_networkx(g1))
    assert np.array_equal(dg0.support, dg1.support)
    fresh = ShardedDegreeHSD(dg1, hop, 0, 1).step().clone()
    before = ShardedDegreeHSD(dg0, hop, 0, 1).step().clone()

    per = shard_rows(n, world, 0)[2]
    ld = engine.roundup(n, 4)
    if peer:
        blocks = [torch.full((per, ld), float("nan"), dtype=torch.float32, device="cuda") for _ in range(world)]
        plans = [ShardedDegreeHSD(dg0, hop, r, world, peer=True, peer_blocks=blocks) for r in range(world)]
        for p in plans:
            p.update_mode = mode
    else:
        plans = [ShardedDegreeHSD(dg0, hop, r, world) for r in range(world)]

    def exchange():                       # what the all-gather / the fused peer stores do
        for p in plans:
            for q in plans:
                if p is not q:
                    sl = slice(q.rank * q.per, q.rank * q.per + q.n_src)
                    p.sig_all[sl] = q.sig_all[sl]

    def assembled():
        if peer:
            return torch.cat([b[:, :n] for b in blocks], 0)[:n]
        return torch.cat([p.out[:p.n_rows] for p in plans], 0)

    for p in plans:
        p.signatures()
    exchange()
    for p in plans:
        p.distances()
    torch.cuda.synchronize()
    assert torch.equal(assembled(), before)

    for p in plans:
        p.update_begin(dg1)
    exchange()
    affs = [p.update_finish()[1] for p in plans]
    torch.cuda.synchronize()
    m = int(affs[0].numel())
    assert all(torch.equal(a, affs[0]) for a in affs)
    assert 0 < m and 2 * m < n            # the incremental branch, not the full-matrix fallback
    assert torch.equal(assembled(), fresh)
    changed = (before != fresh).cpu().numpy()
    mask = np.zeros(n, dtype=bool)
    mask[affs[0].cpu().numpy()] = True
    assert not changed[~mask][:, ~mask].any()
    assert changed.any()


def test_dynamic_model_sharded_api_single_rank():
    """DynamicHSD.structural_distance_update_sharded on one rank: first call = full step, second =
    incremental update through the kept plan; both equal the unsharded method."""
    import torch
    from model import DynamicHSD
    g0 = _ba(2000)
    edges = _low_degree_insertions(g0, 3)
    a = DynamicHSD(g0.copy(), "ba", 2, 1, "wasserstein", signal="degree")
    b = DynamicHSD(g0.copy(), "ba", 2, 1, "wasserstein", signal="degree")
    assert torch.equal(a.structural_distance_update_sharded(0, 1), b.structural_distance_update())
    a.dynamic_add_edges(edges)
    b.dynamic_add_edges(edges)
    Da, Db = a.structural_distance_update_sharded(0, 1), b.structural_distance_update()
    assert torch.equal(Da, Db)
    assert torch.equal(torch.sort(a.last_affected).values, torch.sort(b.last_affected).values)
    assert 0 < a.last_affected.numel() < 1000


@pytest.mark.parametrize("n,include_zero", [(3000, False), (257, True), (5, False)])
def test_device_side_degree_order_equals_host_order(n, include_zero):
    """DeviceGraph.upload_device_order (torch sort / gather on the device) == DeviceGraph.upload (numpy
    on the host), field by field; the small graph has isolated nodes (degree 0 in the support)."""
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import CSRGraph
    if n > 5:
        g = CSRGraph.from_networkx(_ba(n, seed=4))
    else:
        g = CSRGraph.from_edges(5, np.array([[0, 1], [1, 2]]))      # nodes 3, 4 isolated
    a = engine.DeviceGraph.upload(g, include_zero=include_zero)
    b = engine.DeviceGraph.upload_device_order(g, include_zero=include_zero)
    torch.cuda.synchronize()
    assert a.n == b.n and a.n_bins == b.n_bins and np.array_equal(a.support, b.support)
    for f in ("rowptr", "col", "orig_of", "new_of", "bin_end", "delta"):
        x, y = getattr(a, f), getattr(b, f)
        assert x.dtype == y.dtype and x.shape == y.shape and torch.equal(x, y), f
