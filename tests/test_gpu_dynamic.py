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


@pytest.mark.parametrize("n,k,hop", [(3000, 3, 2), (3000, 30, 3)])
def test_incremental_equals_from_scratch(n, k, hop):
    import torch
    from model import DynamicHSD, HSD
    g = _ba(n)
    m = DynamicHSD(g.copy(), "ba", hop, 1, "wasserstein", signal="degree")
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
