"""Row-block sharding on the GPU: the ranks of a 2- and 3-way split are emulated in one
process (their signature slices exchanged by plain copies, which is what the all-gather
does) and must reproduce the single-GPU symmetric result bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_emulated_ranks_equal_single_gpu(world):
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import powerlaw_graph
    from hsd_b200.sharded import ShardedDegreeHSD
    g = powerlaw_graph(2500, 5, seed=0)
    dg = engine.DeviceGraph.upload(g)
    single = ShardedDegreeHSD(dg, 3, 0, 1).step().clone()
    plans = [ShardedDegreeHSD(dg, 3, r, world) for r in range(world)]
    for p in plans:
        p.signatures()
    for p in plans:                       # emulate the in-place all-gather
        for q in plans:
            sl = slice(q.rank * q.per, q.rank * q.per + q.n_src)
            p.sig_all[sl] = q.sig_all[sl]
    blocks = [p.distances() for p in plans]
    torch.cuda.synchronize()
    assert torch.equal(torch.cat(blocks, 0), single)
    assert sum(p.n_rows for p in plans) == g.n


@pytest.mark.parametrize("world,n", [(2, 2500), (4, 1801)])
def test_emulated_peer_memory_symmetric_tiles(world, n):
    """hsd_pairwise_l1_sharded with the peers' blocks standing in as local buffers: every
    rank's launch writes direct + mirrored tiles into the owners' blocks; after all launches
    the blocks hold exactly the single-GPU matrix."""
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import powerlaw_graph
    from hsd_b200.sharded import ShardedDegreeHSD, shard_rows
    g = powerlaw_graph(n, 5, seed=0)
    dg = engine.DeviceGraph.upload(g)
    single = ShardedDegreeHSD(dg, 3, 0, 1).step().clone()
    per = shard_rows(n, world, 0)[2]
    ld = engine.roundup(n, 4)
    blocks = [torch.full((per, ld), float("nan"), dtype=torch.float32, device="cuda") for _ in range(world)]
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
    assert torch.equal(got, single)
    # fused all-gather: the BFS kernel of every "rank" stores its rows into all tables itself
    k_ld = plans[0].ld
    tables = [torch.zeros((world * per, k_ld), dtype=torch.float32, device="cuda") for _ in range(world)]
    blocks2 = [torch.full((per, ld), float("nan"), dtype=torch.float32, device="cuda") for _ in range(world)]
    fused = [ShardedDegreeHSD(dg, 3, r, world, peer=True, peer_blocks=blocks2, peer_tables=tables)
             for r in range(world)]
    for p in fused:
        p.signatures()
    torch.cuda.synchronize()
    for t in tables[1:]:
        assert torch.equal(t, tables[0])
    assert torch.equal(tables[0], plans[0].sig_all)
    for p in fused:
        p.distances()
    torch.cuda.synchronize()
    assert torch.equal(torch.cat([b[:, :n] for b in blocks2], 0)[:n], single)


def test_host_pipeline_matches_device_path():
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import powerlaw_graph
    g = powerlaw_graph(3000, 5, seed=1)
    dg = engine.DeviceGraph.upload(g)
    D, _ = engine.degree_distance_device(dg, 3)
    pipe = engine.HostDegreePipeline(g, 3, n_chunks=5)
    out = torch.empty((g.n, g.n), dtype=torch.float32).pin_memory()
    pipe.run(out)
    assert torch.equal(out, D.cpu())
    assert len(pipe.panels) > 1
    # a row shard through the same pipeline (what a rank of an N>1 e2e run does)
    shard = engine.HostDegreePipeline(g, 3, row0=1000, n_rows=700, n_chunks=3)
    out2 = torch.empty((700, g.n), dtype=torch.float32).pin_memory()
    shard.run(out2)
    assert torch.equal(out2, D[1000:1700].cpu())


def test_model_out_buffer_uses_host_pipeline(golden_graphs):
    import torch
    from conftest import nx_graph
    from model import HSD
    g = nx_graph(golden_graphs, "europe")
    m = HSD(g, "europe", 0, 3, "wasserstein", signal="degree")
    ref = m.calculate_structural_distance(0.0)
    out = torch.empty((m.n_node, m.n_node), dtype=torch.float32).pin_memory()
    m.calculate_structural_distance(0.0, out=out)
    np.testing.assert_array_equal(out.numpy().astype(np.float64), ref)


def test_multihsd_column_shards_compose_the_embedding(golden_graphs):
    """Multi-GPU MultiHSD: ranks own contiguous blocks of impulse columns; emulated here by
    computing each block separately on one GPU and summing the (disjoint) row blocks."""
    import torch
    from conftest import nx_graph
    from model import MultiHSD
    g = nx_graph(golden_graphs, "europe")
    m = MultiHSD(g, "europe", 2, 3)
    full = m.embed_device()
    n = m.n_node
    world = 3
    per = ((n + world - 1) // world + 1) // 2 * 2
    acc = torch.zeros_like(full)
    for r in range(world):
        beg, end = min(r * per, n), min((r + 1) * per, n)
        part = m.embed_device(col_range=(beg, end))
        assert torch.all(part[:beg] == 0) and torch.all(part[end:] == 0)
        acc += part
    # same kernels, same per-column arithmetic; column blocks differ only in how the SpMM is batched
    assert torch.allclose(acc, full, rtol=1e-12, atol=1e-15)
    single = m.embed_device_sharded(0, 1)
    assert torch.equal(single, full)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("n,hops", [(2500, 3), (4100, 2), (900, 4)])
def test_column_split_dense_ring_variant(world, n, hops):
    """hsd_ring_counts_dense_cols + hsd_ring_signature_from_counts: every emulated rank runs the bitmap
    recursion for all nodes on its own 1/world of the bitmap columns; the partial integer counts summed
    over the ranks (what the all-reduce does) give signatures, ring sizes and the empty-ring flag
    bit-identical to the single-GPU kernels — with isolated nodes and both empty-ring policies."""
    import torch
    from hsd_b200 import engine
    from hsd_b200.graph import CSRGraph, powerlaw_graph
    g0 = powerlaw_graph(n, 4, seed=world)
    rows = np.repeat(np.arange(g0.n), np.diff(g0.rowptr))
    e = np.stack([rows, g0.col], 1)
    g = CSRGraph.from_edges(n + 2, e[rows < g0.col])              # two isolated nodes
    dg = engine.DeviceGraph.upload(g, include_zero=True)
    ref_sig, ref_sizes, _, ref_status = engine.ring_signature_degree(dg, hops, empty="zero")
    total = None
    for r in range(world):
        c = engine.ring_counts_cols(dg, hops, r, world).clone()
        total = c if total is None else total + c
    src = dg.new_of.contiguous()
    out_rows = torch.arange(g.n, dtype=torch.int32, device="cuda")
    for empty in ("zero", "raise"):
        sig = torch.zeros_like(ref_sig)
        sizes = torch.zeros_like(ref_sizes)
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        engine.signature_from_counts(dg, hops, total, src, out_rows, sig, sizes, empty, status)
        torch.cuda.synchronize()
        assert torch.equal(sizes, ref_sizes)
        if empty == "zero":
            assert torch.equal(sig, ref_sig) and int(status.item()) == 0
        else:
            assert int(status.item()) & 1                         # the isolated nodes' rings are empty
    assert engine.ring_cols_range(g.n, world - 1, world)[1] == ((g.n + 31) // 32 + 3) // 4
