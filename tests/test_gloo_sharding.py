"""The N>1 path on CPU: world_size-2 gloo run of the sharding plumbing (row blocks,
the in-place all-gather of the signature table, max-over-ranks timing)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, ld, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hsd_b200.sharded import shard_rows
    row0, n_rows, per = shard_rows(n, world, rank)
    # the table a rank would fill with its own signatures: row i, column k -> f(i, k)
    table = torch.zeros((world * per, ld), dtype=torch.float32)
    i = torch.arange(row0, row0 + n_rows, dtype=torch.float32)[:, None]
    k = torch.arange(ld, dtype=torch.float32)[None, :]
    table[row0:row0 + n_rows] = i * 1000 + k
    chunk = table[rank * per:(rank + 1) * per]
    dist.all_gather_into_tensor(table, chunk)            # same call ShardedDegreeHSD.gather makes
    want = torch.arange(n, dtype=torch.float32)[:, None] * 1000 + k
    ok = bool(torch.equal(table[:n], want)) and bool(torch.all(table[n:] == 0))
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)  # "elapsed ms" differs per rank
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ret[rank] = (ok, float(t.item()), row0, n_rows)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1190, 7])
def test_world2_allgather_of_signature_table(n):
    world, ld = 2, 12
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, ld, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r][0] for r in range(world))
    assert all(ret[r][1] == float(world) for r in range(world))         # max over ranks
    assert ret[0][2] == 0 and ret[0][3] + ret[1][3] == n and ret[1][2] == ret[0][3]


def _update_worker(rank, world, port, n, aff_list, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hsd_b200.sharded import deal_affected, shard_rows
    aff = torch.tensor(aff_list, dtype=torch.int64)          # identical on every rank (replicated tables)
    dealt, segments = deal_affected(aff, n, world, rank)
    # what ShardedDegreeHSD.update_finish stores through peer memory: whole rows to their owners
    sends = [(owner, (dealt[lo:hi] - r0).tolist(), dealt[lo:hi].tolist()) for owner, lo, hi, r0 in segments]
    gathered = [None] * world
    dist.all_gather_object(gathered, sends)
    row0, n_rows, _ = shard_rows(n, world, rank)
    received = sorted(g for src in gathered for owner, local, glob in src if owner == rank for g in glob)
    local_ok = all(0 <= l < n_rows and l + row0 == g
                   for src in gathered for owner, local, glob in src if owner == rank for l, g in zip(local, glob))
    mine = [a for a in aff_list if row0 <= a < row0 + n_rows]
    ret[rank] = (received == mine, local_ok, int(dealt.numel()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n,aff", [(1190, list(range(3, 1190, 7))), (50, [0, 1, 2, 3]), (64, [63]), (40, [])])
def test_world2_incremental_update_deals_every_affected_row_to_its_owner(n, aff):
    """Config 5 across ranks (SURVEY §8 e): the affected rows are dealt round-robin; every one must
    reach the rank that owns it exactly once, addressed by its local row index."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_update_worker, args=(r, world, port, n, aff, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r][0] and ret[r][1] for r in range(world))
    assert sum(ret[r][2] for r in range(world)) == len(aff)
    assert abs(ret[0][2] - ret[1][2]) <= 1                 # balanced whatever block the rows fall into


def _status_worker(rank, world, port, flags, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hsd_b200.sharded import agree_status
    status = torch.tensor([flags[rank]], dtype=torch.int32)      # what the BFS kernel left on this rank
    ret[rank] = agree_status(status, world)
    dist.destroy_process_group()


@pytest.mark.parametrize("flags,want", [((1, 0), 1), ((0, 0), 0), ((0, 1), 1), ((2, 1), 3)])
def test_world2_empty_ring_flag_is_global(flags, want):
    """an isolated node is BFS-ed by ONE rank; every rank must see its empty-ring flag so
    that all of them raise (or none), instead of one raising and the others hanging at a barrier."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_status_worker, args=(r, world, port, flags, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] == want and ret[1] == want
