"""NVLink evidence for the two fused compute+exchange kernels (VERDICT r1 #4c), as ONE process that
drives the kernels of rank 0 of a `world`-rank job on cuda:0 while the other ranks' tables / result
blocks live on cuda:1..world-1 (peer access enabled): a plain single-process program ncu can profile
(`nvltx__bytes*`, `nvlrx__bytes*` of the launches on cuda:0).  Run under `gpurun --gpus W`.

    python scripts/nvlink_ncu_probe.py [N_NODES] [HOPS] [WORLD]
prints the algorithmic peer bytes of rank 0's two launches for comparison with the counters."""
import ctypes, sys
sys.path.insert(0, ".")
import torch
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
from hsd_b200.sharded import ShardedDegreeHSD, shard_rows

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world = int(sys.argv[3]) if len(sys.argv) > 3 else torch.cuda.device_count()
assert torch.cuda.device_count() >= world >= 2
torch.cuda.set_device(0)
rt = ctypes.CDLL("libcudart.so.12")
for peer in range(1, world):
    rc = rt.cudaDeviceEnablePeerAccess(peer, 0)
    assert rc in (0, 704), f"cudaDeviceEnablePeerAccess({peer}) -> {rc}"     # 704: already enabled
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g, device=torch.device("cuda", 0))
per = shard_rows(n, world, 0)[2]
ld_out = engine.roundup(n, 4)
k = dg.k_used(hops)
ld = engine.roundup(k, 4)
tables = [torch.zeros((world * per, ld), dtype=torch.float32, device=f"cuda:{r}") for r in range(world)]
blocks = [torch.zeros((per, ld_out), dtype=torch.float32, device=f"cuda:{r}") for r in range(world)]
plan = ShardedDegreeHSD(dg, hops, 0, world, peer=True, peer_blocks=blocks, peer_tables=tables)
for it in range(3):
    plan.signatures()          # rank 0's BFS sources: rows stored into its own table AND the peers' (NVLink)
    torch.cuda.synchronize()
    # rank 0 needs the whole table for its pairwise tiles: fill in the other ranks' rows locally (they would have
    # arrived over NVLink from the peers' BFS kernels)
    if it == 0:
        full = ShardedDegreeHSD(dg, hops, 0, 1)
        full.signatures()
        torch.cuda.synchronize()
        src = full.sig_all[:n]
        plan.sig_all[plan.table_row.long()] = src
    plan.distances()           # rank 0's share of the symmetric tiles, stored direct + mirrored into the owners' blocks
    torch.cuda.synchronize()
own_rows = plan.n_src
sig_bytes = own_rows * (k * 4) * (world - 1)            # every signature row goes to world-1 peers (k_used floats of it)
tiles = (n + 127) // 128
# rank 0 computes tiles 0, world, 2*world, ... of the upper triangle; count entries that land in a peer's block
import numpy as np
T = tiles * (tiles + 1) // 2
t = np.arange(0, T, world)
b = 2.0 * tiles + 1
I = np.floor((b - np.sqrt(np.maximum(b * b - 8.0 * t, 0))) / 2).astype(np.int64)
first = lambda i: i * tiles - i * (i - 1) // 2
I = np.where(first(I) > t, I - 1, I); I = np.where(first(I + 1) <= t, I + 1, I)
J = I + (t - first(I))
def rows_in(tile_idx, lo, hi):
    a = np.clip(np.minimum(tile_idx * 128 + 128, n) - np.maximum(tile_idx * 128, lo), 0, None)
    a2 = np.clip(np.minimum(np.minimum(tile_idx * 128 + 128, n), hi) - np.maximum(tile_idx * 128, lo), 0, None)
    return a2
own_lo, own_hi = 0, min(per, n)
h = lambda ti: np.minimum(ti * 128 + 128, n) - ti * 128
remote_direct = ((h(I) - rows_in(I, own_lo, own_hi)) * h(J)).sum()
remote_mirror = (np.where(I != J, (h(J) - rows_in(J, own_lo, own_hi)) * h(I), 0)).sum()
print(f"graph n={n} hops={hops} world={world} k_used={k}")
print(f"algorithmic peer bytes of rank 0's launches: BFS fused all-gather {sig_bytes/1e6:.2f} MB; "
      f"pairwise direct {remote_direct*4/1e6:.2f} MB + mirrored {remote_mirror*4/1e6:.2f} MB = {(remote_direct+remote_mirror)*4/1e6:.2f} MB")
print("peer block checksum", float(blocks[1].double().sum()), "peer table checksum", float(tables[1].double().sum()))
