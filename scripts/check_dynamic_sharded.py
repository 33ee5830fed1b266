"""torchrun check (config 5 on G GPUs): after edge insertions every rank's row block from
DynamicHSD.structural_distance_update_sharded equals the rows of a from-scratch single-GPU matrix,
in peer mode (affected rows dealt round-robin, stores through NVLink peer memory) and in the
collective-free fallback; prints the affected-row count and the update / full-step times."""
import os, sys, time
sys.path.insert(0, ".")
import numpy as np, networkx as nx
import torch, torch.distributed as dist
from model import DynamicHSD, HSD
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k_ins = int(sys.argv[3]) if len(sys.argv) > 3 else 5
g0 = nx.barabasi_albert_graph(n, 5, seed=0)
deg = dict(g0.degree())
low = [v for v in g0 if deg[v] <= 6]
rng = np.random.default_rng(3)
def draw(g, k):
    out = set()
    while len(out) < k:
        u, v = (int(x) for x in rng.choice(low, 2, replace=False))
        if not g.has_edge(u, v):
            out.add((min(u, v), max(u, v)))
    return sorted(out)
batch1 = draw(g0, k_ins)                      # first update also pays one-time costs (lazy kernel loads, peer views)
g1 = g0.copy(); g1.add_edges_from(batch1)
batch2 = draw(g1, k_ins)                      # second update is the timed one
if len(sys.argv) > 4 and sys.argv[4] == "newdeg":
    # one more neighbour for the biggest hub: a NEW distinct degree, every signature changes length and the
    # plan is rebuilt on the previous plan's symmetric-memory allocations (ShardedDegreeHSD(reuse=...))
    hub = max(deg, key=deg.get)
    batch2 = batch2 + [(min(hub, x), max(hub, x)) for x in low if not g1.has_edge(hub, x)][:1]
g2 = g1.copy(); g2.add_edges_from(batch2)
fresh = HSD(g2, "fresh", 0, hop, "wasserstein", signal="degree").structural_distance_device()
for peer in (False, True):
    m = DynamicHSD(g0.copy(), "ba", hop, 1, "wasserstein", signal="degree")
    m.structural_distance_update_sharded(rank, world, peer=peer)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); m._plan.step(); torch.cuda.synchronize(); dist.barrier(); t_full = time.perf_counter() - t0
    times, affected = [], []
    for batch in (batch1, batch2):
        m.dynamic_add_edges(batch)
        m._device_graph()                     # host CSR rebuild + upload are outside the timed region
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); blk = m.structural_distance_update_sharded(rank, world, peer=peer)
        torch.cuda.synchronize(); dist.barrier(); times.append(time.perf_counter() - t0)
        affected.append(int(m.last_affected.numel()))
    p = m._plan
    ok = torch.equal(blk, fresh[p.row0:p.row0 + p.n_rows])
    t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"peer={peer} world={world} n={n} hop={hop} inserted=2x{k_ins} affected={affected}: "
              f"blocks bit-equal to from-scratch: {bool(t.item())}; full step {t_full*1e3:.2f} ms, "
              f"first update {times[0]*1e3:.2f} ms, second update {times[1]*1e3:.2f} ms", flush=True)
dist.destroy_process_group()
