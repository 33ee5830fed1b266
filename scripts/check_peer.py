"""torchrun check: every rank's peer-memory result block equals the rows a single GPU computes."""
import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
from hsd_b200 import engine
from hsd_b200.graph import powerlaw_graph
from hsd_b200.sharded import ShardedDegreeHSD
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = powerlaw_graph(n, 5, seed=0)
dg = engine.DeviceGraph.upload(g)
single = ShardedDegreeHSD(dg, 3, 0, 1).step()
for peer in (True, False):
    plan = ShardedDegreeHSD(dg, 3, rank, world, peer=peer)
    for _ in range(3):
        blk = plan.step()
    torch.cuda.synchronize(); dist.barrier()
    ok = torch.equal(blk, single[plan.row0:plan.row0 + plan.n_rows])
    t = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"peer={peer} world={world} n={n}: blocks bit-equal to single GPU: {bool(t.item())}")
dist.destroy_process_group()
