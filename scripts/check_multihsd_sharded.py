"""torchrun check: MultiHSD.embed_device_sharded on `world` GPUs equals the single-GPU embedding."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch, torch.distributed as dist, networkx as nx
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from model import MultiHSD
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
G = nx.barabasi_albert_graph(n, 5, seed=0)
m = MultiHSD(G, "ba", 3, 4)
m.scales = np.exp(np.linspace(np.log(0.01), np.log(40.0 / m.lmax), 4)); m.CHEB_ORDER = 30
full = m.embed_device()
torch.cuda.synchronize(); dist.barrier()
import time
t0 = time.perf_counter(); sh = m.embed_device_sharded(rank, world); torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter() - t0
ok = torch.allclose(sh, full, rtol=1e-12, atol=1e-15)
flag = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"MultiHSD sharded world={world} n={n}: equals single GPU: {bool(flag.item())}, {t*1e3:.1f} ms")
dist.destroy_process_group()
