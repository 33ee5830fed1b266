import os, sys
sys.path.insert(0, ".")
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm_mem.empty((1024,), dtype=torch.float32, device="cuda")
h = symm_mem.rendezvous(t, dist.group.WORLD)
x = torch.zeros(1, device="cuda")
for name, fn in [("symm.barrier", lambda: h.barrier()), ("nccl all_reduce(1 float)", lambda: dist.all_reduce(x)),
                 ("fill 256MB", None)]:
    if fn is None:
        buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
        fn = lambda: buf.fill_(1.0)
    for _ in range(10): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): fn()
    b.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name}: {a.elapsed_time(b) * 10:.1f} us per call (world {world})")
dist.destroy_process_group()
