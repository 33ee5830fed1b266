"""e2e path timing (dev aid): HostDegreePipeline with and without the host-side mirror, and the
mirror alone on pinned memory for several thread counts."""
import sys, time, os
sys.path.insert(0, ".")
import numpy as np, torch
from hsd_b200 import engine
from hsd_b200._lib import lib, check
from hsd_b200.graph import powerlaw_graph

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 3
print("cpus", len(os.sched_getaffinity(0)), os.cpu_count())
host = torch.empty((n, n), dtype=torch.float32).pin_memory()
host.uniform_()
for T in (1, 4, 8, 16, 32, 64):
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        check(lib.hsd_mirror_upper_to_lower_host(host.data_ptr(), n, n, 0, n, T))
        ts.append(time.perf_counter() - t)
    print(f"mirror alone threads={T}: {min(ts)*1e3:.1f} ms ({n*n*4/min(ts)/1e9:.0f} GB/s traffic)")
g = powerlaw_graph(n, 5, seed=0)
ref = None
for mirror in (False, True):
    for chunks in ((8,) if not mirror else (8, 12, 16, 24)):
        for T in ((0,) if not mirror else (8, 16, 32)):
            pipe = engine.HostDegreePipeline(g, hops, n_chunks=chunks, host_mirror=mirror, host_threads=T or None)
            for _ in range(2):
                pipe.run(host)
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(5):
                pipe.run(host)
            dt = (time.perf_counter() - t) / 5
            cs = float(host.double().sum())
            if ref is None:
                ref = host.clone()
            print(f"mirror={mirror} chunks={chunks} threads={T}: {dt*1e3:.2f} ms/step, d2h {pipe.d2h_bytes/1e9:.2f} GB, "
                  f"equal_to_full_copy={bool(torch.equal(host, ref))}")
            del pipe
