"""NVLink Tx/Rx bytes per step per GPU from two `nvidia-smi nvlink -gt d` dumps, against the
algorithmic bytes of the fused all-gather (BFS kernel) and the peer-mirrored pairwise kernel."""
import json, re, sys

def parse(path):
    gpus, cur = {}, None
    for line in open(path):
        m = re.match(r"GPU (\d+):", line)
        if m:
            cur = int(m.group(1)); gpus[cur] = {"tx": 0, "rx": 0}
            continue
        m = re.search(r"Link \d+: Data (Tx|Rx): (\d+) KiB", line)
        if m and cur is not None:
            gpus[cur][m.group(1).lower()] += int(m.group(2)) * 1024
    return gpus

before, after, bench, steps = parse(sys.argv[1]), parse(sys.argv[2]), sys.argv[3], int(sys.argv[4])
line = None
for l in open(bench):
    if l.startswith("{"):
        line = json.loads(l)
cfg = line["config"]
n, world = cfg["n_nodes"], line["n_gpus"]
k = cfg["signature_len"]
ld = (k + 3) // 4 * 4
per = ((n + world - 1) // world + 3) // 4 * 4
table = n * ld * 4                                  # every rank stores its n/world rows into world-1 peer tables
alg_sig = table * (world - 1) / world
alg_pair = 4.0 * n * n / world * (world - 1) / world   # a rank produces N^2/world entries (both triangles); (world-1)/world of them land in a peer's block
print(f"workload {cfg['workload']} on {world} GPUs, {steps} steps incl. warm-up, {line['ms_per_step']:.3f} ms/step")
print(f"algorithmic NVLink bytes per step per GPU: signatures (fused all-gather) {alg_sig/1e6:.1f} MB + result tiles (direct + mirrored peer stores) {alg_pair/1e6:.1f} MB = {(alg_sig+alg_pair)/1e6:.1f} MB each way")
for g in sorted(after):
    tx = (after[g]["tx"] - before.get(g, {"tx": 0})["tx"]) / steps
    rx = (after[g]["rx"] - before.get(g, {"rx": 0})["rx"]) / steps
    tot = alg_sig + alg_pair
    print(f"GPU {g}: Tx {tx/1e6:9.1f} MB/step ({tx/tot:5.2f}x algorithmic, {tx/(line['ms_per_step']*1e-3)/1e9:6.1f} GB/s)   "
          f"Rx {rx/1e6:9.1f} MB/step ({rx/tot:5.2f}x, {rx/(line['ms_per_step']*1e-3)/1e9:6.1f} GB/s)")
