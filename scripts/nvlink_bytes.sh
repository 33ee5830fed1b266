#!/bin/bash
# NVLink byte counters around a sharded run (VERDICT r1 #4c): `nvidia-smi nvlink -gt d` before and after
# K steps of the C3 (or given) workload on N GPUs; scripts/nvlink_report.py turns the deltas into
# bytes per step per GPU next to the algorithmic bytes of the two fused compute+exchange kernels.
# usage: scripts/nvlink_bytes.sh N_GPUS WORKLOAD STEPS TAG
N=${1:-2}; W=${2:-c3}; K=${3:-5}; TAG=${4:-nvl}
nvidia-smi nvlink -gt d > gpurun_out/${TAG}_before.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --workload $W --steps $K --warmup 3 --no-extras --no-e2e --no-cpu-baseline \
  > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/${TAG}_after.txt 2>&1
python scripts/nvlink_report.py gpurun_out/${TAG}_before.txt gpurun_out/${TAG}_after.txt gpurun_out/${TAG}_bench.json $((K+3)) | tee gpurun_out/${TAG}_report.txt
