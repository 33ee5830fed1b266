#!/bin/bash
# Round-2 evidence capture (1 GPU): launch list of the bench command + ncu --set full of the two hot kernels.
# usage: scripts/profile_bench.sh TAG
TAG=${1:-r2}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/plain_${TAG}.log 2> gpurun_out/plain_${TAG}.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
python scripts/time_c2.py 20000 3 > gpurun_out/plain_c2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'pairwise_l1|bfs_ring|ball_or|ring_cdf' -s 7 -c 7 -f -o gpurun_out/prof_${TAG} python scripts/time_c2.py 20000 3 > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
