#!/bin/bash
# usage: scripts/nvlink_ncu.sh N HOPS WORLD TAG   (under gpurun --gpus WORLD)
N=${1:-20000}; H=${2:-3}; W=${3:-2}; TAG=${4:-r2_nvl}
python scripts/nvlink_ncu_probe.py $N $H $W > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --devices 0 --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes.sum,nvlrx__bytes_data_user.sum,dram__bytes_write.sum,dram__bytes_read.sum \
    --clock-control none -k regex:'bfs_ring|pairwise_l1' --csv --log-file gpurun_out/${TAG}_ncu.csv \
    python scripts/nvlink_ncu_probe.py $N $H $W > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/${TAG}_plain.log | tail -4
