#!/bin/bash
# ncu --set full of the pairwise kernel for a list of HSD_PAIR_V2 variants (C2 graph); one report each.
# usage: scripts/ncu_pair.sh tag variant [variant ...]      -> gpurun_out/prof_<tag>_<i>.ncu-rep
TAG=$1; shift
i=0
for v in "$@"; do
  export HSD_PAIR_V2=$v
  python scripts/time_c2.py 20000 3 > gpurun_out/plain_${TAG}_$i.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:pairwise_l1 -s 1 -c 1 \
      -f -o gpurun_out/prof_${TAG}_$i python scripts/time_c2.py 20000 3 > gpurun_out/ncu_${TAG}_$i.log 2>&1
  echo "variant $v -> prof_${TAG}_$i rc=$?"
  i=$((i+1))
done
