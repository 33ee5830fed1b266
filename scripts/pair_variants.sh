#!/bin/bash
# Times the pairwise kernel variants (HSD_PAIR_V2="kc,stages,packed,prodw,lookahead"; 0 = round-1 kernel)
# on the C2 / C3 graphs through the C-ABI; prints per-variant ms, % of the live FP32 issue peak, checksum.
# usage: scripts/pair_variants.sh N HOPS "variant variant ..."
N=${1:-20000}; H=${2:-3}; shift 2
for v in ${@:-0 16,4,0,0,2 16,4,1,0,2}; do
  echo "=== HSD_PAIR_V2=$v  N=$N hops=$H"
  HSD_PAIR_V2=$v python scripts/time_c2.py $N $H 2>&1 | grep -E "iter 3|checksum|Error|error"
done
