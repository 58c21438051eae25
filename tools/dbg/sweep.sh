#!/bin/bash
# sweep scheduling knobs of the traversal kernel on the C2 workload
for cfg in "8 6 2" "4 4 2" "12 8 2" "16 8 2" "8 6 4" "8 6 1" "12 12 4" "16 16 4" "6 2 2" "24 16 4" "12 4 3"; do
  set -- $cfg
  SG_LEAF_THRESHOLD=$1 SG_REFILL_THRESHOLD=$2 SG_INTERIOR_BURST=$3 timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline | sed "s/^/[L=$1 R=$2 B=$3] /" | cut -c1-200
done
