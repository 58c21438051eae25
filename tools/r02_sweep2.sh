#!/bin/bash
# round 2: scheduling knobs around the new default (refill 12) on C2 and on C5's per-GPU share at N = 8 (128 spp)
mkdir -p gpurun_out
python tools/perf_ab.py --workload mesh1m base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=14 SG_REFILL_THRESHOLD=16 \
  SG_REFILL_THRESHOLD=14,SG_LEAF_THRESHOLD=6 SG_REFILL_THRESHOLD=14,SG_LEAF_THRESHOLD=10 SG_REFILL_THRESHOLD=12,SG_LEAF_THRESHOLD=6 \
  SG_REFILL_THRESHOLD=12,SG_INTERIOR_BURST=3 SG_REFILL_THRESHOLD=12,SG_INTERIOR_BURST=6 SG_REFILL_THRESHOLD=16,SG_INTERIOR_BURST=6 \
  SG_SMEM_LEVELS=16 SG_SMEM_LEVELS=24 SG_PREFETCH=1 > gpurun_out/r02_sweep2_c2.log 2> gpurun_out/r02_sweep2_c2.err
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=20 \
  SG_REFILL_THRESHOLD=16,SG_LEAF_THRESHOLD=6 SG_REFILL_THRESHOLD=16,SG_LEAF_THRESHOLD=12 SG_INTERIOR_BURST=8 > gpurun_out/r02_sweep2_c5.log 2> gpurun_out/r02_sweep2_c5.err
cat gpurun_out/r02_sweep2_c2.log gpurun_out/r02_sweep2_c5.log
