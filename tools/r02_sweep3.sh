#!/bin/bash
# scheduling knobs of the INSTANCED traversal kernels on C4 (the round-2 defaults were tuned on the triangle-only kernels)
mkdir -p gpurun_out
python tools/perf_ab.py --workload instanced --reps 1 base SG_LEAF_THRESHOLD=2 SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=8 SG_LEAF_THRESHOLD=12 SG_LEAF_THRESHOLD=16 \
  SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=20 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=8 SG_LEAF_THRESHOLD=10,SG_INTERIOR_BURST=8 \
  2> gpurun_out/r02_sweep3_c4.err | cut -c1-170 | tee gpurun_out/r02_sweep3_c4.log
