#!/bin/bash
# round 2, call 1: wavefront-width sweep + scheduling-knob sweep on the round-1 build
mkdir -p gpurun_out
python tools/perf_ab.py --workload mesh1m base PIF=1048576 PIF=4194304 PIF=16777216 PIF=67108864 \
  SG_REFILL_THRESHOLD=4 SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=12 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=24 \
  SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=12 SG_LEAF_THRESHOLD=16 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=8 \
  > gpurun_out/r02_sweep_c2.log 2> gpurun_out/r02_sweep_c2.err
python tools/perf_ab.py --workload instanced --reps 1 base PIF=1048576 PIF=4194304 PIF=16777216 PIF=67108864 \
  > gpurun_out/r02_sweep_c4.log 2> gpurun_out/r02_sweep_c4.err
python tools/perf_ab.py --workload composite --spp 128 --reps 1 base PIF=4194304 PIF=16777216 \
  > gpurun_out/r02_sweep_c5.log 2> gpurun_out/r02_sweep_c5.err
cat gpurun_out/r02_sweep_c2.log gpurun_out/r02_sweep_c4.log gpurun_out/r02_sweep_c5.log
