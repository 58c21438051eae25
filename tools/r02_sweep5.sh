#!/bin/bash
# depth-0 closest-hit knobs (camera rays) on C5 (64 spp), C2 and C4
mkdir -p gpurun_out
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD_D0=20 SG_REFILL_THRESHOLD_D0=24 SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 \
   SG_REFILL_THRESHOLD_D0=28,SG_INTERIOR_BURST_D0=8 SG_REFILL_THRESHOLD_D0=28,SG_INTERIOR_BURST_D0=2 2> gpurun_out/r02_sweep5_c5.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c5.log
python tools/perf_ab.py --workload mesh1m --reps 3 base SG_REFILL_THRESHOLD_D0=24 SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 2> gpurun_out/r02_sweep5_c2.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c2.log
python tools/perf_ab.py --workload instanced --reps 1 base SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 SG_INTERIOR_BURST_D0=2 2> gpurun_out/r02_sweep5_c4.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c4.log
