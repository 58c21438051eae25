#!/bin/bash
# refill / leaf / burst of the depth >= 1 launches again, now that the shade queues no longer depend on the retire order
mkdir -p gpurun_out
python tools/perf_ab.py --workload mesh1m --reps 3 base SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=12 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=20 \
  SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=8 SG_INTERIOR_BURST=3 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c2.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c2.log
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=18 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c5.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c5.log
python tools/perf_ab.py --workload glass --reps 1 base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=18 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c3.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c3.log
