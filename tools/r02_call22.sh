#!/bin/bash
# 256-thread lean shade CTAs + barriers around sample_ld as defaults: GPU tests, all five configs, textured-kernel barrier sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c22_pytest.log 2>&1; tail -3 gpurun_out/r02_c22_pytest.log
L=gpurun_out/r02_shade_sync2.log; : > $L
echo "== instanced" >> $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base SG_SHADE_SYNC_TEX=2 SG_SHADE_SYNC_TEX=4 SG_SHADE_SYNC_TEX=6 SG_SHADE_SYNC_TEX=14 SG_SHADE_SYNC=0 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== glass" >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 SG_SHADE_SYNC=8 SG_SHADE_SYNC=14 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== cornell" >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== composite" >> $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 SG_SHADE_SYNC=6 SG_SHADE_SYNC=14 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
cat $L
