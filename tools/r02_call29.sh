#!/bin/bash
# wavefront width on C5 again (two wavefronts in flight): 32 / 64 (default) / 128 Mi paths per wavefront; refill thresholds on the new build
mkdir -p gpurun_out
L=gpurun_out/r02_c29_perf.log; : > $L
timeout 900 python tools/perf_ab.py --workload composite --spp 128 --reps 2 base PIF=33554432 PIF=134217728 SG_REFILL_THRESHOLD=18 SG_REFILL_THRESHOLD=18,SG_INTERIOR_BURST=6 2>> gpurun_out/r02_c29.err | cut -c1-200 >> $L
cat $L; tail -3 gpurun_out/r02_c29.err
