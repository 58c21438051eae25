#!/bin/bash
# co-residency of the issue-bound traversal kernels and the latency-bound shade kernels of the two wavefronts: cap the persistent traversal grid
# (9 CTAs per SM fill the register file) and size the shade grid to what fits beside it
mkdir -p gpurun_out
L=gpurun_out/r02_c35_perf.log; : > $L
timeout 900 python tools/perf_ab.py --workload composite --spp 128 --reps 2 base SG_TRACE_BLOCKS_PER_SM=7 SG_TRACE_BLOCKS_PER_SM=5 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=2 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=4 SG_TRACE_BLOCKS_PER_SM=6,SG_SHADE_GRID=2 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=2,SG_OVERLAP=3 2>> gpurun_out/r02_c35.err | cut -c1-200 >> $L
cat $L
