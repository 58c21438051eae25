#!/bin/bash
# textured shade kernels: CTAs of 256 threads (x2 = 2 per SM / 128 regs, x3 = 3 per SM / 80 regs) with stage barriers, against 128 x 5
mkdir -p gpurun_out
L=gpurun_out/r02_c33_perf.log; : > $L
for V in base x2 x3; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 600 python tools/perf_ab.py --workload instanced --spp 32 --reps 2 base SG_SHADE_SYNC_TEX=2 SG_SHADE_SYNC_TEX=3 SG_SHADE_SYNC_TEX=19 2>> gpurun_out/r02_c33.err | cut -c1-200 >> $L
done
cat $L
