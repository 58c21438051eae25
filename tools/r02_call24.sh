#!/bin/bash
# occupancy of the lean shade kernels again, now that they are no longer instruction-fetch bound:
# base = 2 x 256 threads (16 warps, <= 128 regs), t320 = 2 x 320 (20 warps, 96 regs), b3 = 3 x 256 (24 warps, 80 regs), b4 = 4 x 256 (32 warps, 64 regs)
mkdir -p gpurun_out
L=gpurun_out/r02_shade_occ.log; : > $L
for V in base t320 b3 b4; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
done
cat $L
