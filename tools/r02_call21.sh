#!/bin/bash
# stage barriers + CTA size of the lean shade kernels (instruction-fetch sharing between the warps of an SM)
mkdir -p gpurun_out
L=gpurun_out/r02_shade_sync.log; : > $L
for V in base t256 t512; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V composite" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_SYNC=16 SG_SHADE_SYNC=4 SG_SHADE_SYNC=12 SG_SHADE_SYNC=31 2>> gpurun_out/r02_shade_sync.err | cut -c1-200 >> $L
  echo "== $V mesh1m" >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_SHADE_SYNC=12 SG_SHADE_SYNC=31 2>> gpurun_out/r02_shade_sync.err | cut -c1-200 >> $L
done
cat $L
