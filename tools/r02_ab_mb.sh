#!/bin/bash
# A/B: resident blocks per SM of the lean shade kernels (4 = default build, 5 / 6 / 8 = register caps 96 / 80 / 64)
mkdir -p gpurun_out
for V in base mb5 mb6 mb8; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> gpurun_out/r02_ab_mb.log
  python tools/perf_ab.py --workload composite --spp 64 --reps 2 base >> gpurun_out/r02_ab_mb.log 2>> gpurun_out/r02_ab_mb.err
  python tools/perf_ab.py --workload mesh1m --reps 3 base >> gpurun_out/r02_ab_mb.log 2>> gpurun_out/r02_ab_mb.err
done
cat gpurun_out/r02_ab_mb.log
