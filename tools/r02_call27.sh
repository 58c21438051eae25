#!/bin/bash
# depth-0 closest-hit launch reads path = index (no queue indirection); pf = software prefetch of the next item's load chain in the lean shade kernels;
# racecheck / synccheck of the barrier-synchronised shade kernels
mkdir -p gpurun_out
L=gpurun_out/r02_c27_perf.log; : > $L
for V in base pf; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
done
unset SHIMMER_GPU_LIB
cat $L
bash tools/sanitize_race.sh
