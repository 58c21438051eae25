#!/bin/bash
for t in 128 64 256; do
  lib=$PWD/shimmer_b200/libshimmer_gpu.so; [ $t != 128 ] && lib=$PWD/shimmer_b200/libshimmer_gpu_t$t.so
  SHIMMER_GPU_LIB=$lib timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline | sed "s/^/[threads=$t] /" | cut -c1-200
done
for pf in 2097152 8388608 16777216; do
  timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline --paths-in-flight $pf | sed "s/^/[pf=$pf] /" | cut -c1-200
done
