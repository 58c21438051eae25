#!/bin/bash
for b in 4 5 6 8; do
  lib=$PWD/shimmer_b200/libshimmer_gpu.so; [ $b != 4 ] && lib=$PWD/shimmer_b200/libshimmer_gpu_b$b.so
  SHIMMER_GPU_LIB=$lib timeout 300 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline | sed "s/^/[shade blocks=$b] /" | cut -c1-200
done
