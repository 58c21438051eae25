#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in new base; do
  lib=$PWD/shimmer_b200/libshimmer_gpu.so; [ $v = base ] && lib=$PWD/shimmer_b200/libshimmer_gpu_base.so
  SHIMMER_GPU_LIB=$lib timeout 300 python tools/bench_brief.py --workload instanced --steps 2 --warmup 3 --no-cpu-baseline | sed "s/^/[$v C4] /" | cut -c1-200
  SHIMMER_GPU_LIB=$lib timeout 300 python tools/bench_brief.py --steps 4 --warmup 3 --no-cpu-baseline | sed "s/^/[$v C2] /" | cut -c1-200
done
