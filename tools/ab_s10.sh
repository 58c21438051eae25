#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in post nopost; do
  lib=$PWD/shimmer_b200/libshimmer_gpu.so; [ $v = nopost ] && lib=$PWD/shimmer_b200/libshimmer_gpu_nopost.so
  SHIMMER_GPU_LIB=$lib timeout 300 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline | sed "s/^/[$v] /" | cut -c1-260
done
for lt in 4 6 12 16; do
  SG_LEAF_THRESHOLD=$lt timeout 300 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline | sed "s/^/[post leaf=$lt] /" | cut -c1-260
done
for ib in 2 8; do
  SG_INTERIOR_BURST=$ib timeout 300 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline | sed "s/^/[post burst=$ib] /" | cut -c1-260
done
