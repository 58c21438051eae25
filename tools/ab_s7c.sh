#!/bin/bash
# same-box A/B: multi-TU build (lean kernels, queue mask) vs the same sources as one TU vs the pre-ABI-v7 tree
run() { (cd $2 && SHIMMER_GPU_LIB=$3 timeout 400 python tools/bench_brief.py $4 --no-cpu-baseline | head -1 | sed "s/^/[$1] /" | cut -c1-200); }
R=$PWD
for rep in 1 2; do
  run "multi  C2" $R $R/shimmer_b200/libshimmer_gpu.so "--steps 4 --warmup 3"
  run "single C2" $R $R/shimmer_b200/libshimmer_gpu_single.so "--steps 4 --warmup 3"
  run "old    C2" $R/_ab_old $R/_ab_old/shimmer_b200/libshimmer_gpu.so "--steps 4 --warmup 3"
done
run "multi  C4" $R $R/shimmer_b200/libshimmer_gpu.so "--workload instanced --steps 2 --warmup 2"
run "old    C4" $R/_ab_old $R/_ab_old/shimmer_b200/libshimmer_gpu.so "--workload instanced --steps 2 --warmup 2"
run "multi  C1" $R $R/shimmer_b200/libshimmer_gpu.so "--workload cornell --steps 10 --warmup 3"
run "old    C1" $R/_ab_old $R/_ab_old/shimmer_b200/libshimmer_gpu.so "--workload cornell --steps 10 --warmup 3"
run "multi  C3" $R $R/shimmer_b200/libshimmer_gpu.so "--workload glass --steps 2 --warmup 2"
run "old    C3" $R/_ab_old $R/_ab_old/shimmer_b200/libshimmer_gpu.so "--workload glass --steps 2 --warmup 2"
