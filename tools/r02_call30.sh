#!/bin/bash
# 2 GPUs: the multi-GPU tests (both forms behind the C ABI) and bench.py --gpus 2 the way the driver launches it
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_c30_multi.log 2>&1; tail -3 gpurun_out/r02_c30_multi.log
bash tools/r02_bench_n.sh 2 5 3 | tail -c 600
