#!/bin/bash
# bench.py at N GPUs the way the driver launches it (torchrun for N > 1) + the reference arm at N = 1
N=${1:-1}; STEPS=${2:-5}; WARM=${3:-3}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps $STEPS --warmup $WARM > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
  python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps $STEPS --warmup $WARM \
     > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
fi
tail -c 2500 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
