#!/bin/bash
# multi-GPU check on an 8-GPU box: C2 weak scaling at N=8 and N=4, C5 (4K, 1024 spp split 8 x 128) at N=8
run() { # N args...
  N=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2>gpurun_out/scale_err_$N.log | grep '^{' | tail -1
}
nvidia-smi -L | head -8
echo "C2 N=8"; run 8 --steps 3 --warmup 3 --no-cpu-baseline | tee gpurun_out/scale_c2_n8.json | cut -c1-400
echo "C2 N=4"; run 4 --steps 3 --warmup 3 --no-cpu-baseline | tee gpurun_out/scale_c2_n4.json | cut -c1-400
echo "C5 N=8"; run 8 --workload composite --spp 128 --steps 2 --warmup 3 --no-cpu-baseline | tee gpurun_out/scale_c5_n8.json | cut -c1-400
tail -3 gpurun_out/scale_err_8.log
