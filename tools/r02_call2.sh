#!/bin/bash
# round 2, call 2 (2 GPUs): regression run of the whole GPU suite after the multi-GPU refactor, the new tests, and bench at N = 1 and 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_c2_gpus.txt
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r02_c2_pytest.log 2>&1
tail -25 gpurun_out/r02_c2_pytest.log
python bench.py --steps 2 --warmup 3 --e2e-steps 2 > gpurun_out/r02_c2_bench_n1.json 2> gpurun_out/r02_c2_bench_n1.err
tail -c 1500 gpurun_out/r02_c2_bench_n1.json; tail -5 gpurun_out/r02_c2_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 \
   > gpurun_out/r02_c2_bench_n2.json 2> gpurun_out/r02_c2_bench_n2.err
tail -c 1500 gpurun_out/r02_c2_bench_n2.json; tail -5 gpurun_out/r02_c2_bench_n2.err
