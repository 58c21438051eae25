#!/bin/bash
# run on the GPU box under gpurun: bench line, launch list, and one full ncu capture of the traversal kernels
TAG=${1:-r01}
python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -s 0 -c 4 -o gpurun_out/${TAG}_trace python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench.json
