#!/bin/bash
# k_generate without 64-bit divisions + one-pass 8-way k_queue_scan: GPU tests, five configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c26_pytest.log 2>&1; tail -3 gpurun_out/r02_c26_pytest.log
L=gpurun_out/r02_c26_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
cat $L
