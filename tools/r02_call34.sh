#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c34_pytest.log 2>&1; tail -3 gpurun_out/r02_c34_pytest.log
L=gpurun_out/r02_c34_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base 2>> gpurun_out/r02_c34.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c34.err | cut -c1-200 >> $L
cat $L
