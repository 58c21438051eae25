#!/bin/bash
mkdir -p gpurun_out
python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_c15.err | cut -c1-170 | tee gpurun_out/r02_c15.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c15_pytest.log 2>&1; tail -4 gpurun_out/r02_c15_pytest.log
