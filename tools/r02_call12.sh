#!/bin/bash
# instanced traversal with the render-space ray parked in shared memory: C4 timing + instancing / ray-cast parity tests
mkdir -p gpurun_out
python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_c12.err | tee gpurun_out/r02_c12.log
python -m pytest tests -m gpu -x -q -k "inst or raycast or sphere or patch or configs and not c4_converged" > gpurun_out/r02_c12_pytest.log 2>&1; tail -4 gpurun_out/r02_c12_pytest.log
