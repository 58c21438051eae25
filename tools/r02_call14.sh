#!/bin/bash
# staged shading of textured scenes: off / on (C4), then the GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r02_staged.log
for O in 0 1; do
  echo "== instanced SG_STAGED_SHADING=$O" >> gpurun_out/r02_staged.log
  SG_STAGED_SHADING=$O python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_staged.err >> gpurun_out/r02_staged.log
done
cat gpurun_out/r02_staged.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c14_pytest.log 2>&1; tail -4 gpurun_out/r02_c14_pytest.log
