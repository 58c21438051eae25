#!/bin/bash
# shadow traversal of depth d on a side stream, overlapping the closest-hit traversal of depth d + 1: off / on
mkdir -p gpurun_out; rm -f gpurun_out/r02_side.log
for W in "mesh1m --reps 3" "composite --spp 64 --reps 2" "instanced --reps 1" "glass --reps 1" "cornell --reps 3"; do
  for O in 0 1; do
    echo "== $W SG_SHADOW_SIDE_STREAM=$O" >> gpurun_out/r02_side.log
    SG_SHADOW_SIDE_STREAM=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_side.err | cut -c1-110 >> gpurun_out/r02_side.log
  done
done
cat gpurun_out/r02_side.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c9_pytest.log 2>&1; tail -4 gpurun_out/r02_c9_pytest.log
