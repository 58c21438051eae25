#!/bin/bash
# wavefronts in flight: 1 / 2 / 3 / 4
mkdir -p gpurun_out; rm -f gpurun_out/r02_overlap2.log
for W in "composite --spp 64 --reps 2" "instanced --reps 1" "glass --reps 1" "cornell --reps 3" "mesh1m --reps 2"; do
  for O in 1 2 3 4; do
    echo "== $W SG_OVERLAP=$O" >> gpurun_out/r02_overlap2.log
    SG_OVERLAP=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_overlap2.err | cut -c1-110 >> gpurun_out/r02_overlap2.log
  done
done
cat gpurun_out/r02_overlap2.log
