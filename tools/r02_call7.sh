#!/bin/bash
# two wavefronts in flight (SG_OVERLAP=2, the new default) against one (SG_OVERLAP=1), then the GPU suite on the default
mkdir -p gpurun_out; rm -f gpurun_out/r02_overlap.log
for W in "composite --spp 64 --reps 2" "mesh1m --reps 3" "glass --reps 1" "instanced --reps 1" "cornell --reps 3"; do
  for O in 1 2; do
    echo "== $W SG_OVERLAP=$O" >> gpurun_out/r02_overlap.log
    SG_OVERLAP=$O python tools/perf_ab.py --workload $W base >> gpurun_out/r02_overlap.log 2>> gpurun_out/r02_overlap.err
  done
done
cat gpurun_out/r02_overlap.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c7_pytest.log 2>&1; tail -4 gpurun_out/r02_c7_pytest.log
