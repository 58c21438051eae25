#!/bin/bash
# order-preserving shade queues: off / on for every configuration, then the GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r02_ordered.log
for W in "mesh1m --reps 3" "composite --spp 64 --reps 2" "glass --reps 1" "instanced --reps 1" "cornell --reps 3"; do
  for O in 0 1; do
    echo "== $W SG_ORDERED_QUEUES=$O" >> gpurun_out/r02_ordered.log
    SG_ORDERED_QUEUES=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_ordered.err | cut -c1-170 >> gpurun_out/r02_ordered.log
  done
done
cat gpurun_out/r02_ordered.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c16_pytest.log 2>&1; tail -4 gpurun_out/r02_c16_pytest.log
