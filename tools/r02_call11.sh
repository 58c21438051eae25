#!/bin/bash
# shade queues re-ordered by material id (textured scenes): off / on, then the textured parity tests
mkdir -p gpurun_out; rm -f gpurun_out/r02_sortmat.log
for O in 0 1; do
  echo "== instanced SG_SORT_MATERIALS=$O" >> gpurun_out/r02_sortmat.log
  SG_SORT_MATERIALS=$O python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_sortmat.err >> gpurun_out/r02_sortmat.log
done
cat gpurun_out/r02_sortmat.log
python -m pytest tests -m gpu -x -q -k "tex or variety or instanced or mix or configs and not c4_converged" > gpurun_out/r02_c11_pytest.log 2>&1; tail -4 gpurun_out/r02_c11_pytest.log
