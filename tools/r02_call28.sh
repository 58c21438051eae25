#!/bin/bash
# scheduling-invariance test + full GPU suite; shade grid size sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c28_pytest.log 2>&1; tail -3 gpurun_out/r02_c28_pytest.log
L=gpurun_out/r02_c28_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_GRID=4 SG_SHADE_GRID=6 SG_SHADE_GRID=12 SG_SHADE_GRID=16 SG_SHADE_GRID=32 2>> gpurun_out/r02_c28.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_SHADE_GRID=4 SG_SHADE_GRID=16 SG_SHADE_GRID=32 2>> gpurun_out/r02_c28.err | cut -c1-200 >> $L
cat $L
