#!/bin/bash
# k_film: segmented butterfly reduction over aligned runs of one pixel
mkdir -p gpurun_out
L=gpurun_out/r02_c31_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload instanced --spp 16 --reps 2 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
cat $L
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_film|k_generate' --csv --log-file gpurun_out/r02_c31_film.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > /dev/null 2>&1
grep -E "k_film|k_generate" gpurun_out/r02_c31_film.csv | cut -d, -f5,15-
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scheduling.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
