#!/bin/bash
mkdir -p gpurun_out
SG_OVERLAP=1 ncu --set full --clock-control none -k regex:'k_shade|k_trace' -c 9 -o /tmp/r02_final_c5 python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_final_c5.log 2>&1
ncu -i /tmp/r02_final_c5.ncu-rep --page raw --csv > gpurun_out/r02_final_c5_raw.csv 2>/dev/null
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
