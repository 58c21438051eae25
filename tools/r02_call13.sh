#!/bin/bash
# C4 after the material sort + parked ray: ncu of the shade / sort kernels and the traversal kernels
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_sort' -c 6 -o /tmp/r02_c4b_shade \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4b_shade.log 2>&1
ncu -i /tmp/r02_c4b_shade.ncu-rep --page raw --csv > gpurun_out/r02_c4b_shade_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r02_c4b_trace \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4b_trace.log 2>&1
ncu -i gpurun_out/r02_c4b_trace.ncu-rep --page raw --csv > gpurun_out/r02_c4b_trace_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_instanced_b.csv \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_launches_instanced_b.log 2>&1
ls -la gpurun_out
