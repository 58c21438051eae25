#!/bin/bash
# C4: full ncu capture of the textured shade kernel and the instanced traversal kernel (first launches of one batch)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -o gpurun_out/r02_c4_shade \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r02_c4_trace \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4_trace.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_instanced.csv \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_launches_instanced.log 2>&1
ls -la gpurun_out
