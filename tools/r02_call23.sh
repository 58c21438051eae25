#!/bin/bash
# source-level capture of the lean shade kernels with 256-thread CTAs + stage barriers: depth-0 diffuse, depth-0 conductor, depth-1 diffuse
mkdir -p gpurun_out
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -f -o gpurun_out/r02_shade_src2 \
   python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_shade_src2.log 2>&1
ls -la gpurun_out/r02_shade_src2.ncu-rep
