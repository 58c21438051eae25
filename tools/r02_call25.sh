#!/bin/bash
# ncu --set full of the streaming kernels of a C5 batch: k_generate, k_film, the queue compaction
mkdir -p gpurun_out
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_generate|k_film|k_queue' -c 5 -f -o gpurun_out/r02_stream_src \
   python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_stream_src.log 2>&1
ls -la gpurun_out/r02_stream_src.ncu-rep
