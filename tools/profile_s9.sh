#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o /tmp/s9_trace python bench.py --steps 1 --warmup 1 --spp 16 --no-cpu-baseline > gpurun_out/s9_ncu_trace.log 2>&1
ncu -i /tmp/s9_trace.ncu-rep --page raw --csv > gpurun_out/s9_trace_raw.csv 2>/dev/null
ncu -i /tmp/s9_trace.ncu-rep --page source --csv --print-source sass > gpurun_out/s9_trace_sass.csv 2>/dev/null
gzip -f gpurun_out/s9_trace_sass.csv
ls -la gpurun_out | tail -5
