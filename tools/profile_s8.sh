#!/bin/bash
# GPU-box script: ncu captures, exported to CSV on the box (the .ncu-rep files are too large to bring back together)
exp() { # rep tag
  ncu -i $1 --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
  ncu -i $1 --page source --csv --print-source sass > gpurun_out/$2_sass.csv 2>/dev/null
  gzip -f gpurun_out/$2_sass.csv
}
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 4 -o /tmp/s8_shade python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/s8_ncu_shade.log 2>&1
exp /tmp/s8_shade.ncu-rep s8_shade
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s8_inst_launches.csv python tools/perf_ab.py --workload instanced --spp 8 --reps 1 base > gpurun_out/s8_inst_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -c 8 -o /tmp/s8_inst python tools/perf_ab.py --workload instanced --spp 4 --reps 1 base > gpurun_out/s8_ncu_inst.log 2>&1
exp /tmp/s8_inst.ncu-rep s8_inst
ncu --set full --clock-control none -k regex:k_trace -c 2 -o /tmp/s8_trace_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s8_ncu_trace_full.log 2>&1
ncu -i /tmp/s8_trace_full.ncu-rep --page raw --csv > gpurun_out/s8_trace_full_raw.csv
ls -la gpurun_out /tmp/*.ncu-rep | tail -20
