#!/bin/bash
# round 2, call 4 (1 GPU): profiler passes on C5 and C2 (reports are exported to CSV on the box: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_composite.csv \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_issue_composite.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_issue_mesh1m.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
for K in k_trace k_shade; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 4 -o /tmp/r02_full_${K} \
      python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_full_${K}.log 2>&1
  ncu -i /tmp/r02_full_${K}.ncu-rep --page raw --csv > gpurun_out/r02_full_composite_${K}_raw.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:'k_generate|k_film' -c 2 -o /tmp/r02_full_gf python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_full_gf.log 2>&1
ncu -i /tmp/r02_full_gf.ncu-rep --page raw --csv > gpurun_out/r02_full_composite_genfilm_raw.csv 2>/dev/null
# keep the shade report itself when it fits (per-instruction stall analysis happens off the box)
S=$(stat -c %s /tmp/r02_full_k_shade.ncu-rep); if [ "$S" -lt 40000000 ]; then cp /tmp/r02_full_k_shade.ncu-rep gpurun_out/r02_full_composite_k_shade.ncu-rep; fi
cat gpurun_out/r02_issue_composite.log gpurun_out/r02_issue_mesh1m.log
ls -la gpurun_out /tmp/*.ncu-rep
