#!/bin/bash
# round 2, call 6: racecheck / synccheck, the five-config results table, instruction counts of the current build
mkdir -p gpurun_out
bash tools/sanitize_race.sh > gpurun_out/r02_sanitize_summary.txt 2>&1; cat gpurun_out/r02_sanitize_summary.txt
python tools/results_table.py --out gpurun_out/r02_results.json --md gpurun_out/r02_results.md > gpurun_out/r02_results.log 2>&1; cat gpurun_out/r02_results.md
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_composite.csv \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_issue_composite.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_issue_mesh1m.log 2>&1
cat gpurun_out/r02_issue_composite.log gpurun_out/r02_issue_mesh1m.log
