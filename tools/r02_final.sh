#!/bin/bash
# final round-2 measurements on one GPU: results table, bench N = 1 with the reference arm, launch lists and issue captures of the final build
mkdir -p gpurun_out
python tools/results_table.py --out gpurun_out/r02_results.json --md gpurun_out/r02_results.md > gpurun_out/r02_results.log 2>&1; cat gpurun_out/r02_results.md
bash tools/r02_bench_n.sh 1 5 3 > /dev/null 2>&1
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for W in composite:8 mesh1m:0 instanced:8; do
  N=${W%%:*}; S=${W##*:}; A=""; if [ "$S" != "0" ]; then A="--spp $S"; fi
  SG_OVERLAP=1 ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_$N.csv \
      python tools/render_once.py --workload $N $A --warm 0 > gpurun_out/r02_issue_$N.log 2>&1
done
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_instanced.csv \
    python tools/render_once.py --workload instanced --spp 16 --warm 0 > gpurun_out/r02_launches_instanced.log 2>&1
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_launches_mesh1m.log 2>&1
cat gpurun_out/r02_issue_*.log
