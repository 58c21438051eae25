#!/bin/bash
# round 2, call 3 (1 GPU): whole GPU suite after the ABI v9 / ADVICE changes, then the profiler passes on C5 and C2
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -q --durations=8) > gpurun_out/r02_c3_pytest.log 2>&1
tail -15 gpurun_out/r02_c3_pytest.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_composite.csv \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_issue_composite.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_issue_mesh1m.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_trace|k_generate|k_film' -c 14 -o gpurun_out/r02_full_composite \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_full_composite.log 2>&1
tail -3 gpurun_out/r02_issue_composite.log gpurun_out/r02_issue_mesh1m.log gpurun_out/r02_launches_composite.log gpurun_out/r02_full_composite.log
ls -la gpurun_out | tail -12
