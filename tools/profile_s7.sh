#!/bin/bash
# GPU-box script: new gpu tests, results table, shade-kernel + instanced-scene ncu captures
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s7_pytest.log 2>&1; tail -3 gpurun_out/s7_pytest.log
timeout 600 python tools/results_table.py --out gpurun_out/results.json --md gpurun_out/results.md > gpurun_out/results.log 2>&1; cat gpurun_out/results.md
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 4 -o gpurun_out/s7_shade python bench.py --steps 1 --warmup 1 --spp 8 --no-cpu-baseline > gpurun_out/s7_ncu_shade.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/s7_inst_launches.csv python tools/perf_ab.py --workload instanced --spp 8 --reps 1 base > gpurun_out/s7_inst_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_shade" -c 8 -o gpurun_out/s7_inst python tools/perf_ab.py --workload instanced --spp 4 --reps 1 base > gpurun_out/s7_ncu_inst.log 2>&1
ncu --set full --clock-control none -k regex:k_trace -c 2 -o gpurun_out/s7_trace_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s7_ncu_trace_full.log 2>&1
ls -la gpurun_out | tail -12
