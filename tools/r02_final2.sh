#!/bin/bash
# last checks of round 2 on one GPU: the GPU test suite, smoke(), the default bench line, and the ncu launch list of the bench command itself
# (at 16 of the 1024 spp: two 8-spp wavefronts, the same launch sequence the full job repeats 64 times)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final2_pytest.log 2>&1; tail -2 gpurun_out/r02_final2_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -c 400 gpurun_out/r02_bench_default.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --spp 16 --no-c2 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
grep -c k_ gpurun_out/r02_launches_bench.csv
