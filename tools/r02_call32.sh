#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/kernel_trace.py --workload composite --spp 16 > gpurun_out/r02_ktrace_c5.log 2>&1; tail -20 gpurun_out/r02_ktrace_c5.log
timeout 600 python tools/kernel_trace.py --workload mesh1m > gpurun_out/r02_ktrace_c2.log 2>&1; tail -14 gpurun_out/r02_ktrace_c2.log
