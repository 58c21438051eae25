#!/bin/bash
# ray-order experiment: what would sorting the ray queue buy the closest-hit kernel?
mkdir -p gpurun_out
timeout 600 python tools/exp_ray_order.py --workload mesh1m --spp 16 > gpurun_out/r02_ray_order_c2.log 2>&1
timeout 600 python tools/exp_ray_order.py --workload composite --spp 8 > gpurun_out/r02_ray_order_c5.log 2>&1
tail -30 gpurun_out/r02_ray_order_c2.log gpurun_out/r02_ray_order_c5.log
