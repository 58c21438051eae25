#!/bin/bash
for pf in 16777216 33554432 67108864; do
  timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline --paths-in-flight $pf | sed "s/^/[pf=$pf] /" | cut -c1-210
done
