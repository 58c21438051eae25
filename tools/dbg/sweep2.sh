#!/bin/bash
for pf in 0 1; do
  SG_PREFETCH=$pf timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline | sed "s/^/[prefetch=$pf] /" | cut -c1-220
done
