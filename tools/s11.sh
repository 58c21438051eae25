#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 300 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline | cut -c1-260
timeout 300 python tools/bench_brief.py --workload instanced --steps 2 --warmup 3 --no-cpu-baseline | cut -c1-260
