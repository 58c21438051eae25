#!/bin/bash
mkdir -p gpurun_out
python tools/perf_ab.py --workload mesh1m --reps 3 base 2>> gpurun_out/r02_c17.err | cut -c1-170 | tee gpurun_out/r02_c17.log
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_mesh1m_b.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_launches_mesh1m_b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_mesh1m_b.csv')) if len(r)>10]
h=rows[0]; ki,vi=h.index("Kernel Name"),h.index("Metric Value")
for r in rows[1:45]:
    print("%-40s %9.3f ms" % (r[ki].split("(")[0].replace("void ","")[:40], float(r[vi].replace(",",""))/1e6))
PY
