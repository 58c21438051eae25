#!/bin/bash
# Timing experiment behind the "split k_shade?" decision: per-kernel device times (ncu launch list) of the lean diffuse
# shade kernel at depth 0 of one C5 batch, for builds with NEE / BSDF sampling / both compiled out.  Depth-0 inputs are identical in
# all variants.  If T(full) is well above T(neither) + [T(nosample) - T(neither)] + [T(nonee) - T(neither)] the stages interfere
# (instruction cache, registers) and separate kernels would win; if it is additive, a split only adds state traffic.
mkdir -p gpurun_out
for V in base nonee nosample neither; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio \
      --clock-control none -k regex:k_shade -c 4 --csv --log-file gpurun_out/r02_exp_split_$V.csv \
      python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_exp_split_$V.log 2>&1
done
python - <<'PY'
import csv, collections
for v in ["base", "nonee", "nosample", "neither"]:
    rows = [r for r in csv.reader(open("gpurun_out/r02_exp_split_%s.csv" % v)) if len(r) > 10]
    h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[ii], {"k": r[ki][:40]})[r[mi]] = r[vi]
    for i, d in per.items():
        print(v, i, d["k"], {k: x for k, x in d.items() if k != "k"})
PY
