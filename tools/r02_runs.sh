#!/bin/bash
# Every GPU-box run of round 2, one function per gpurun call (usage: tools/r02_runs.sh <name>; `list` prints the names).
# They are the commands behind profiles/r02_experiments.md and the profiles/r02_* summaries; outputs land in gpurun_out/.
# Runs that loaded A/B builds expect shimmer_b200/ab/libshimmer_gpu_<variant>.so from
#   python -m shimmer_b200.build --variant <name> -D<MACRO>=<value> ...
mkdir -p gpurun_out

# round 2, call 2 (2 GPUs): regression run of the whole GPU suite after the multi-GPU refactor, the new tests, and bench at N = 1 and 2
call2() {
nvidia-smi -L > gpurun_out/r02_c2_gpus.txt
(time python -m pytest tests -m gpu -x -q --durations=15) > gpurun_out/r02_c2_pytest.log 2>&1
tail -25 gpurun_out/r02_c2_pytest.log
python bench.py --steps 2 --warmup 3 --e2e-steps 2 > gpurun_out/r02_c2_bench_n1.json 2> gpurun_out/r02_c2_bench_n1.err
tail -c 1500 gpurun_out/r02_c2_bench_n1.json; tail -5 gpurun_out/r02_c2_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 \
   > gpurun_out/r02_c2_bench_n2.json 2> gpurun_out/r02_c2_bench_n2.err
tail -c 1500 gpurun_out/r02_c2_bench_n2.json; tail -5 gpurun_out/r02_c2_bench_n2.err
}

# round 2, call 3 (1 GPU): whole GPU suite after the ABI v9 / ADVICE changes, then the profiler passes on C5 and C2
call3() {
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
}

# round 2, call 4 (1 GPU): profiler passes on C5 and C2 (reports are exported to CSV on the box: gpurun_out is capped at 64 MiB)
# keep the shade report itself when it fits (per-instruction stall analysis happens off the box)
call4() {
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_composite.csv \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_issue_composite.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_issue_mesh1m.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
for K in k_trace k_shade; do
  ncu --set full --clock-control none --import-source on -k regex:$K -c 4 -o /tmp/r02_full_${K} \
      python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_full_${K}.log 2>&1
  ncu -i /tmp/r02_full_${K}.ncu-rep --page raw --csv > gpurun_out/r02_full_composite_${K}_raw.csv 2>/dev/null
done
ncu --set full --clock-control none -k regex:'k_generate|k_film' -c 2 -o /tmp/r02_full_gf python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_full_gf.log 2>&1
ncu -i /tmp/r02_full_gf.ncu-rep --page raw --csv > gpurun_out/r02_full_composite_genfilm_raw.csv 2>/dev/null
S=$(stat -c %s /tmp/r02_full_k_shade.ncu-rep); if [ "$S" -lt 40000000 ]; then cp /tmp/r02_full_k_shade.ncu-rep gpurun_out/r02_full_composite_k_shade.ncu-rep; fi
cat gpurun_out/r02_issue_composite.log gpurun_out/r02_issue_mesh1m.log
ls -la gpurun_out /tmp/*.ncu-rep
}

# parity after the k_generate / k_film trims, then the occupancy A/B of the lean shade kernels
call5() {
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_variety.py -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c5_pytest.log 2>&1; tail -4 gpurun_out/r02_c5_pytest.log
rm -f gpurun_out/r02_ab_mb.log
bash tools/r02_ab_mb.sh
}

# round 2, call 6: racecheck / synccheck, the five-config results table, instruction counts of the current build
call6() {
bash tools/sanitize_race.sh > gpurun_out/r02_sanitize_summary.txt 2>&1; cat gpurun_out/r02_sanitize_summary.txt
python tools/results_table.py --out gpurun_out/r02_results.json --md gpurun_out/r02_results.md > gpurun_out/r02_results.log 2>&1; cat gpurun_out/r02_results.md
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_composite.csv \
    python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_issue_composite.log 2>&1
ncu --metrics $M --clock-control none -k regex:k_trace --csv --log-file gpurun_out/r02_issue_mesh1m.csv \
    python tools/render_once.py --workload mesh1m --warm 0 > gpurun_out/r02_issue_mesh1m.log 2>&1
cat gpurun_out/r02_issue_composite.log gpurun_out/r02_issue_mesh1m.log
}

# two wavefronts in flight (SG_OVERLAP=2, the new default) against one (SG_OVERLAP=1), then the GPU suite on the default
call7() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_overlap.log
for W in "composite --spp 64 --reps 2" "mesh1m --reps 3" "glass --reps 1" "instanced --reps 1" "cornell --reps 3"; do
  for O in 1 2; do
    echo "== $W SG_OVERLAP=$O" >> gpurun_out/r02_overlap.log
    SG_OVERLAP=$O python tools/perf_ab.py --workload $W base >> gpurun_out/r02_overlap.log 2>> gpurun_out/r02_overlap.err
  done
done
cat gpurun_out/r02_overlap.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c7_pytest.log 2>&1; tail -4 gpurun_out/r02_c7_pytest.log
}

# wavefronts in flight: 1 / 2 / 3 / 4
call8() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_overlap2.log
for W in "composite --spp 64 --reps 2" "instanced --reps 1" "glass --reps 1" "cornell --reps 3" "mesh1m --reps 2"; do
  for O in 1 2 3 4; do
    echo "== $W SG_OVERLAP=$O" >> gpurun_out/r02_overlap2.log
    SG_OVERLAP=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_overlap2.err | cut -c1-110 >> gpurun_out/r02_overlap2.log
  done
done
cat gpurun_out/r02_overlap2.log
}

# shadow traversal of depth d on a side stream, overlapping the closest-hit traversal of depth d + 1: off / on
call9() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_side.log
for W in "mesh1m --reps 3" "composite --spp 64 --reps 2" "instanced --reps 1" "glass --reps 1" "cornell --reps 3"; do
  for O in 0 1; do
    echo "== $W SG_SHADOW_SIDE_STREAM=$O" >> gpurun_out/r02_side.log
    SG_SHADOW_SIDE_STREAM=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_side.err | cut -c1-110 >> gpurun_out/r02_side.log
  done
done
cat gpurun_out/r02_side.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c9_pytest.log 2>&1; tail -4 gpurun_out/r02_c9_pytest.log
}

# C4: full ncu capture of the textured shade kernel and the instanced traversal kernel (first launches of one batch)
call10() {
ncu --set full --clock-control none --import-source on -k regex:k_shade -c 2 -o gpurun_out/r02_c4_shade \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4_shade.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r02_c4_trace \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4_trace.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_instanced.csv \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_launches_instanced.log 2>&1
ls -la gpurun_out
}

# shade queues re-ordered by material id (textured scenes): off / on, then the textured parity tests
call11() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_sortmat.log
for O in 0 1; do
  echo "== instanced SG_SORT_MATERIALS=$O" >> gpurun_out/r02_sortmat.log
  SG_SORT_MATERIALS=$O python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_sortmat.err >> gpurun_out/r02_sortmat.log
done
cat gpurun_out/r02_sortmat.log
python -m pytest tests -m gpu -x -q -k "tex or variety or instanced or mix or configs and not c4_converged" > gpurun_out/r02_c11_pytest.log 2>&1; tail -4 gpurun_out/r02_c11_pytest.log
}

# instanced traversal with the render-space ray parked in shared memory: C4 timing + instancing / ray-cast parity tests
call12() {
python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_c12.err | tee gpurun_out/r02_c12.log
python -m pytest tests -m gpu -x -q -k "inst or raycast or sphere or patch or configs and not c4_converged" > gpurun_out/r02_c12_pytest.log 2>&1; tail -4 gpurun_out/r02_c12_pytest.log
}

# C4 after the material sort + parked ray: ncu of the shade / sort kernels and the traversal kernels
call13() {
ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_sort' -c 6 -o /tmp/r02_c4b_shade \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4b_shade.log 2>&1
ncu -i /tmp/r02_c4b_shade.ncu-rep --page raw --csv > gpurun_out/r02_c4b_shade_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_trace -c 3 -o gpurun_out/r02_c4b_trace \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_c4b_trace.log 2>&1
ncu -i gpurun_out/r02_c4b_trace.ncu-rep --page raw --csv > gpurun_out/r02_c4b_trace_raw.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_instanced_b.csv \
    python tools/render_once.py --workload instanced --spp 8 --warm 0 > gpurun_out/r02_launches_instanced_b.log 2>&1
ls -la gpurun_out
}

# staged shading of textured scenes: off / on (C4), then the GPU suite
call14() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_staged.log
for O in 0 1; do
  echo "== instanced SG_STAGED_SHADING=$O" >> gpurun_out/r02_staged.log
  SG_STAGED_SHADING=$O python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_staged.err >> gpurun_out/r02_staged.log
done
cat gpurun_out/r02_staged.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c14_pytest.log 2>&1; tail -4 gpurun_out/r02_c14_pytest.log
}

call15() {
python tools/perf_ab.py --workload instanced --reps 2 base 2>> gpurun_out/r02_c15.err | cut -c1-170 | tee gpurun_out/r02_c15.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c15_pytest.log 2>&1; tail -4 gpurun_out/r02_c15_pytest.log
}

# order-preserving shade queues: off / on for every configuration, then the GPU suite
call16() {
mkdir -p gpurun_out; rm -f gpurun_out/r02_ordered.log
for W in "mesh1m --reps 3" "composite --spp 64 --reps 2" "glass --reps 1" "instanced --reps 1" "cornell --reps 3"; do
  for O in 0 1; do
    echo "== $W SG_ORDERED_QUEUES=$O" >> gpurun_out/r02_ordered.log
    SG_ORDERED_QUEUES=$O python tools/perf_ab.py --workload $W base 2>> gpurun_out/r02_ordered.err | cut -c1-170 >> gpurun_out/r02_ordered.log
  done
done
cat gpurun_out/r02_ordered.log
python -m pytest tests -m gpu -x -q -k "not c4_converged" > gpurun_out/r02_c16_pytest.log 2>&1; tail -4 gpurun_out/r02_c16_pytest.log
}

call17() {
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
}

call18() {
SG_OVERLAP=1 ncu --set full --clock-control none -k regex:'k_shade|k_trace' -c 9 -o /tmp/r02_final_c5 python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_final_c5.log 2>&1
ncu -i /tmp/r02_final_c5.ncu-rep --page raw --csv > gpurun_out/r02_final_c5_raw.csv 2>/dev/null
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_composite.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > gpurun_out/r02_launches_composite.log 2>&1
}

# ray-order experiment: what would sorting the ray queue buy the closest-hit kernel?
call19() {
timeout 600 python tools/exp_ray_order.py --workload mesh1m --spp 16 > gpurun_out/r02_ray_order_c2.log 2>&1
timeout 600 python tools/exp_ray_order.py --workload composite --spp 8 > gpurun_out/r02_ray_order_c5.log 2>&1
tail -30 gpurun_out/r02_ray_order_c2.log gpurun_out/r02_ray_order_c5.log
}

# source-level capture of the lean shade kernels at the order-preserving-queue build: depth-0 diffuse, depth-0 conductor, depth-1 diffuse
call20() {
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -f -o gpurun_out/r02_shade_src \
   python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_shade_src.log 2>&1
ls -la gpurun_out/r02_shade_src.ncu-rep
}

# stage barriers + CTA size of the lean shade kernels (instruction-fetch sharing between the warps of an SM)
call21() {
L=gpurun_out/r02_shade_sync.log; : > $L
for V in base t256 t512; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V composite" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_SYNC=16 SG_SHADE_SYNC=4 SG_SHADE_SYNC=12 SG_SHADE_SYNC=31 2>> gpurun_out/r02_shade_sync.err | cut -c1-200 >> $L
  echo "== $V mesh1m" >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_SHADE_SYNC=12 SG_SHADE_SYNC=31 2>> gpurun_out/r02_shade_sync.err | cut -c1-200 >> $L
done
cat $L
}

# 256-thread lean shade CTAs + barriers around sample_ld as defaults: GPU tests, all five configs, textured-kernel barrier sweep
call22() {
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c22_pytest.log 2>&1; tail -3 gpurun_out/r02_c22_pytest.log
L=gpurun_out/r02_shade_sync2.log; : > $L
echo "== instanced" >> $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base SG_SHADE_SYNC_TEX=2 SG_SHADE_SYNC_TEX=4 SG_SHADE_SYNC_TEX=6 SG_SHADE_SYNC_TEX=14 SG_SHADE_SYNC=0 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== glass" >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 SG_SHADE_SYNC=8 SG_SHADE_SYNC=14 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== cornell" >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
echo "== composite" >> $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_SYNC=0 SG_SHADE_SYNC=4 SG_SHADE_SYNC=6 SG_SHADE_SYNC=14 2>> gpurun_out/r02_shade_sync2.err | cut -c1-200 >> $L
cat $L
}

# source-level capture of the lean shade kernels with 256-thread CTAs + stage barriers: depth-0 diffuse, depth-0 conductor, depth-1 diffuse
call23() {
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_shade -c 3 -f -o gpurun_out/r02_shade_src2 \
   python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_shade_src2.log 2>&1
ls -la gpurun_out/r02_shade_src2.ncu-rep
}

# occupancy of the lean shade kernels again, now that they are no longer instruction-fetch bound:
# base = 2 x 256 threads (16 warps, <= 128 regs), t320 = 2 x 320 (20 warps, 96 regs), b3 = 3 x 256 (24 warps, 80 regs), b4 = 4 x 256 (32 warps, 64 regs)
call24() {
L=gpurun_out/r02_shade_occ.log; : > $L
for V in base t320 b3 b4; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_shade_occ.err | cut -c1-200 >> $L
done
cat $L
}

# ncu --set full of the streaming kernels of a C5 batch: k_generate, k_film, the queue compaction
call25() {
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_generate|k_film|k_queue' -c 5 -f -o gpurun_out/r02_stream_src \
   python tools/render_once.py --workload composite --spp 8 --warm 0 > gpurun_out/r02_stream_src.log 2>&1
ls -la gpurun_out/r02_stream_src.ncu-rep
}

# k_generate without 64-bit divisions + one-pass 8-way k_queue_scan: GPU tests, five configs
call26() {
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c26_pytest.log 2>&1; tail -3 gpurun_out/r02_c26_pytest.log
L=gpurun_out/r02_c26_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base 2>> gpurun_out/r02_c26.err | cut -c1-200 >> $L
cat $L
}

# depth-0 closest-hit launch reads path = index (no queue indirection); pf = software prefetch of the next item's load chain in the lean shade kernels;
# racecheck / synccheck of the barrier-synchronised shade kernels
call27() {
L=gpurun_out/r02_c27_perf.log; : > $L
for V in base pf; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 400 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
  timeout 400 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c27.err | cut -c1-200 >> $L
done
unset SHIMMER_GPU_LIB
cat $L
bash tools/sanitize_race.sh
}

# scheduling-invariance test + full GPU suite; shade grid size sweep
call28() {
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c28_pytest.log 2>&1; tail -3 gpurun_out/r02_c28_pytest.log
L=gpurun_out/r02_c28_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_SHADE_GRID=4 SG_SHADE_GRID=6 SG_SHADE_GRID=12 SG_SHADE_GRID=16 SG_SHADE_GRID=32 2>> gpurun_out/r02_c28.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_SHADE_GRID=4 SG_SHADE_GRID=16 SG_SHADE_GRID=32 2>> gpurun_out/r02_c28.err | cut -c1-200 >> $L
cat $L
}

# wavefront width on C5 again (two wavefronts in flight): 32 / 64 (default) / 128 Mi paths per wavefront; refill thresholds on the new build
call29() {
L=gpurun_out/r02_c29_perf.log; : > $L
timeout 900 python tools/perf_ab.py --workload composite --spp 128 --reps 2 base PIF=33554432 PIF=134217728 SG_REFILL_THRESHOLD=18 SG_REFILL_THRESHOLD=18,SG_INTERIOR_BURST=6 2>> gpurun_out/r02_c29.err | cut -c1-200 >> $L
cat $L; tail -3 gpurun_out/r02_c29.err
}

# 2 GPUs: the multi-GPU tests (both forms behind the C ABI) and bench.py --gpus 2 the way the driver launches it
call30() {
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r02_c30_multi.log 2>&1; tail -3 gpurun_out/r02_c30_multi.log
bash tools/r02_bench_n.sh 2 5 3 | tail -c 600
}

# k_film: segmented butterfly reduction over aligned runs of one pixel
call31() {
L=gpurun_out/r02_c31_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload instanced --spp 16 --reps 2 base 2>> gpurun_out/r02_c31.err | cut -c1-200 >> $L
cat $L
SG_OVERLAP=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_film|k_generate' --csv --log-file gpurun_out/r02_c31_film.csv \
    python tools/render_once.py --workload composite --spp 16 --warm 0 > /dev/null 2>&1
grep -E "k_film|k_generate" gpurun_out/r02_c31_film.csv | cut -d, -f5,15-
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scheduling.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
}

call32() {
timeout 600 python tools/kernel_trace.py --workload composite --spp 16 > gpurun_out/r02_ktrace_c5.log 2>&1; tail -20 gpurun_out/r02_ktrace_c5.log
timeout 600 python tools/kernel_trace.py --workload mesh1m > gpurun_out/r02_ktrace_c2.log 2>&1; tail -14 gpurun_out/r02_ktrace_c2.log
}

# textured shade kernels: CTAs of 256 threads (x2 = 2 per SM / 128 regs, x3 = 3 per SM / 80 regs) with stage barriers, against 128 x 5
call33() {
L=gpurun_out/r02_c33_perf.log; : > $L
for V in base x2 x3; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> $L
  timeout 600 python tools/perf_ab.py --workload instanced --spp 32 --reps 2 base SG_SHADE_SYNC_TEX=2 SG_SHADE_SYNC_TEX=3 SG_SHADE_SYNC_TEX=19 2>> gpurun_out/r02_c33.err | cut -c1-200 >> $L
done
cat $L
}

call34() {
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c34_pytest.log 2>&1; tail -3 gpurun_out/r02_c34_pytest.log
L=gpurun_out/r02_c34_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base 2>> gpurun_out/r02_c34.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c34.err | cut -c1-200 >> $L
cat $L
}

# co-residency of the issue-bound traversal kernels and the latency-bound shade kernels of the two wavefronts: cap the persistent traversal grid
# (9 CTAs per SM fill the register file) and size the shade grid to what fits beside it
call35() {
L=gpurun_out/r02_c35_perf.log; : > $L
timeout 900 python tools/perf_ab.py --workload composite --spp 128 --reps 2 base SG_TRACE_BLOCKS_PER_SM=7 SG_TRACE_BLOCKS_PER_SM=5 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=2 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=4 SG_TRACE_BLOCKS_PER_SM=6,SG_SHADE_GRID=2 SG_TRACE_BLOCKS_PER_SM=5,SG_SHADE_GRID=2,SG_OVERLAP=3 2>> gpurun_out/r02_c35.err | cut -c1-200 >> $L
cat $L
}

# round 2, call 1: wavefront-width sweep + scheduling-knob sweep on the round-1 build
sweep1() {
python tools/perf_ab.py --workload mesh1m base PIF=1048576 PIF=4194304 PIF=16777216 PIF=67108864 \
  SG_REFILL_THRESHOLD=4 SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=12 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=24 \
  SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=12 SG_LEAF_THRESHOLD=16 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=8 \
  > gpurun_out/r02_sweep_c2.log 2> gpurun_out/r02_sweep_c2.err
python tools/perf_ab.py --workload instanced --reps 1 base PIF=1048576 PIF=4194304 PIF=16777216 PIF=67108864 \
  > gpurun_out/r02_sweep_c4.log 2> gpurun_out/r02_sweep_c4.err
python tools/perf_ab.py --workload composite --spp 128 --reps 1 base PIF=4194304 PIF=16777216 \
  > gpurun_out/r02_sweep_c5.log 2> gpurun_out/r02_sweep_c5.err
cat gpurun_out/r02_sweep_c2.log gpurun_out/r02_sweep_c4.log gpurun_out/r02_sweep_c5.log
}

# round 2: scheduling knobs around the new default (refill 12) on C2 and on C5's per-GPU share at N = 8 (128 spp)
sweep2() {
python tools/perf_ab.py --workload mesh1m base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=14 SG_REFILL_THRESHOLD=16 \
  SG_REFILL_THRESHOLD=14,SG_LEAF_THRESHOLD=6 SG_REFILL_THRESHOLD=14,SG_LEAF_THRESHOLD=10 SG_REFILL_THRESHOLD=12,SG_LEAF_THRESHOLD=6 \
  SG_REFILL_THRESHOLD=12,SG_INTERIOR_BURST=3 SG_REFILL_THRESHOLD=12,SG_INTERIOR_BURST=6 SG_REFILL_THRESHOLD=16,SG_INTERIOR_BURST=6 \
  SG_SMEM_LEVELS=16 SG_SMEM_LEVELS=24 SG_PREFETCH=1 > gpurun_out/r02_sweep2_c2.log 2> gpurun_out/r02_sweep2_c2.err
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=20 \
  SG_REFILL_THRESHOLD=16,SG_LEAF_THRESHOLD=6 SG_REFILL_THRESHOLD=16,SG_LEAF_THRESHOLD=12 SG_INTERIOR_BURST=8 > gpurun_out/r02_sweep2_c5.log 2> gpurun_out/r02_sweep2_c5.err
cat gpurun_out/r02_sweep2_c2.log gpurun_out/r02_sweep2_c5.log
}

# scheduling knobs of the INSTANCED traversal kernels on C4 (the round-2 defaults were tuned on the triangle-only kernels)
sweep3() {
python tools/perf_ab.py --workload instanced --reps 1 base SG_LEAF_THRESHOLD=2 SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=8 SG_LEAF_THRESHOLD=12 SG_LEAF_THRESHOLD=16 \
  SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=20 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=8 SG_LEAF_THRESHOLD=10,SG_INTERIOR_BURST=8 \
  2> gpurun_out/r02_sweep3_c4.err | cut -c1-170 | tee gpurun_out/r02_sweep3_c4.log
}

sweep4() {
python tools/perf_ab.py --workload instanced --reps 1 SG_INTERIOR_BURST=1 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=3 SG_INTERIOR_BURST=1,SG_REFILL_THRESHOLD=20 SG_INTERIOR_BURST=2,SG_REFILL_THRESHOLD=20 \
  SG_INTERIOR_BURST=2,SG_REFILL_THRESHOLD=24 SG_INTERIOR_BURST=1,SG_REFILL_THRESHOLD=20,SG_LEAF_THRESHOLD=4 SG_INTERIOR_BURST=2,SG_REFILL_THRESHOLD=20,SG_LEAF_THRESHOLD=4 SG_INTERIOR_BURST=2,SG_REFILL_THRESHOLD=20,SG_LEAF_THRESHOLD=8 \
  2> gpurun_out/r02_sweep4_c4.err | cut -c1-170 | tee gpurun_out/r02_sweep4_c4.log
}

# depth-0 closest-hit knobs (camera rays) on C5 (64 spp), C2 and C4
sweep5() {
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD_D0=20 SG_REFILL_THRESHOLD_D0=24 SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 \
   SG_REFILL_THRESHOLD_D0=28,SG_INTERIOR_BURST_D0=8 SG_REFILL_THRESHOLD_D0=28,SG_INTERIOR_BURST_D0=2 2> gpurun_out/r02_sweep5_c5.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c5.log
python tools/perf_ab.py --workload mesh1m --reps 3 base SG_REFILL_THRESHOLD_D0=24 SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 2> gpurun_out/r02_sweep5_c2.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c2.log
python tools/perf_ab.py --workload instanced --reps 1 base SG_REFILL_THRESHOLD_D0=28 SG_REFILL_THRESHOLD_D0=32 SG_INTERIOR_BURST_D0=2 2> gpurun_out/r02_sweep5_c4.err | cut -c1-170 | tee gpurun_out/r02_sweep5_c4.log
}

# refill / leaf / burst of the depth >= 1 launches again, now that the shade queues no longer depend on the retire order
sweep6() {
python tools/perf_ab.py --workload mesh1m --reps 3 base SG_REFILL_THRESHOLD=8 SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=12 SG_REFILL_THRESHOLD=16 SG_REFILL_THRESHOLD=20 \
  SG_LEAF_THRESHOLD=4 SG_LEAF_THRESHOLD=8 SG_INTERIOR_BURST=3 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c2.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c2.log
python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=18 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c5.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c5.log
python tools/perf_ab.py --workload glass --reps 1 base SG_REFILL_THRESHOLD=10 SG_REFILL_THRESHOLD=18 SG_INTERIOR_BURST=2 SG_INTERIOR_BURST=6 2> gpurun_out/r02_sweep6_c3.err | cut -c1-170 | tee gpurun_out/r02_sweep6_c3.log
}

# A/B: resident blocks per SM of the lean shade kernels (4 = default build, 5 / 6 / 8 = register caps 96 / 80 / 64)
ab_mb() {
for V in base mb5 mb6 mb8; do
  if [ $V = base ]; then unset SHIMMER_GPU_LIB; else export SHIMMER_GPU_LIB=$PWD/shimmer_b200/ab/libshimmer_gpu_$V.so; fi
  echo "== $V" >> gpurun_out/r02_ab_mb.log
  python tools/perf_ab.py --workload composite --spp 64 --reps 2 base >> gpurun_out/r02_ab_mb.log 2>> gpurun_out/r02_ab_mb.err
  python tools/perf_ab.py --workload mesh1m --reps 3 base >> gpurun_out/r02_ab_mb.log 2>> gpurun_out/r02_ab_mb.err
done
cat gpurun_out/r02_ab_mb.log
}

# spectrum interval table (piecewise-linear spectra): GPU tests + five configs
call36() {
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_c36_pytest.log 2>&1; tail -3 gpurun_out/r02_c36_pytest.log
L=gpurun_out/r02_c36_perf.log; : > $L
timeout 600 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c36.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload mesh1m --reps 2 base 2>> gpurun_out/r02_c36.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload glass --reps 1 base 2>> gpurun_out/r02_c36.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload instanced --reps 1 base 2>> gpurun_out/r02_c36.err | cut -c1-200 >> $L
timeout 600 python tools/perf_ab.py --workload cornell --reps 3 base 2>> gpurun_out/r02_c36.err | cut -c1-200 >> $L
cat $L
}

# source-level capture of the depth-1 traversal kernels on C2 (profiles/r02_trace_ncu.md)
call37() {
SG_OVERLAP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 2 -f -o gpurun_out/r02_trace_src \
   python tools/render_once.py --workload mesh1m --spp 16 --warm 0 > gpurun_out/r02_trace_src.log 2>&1
ls -la gpurun_out/r02_trace_src.ncu-rep
}

# two rays per lane (k_trace_dual): scheduling-invariance test, C2 / C5 A/B
call38() {
timeout 300 python -m pytest tests/test_gpu_scheduling.py -m gpu -x -q > gpurun_out/r02_c38_pytest.log 2>&1; tail -5 gpurun_out/r02_c38_pytest.log
L=gpurun_out/r02_c38_perf.log; : > $L
timeout 300 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_TRACE_DUAL=1 SG_TRACE_DUAL=2 SG_TRACE_DUAL=3 SG_TRACE_DUAL=3,SG_DUAL_LEAF=8,SG_DUAL_REFILL=16 SG_TRACE_DUAL=3,SG_DUAL_LEAF=16,SG_DUAL_REFILL=32 SG_TRACE_DUAL=3,SG_DUAL_LEVELS=16 2>> gpurun_out/r02_c38.err | cut -c1-200 >> $L
timeout 300 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base SG_TRACE_DUAL=3 2>> gpurun_out/r02_c38.err | cut -c1-200 >> $L
cat $L
}

# two rays per lane: stack levels / thresholds, instruction counts of the depth-1 closest-hit launch with and without
call39() {
L=gpurun_out/r02_c39_perf.log; : > $L
timeout 300 python tools/perf_ab.py --workload mesh1m --reps 2 base SG_TRACE_DUAL=1,SG_DUAL_LEVELS=6 SG_TRACE_DUAL=1,SG_DUAL_LEVELS=8 SG_TRACE_DUAL=1,SG_DUAL_LEVELS=10 SG_TRACE_DUAL=1,SG_DUAL_LEVELS=8,SG_DUAL_LEAF=20,SG_DUAL_REFILL=40 SG_TRACE_DUAL=1,SG_DUAL_LEVELS=8,SG_INTERIOR_BURST=8 2>> gpurun_out/r02_c39.err | cut -c1-200 >> $L
cat $L
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers
for D in 0 1; do
SG_TRACE_DUAL=$D SG_OVERLAP=1 timeout 300 ncu --metrics $M --clock-control none -k regex:k_trace -s 2 -c 1 --csv --log-file gpurun_out/r02_c39_dual$D.csv python tools/render_once.py --workload mesh1m --spp 16 --warm 0 > /dev/null 2>&1
grep -E "k_trace" gpurun_out/r02_c39_dual$D.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
done
}

# vote-loop trims of the triangle-only traversal kernels: parity + scheduling tests, C2 / C5 / C4
call40() {
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scheduling.py -m gpu -x -q > gpurun_out/r02_c40_pytest.log 2>&1; tail -2 gpurun_out/r02_c40_pytest.log
L=gpurun_out/r02_c40_perf.log; : > $L
timeout 300 python tools/perf_ab.py --workload mesh1m --reps 3 base 2>> gpurun_out/r02_c40.err | cut -c1-200 >> $L
timeout 300 python tools/perf_ab.py --workload composite --spp 64 --reps 2 base 2>> gpurun_out/r02_c40.err | cut -c1-200 >> $L
timeout 300 python tools/perf_ab.py --workload instanced --spp 32 --reps 2 base 2>> gpurun_out/r02_c40.err | cut -c1-200 >> $L
cat $L
}

if [ "$1" = "list" ] || [ -z "$1" ]; then echo call2 call3 call4 call5 call6 call7 call8 call9 call10 call11 call12 call13 call14 call15 call16 call17 call18 call19 call20 call21 call22 call23 call24 call25 call26 call27 call28 call29 call30 call31 call32 call33 call34 call35 call36 call37 call38 call39 call40 sweep1 sweep2 sweep3 sweep4 sweep5 sweep6 ab_mb; else "$@"; fi
