#!/usr/bin/env python3
"""Print a compact per-launch metric table from an `ncu --page raw --csv` export.  usage: ncu_raw_table.py raw.csv [--md]"""
import csv, sys
WANT = [("gpu__time_duration.sum", "time ms", None), ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1),
        ("launch__block_size", "block", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst", 1),
        ("smsp__inst_executed.sum", "warp inst (M)", 1e-6),
        ("dram__bytes_read.sum", "dram rd MB", None), ("dram__bytes_write.sum", "dram wr MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", 1),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_sel", 1),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math", 1),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_thr", 1),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier", 1),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch", 1),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_inst", 1),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch", 1),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio", 1),
        ("smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio", "stall imc", 1),
        ("local_load_bytes", "", 1),
        ("smsp__inst_executed_op_local_ld.sum", "local ld inst", 1), ("smsp__inst_executed_op_local_st.sum", "local st inst", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %", 1),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %", 1),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu pipe %", 1)]
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]; units = rows[1]; data = rows[2:]
idx = {n: i for i, n in enumerate(h)}
names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("sg::", "") for r in data]
print("| metric | " + " | ".join("#%d" % i for i in range(len(data))) + " |")
print("|---|" + "---:|" * len(data))
print("| kernel | " + " | ".join(names) + " |")
for key, label, scale in WANT:
    if key not in idx:
        continue
    vals = []
    for r in data:
        try:
            v = float(r[idx[key]].replace(",", ""))
        except ValueError:
            vals.append(r[idx[key]]); continue
        if scale is None:          # bytes with unit column
            u = units[idx[key]].lower()
            v = v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6,
                     "ms": 1.0, "msecond": 1.0, "second": 1e3}.get(u, 1.0)
        else:
            v *= scale
        vals.append("%.3f" % v)
    print("| %s | " % label + " | ".join(vals) + " |")
