#!/usr/bin/env python3
"""Build profiles/<tag>_ncu_summary.md from an ncu launch list (csv) and a `--set full` .ncu-rep.
usage: make_profile_summary.py TAG launches.csv prof.ncu-rep [bench.json]"""
import collections, csv, io, json, subprocess, sys

tag, launches, rep = sys.argv[1:4]
bench = sys.argv[4] if len(sys.argv) > 4 else None
out = [f"# {tag}: ncu evidence for the wavefront path tracer (B200, sm_100a)\n"]
out.append("Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py` "
           "(cold-cache, serialised: compare SHARES, not absolutes).\n")
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("sg::", "")
    agg[name][0] += 1; agg[name][1] += v / 1e6
tot = sum(v[1] for v in agg.values())
out.append("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.append(f"| `{k[:60]}` | {v[0]} | {v[1]:.3f} | {v[1] / tot:.3f} |")
out.append("")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]; idx = {x: i for i, x in enumerate(h)}
want = [("gpu__time_duration.sum", "time ms"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_throttle")]
out.append(f"`ncu --set full --clock-control none --import-source on -k regex:k_trace` ({rep.split('/')[-1]}), first traversal launches of a batch "
           "(depth-0 closest, depth-0 any-hit, depth-1 closest, depth-1 any-hit):\n")
out.append("| metric | " + " | ".join(f"#{i}" for i in range(len(rr) - 2)) + " |")
out.append("|---|" + "---:|" * (len(rr) - 2))
out.append("| kernel | " + " | ".join(r[idx["Kernel Name"]].split("(")[0].replace("void ", "") for r in rr[2:]) + " |")
for m, label in want:
    if m in idx:
        vals = []
        for r in rr[2:]:
            try:
                vals.append(f"{float(r[idx[m]].replace(',', '')):.3f}")
            except ValueError:
                vals.append(r[idx[m]])
        out.append(f"| {label} | " + " | ".join(vals) + " |")
out.append("")
if bench:
    d = json.loads([l for l in open(bench) if l.startswith("{")][-1])
    out.append("Bench line of the same build (NOT under a profiler):\n\n```json\n" + json.dumps(d, indent=1) + "\n```\n")
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("wrote", f"profiles/{tag}_ncu_summary.md")
