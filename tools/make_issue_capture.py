#!/usr/bin/env python3
"""profiles/r02_issue.json from an ncu metrics list of tools/render_once.py (run here, on the CSV gpurun brought back).

  python tools/make_issue_capture.py WORKLOAD gpurun_out/<csv> [--warm-renders 1] [--desc "..."]

The CSV comes from
  ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:k_trace --csv --log-file <csv> python tools/render_once.py --workload W --spp S --warm 0
and holds one row per (launch, metric).  The closest-hit launches are k_trace<0, ...> (ANY = false).  bench.py multiplies
`warp_inst_per_closest_ray` by the rays it traced and divides by the live CUDA-event time of the same kernels to get the
issue-slot roofline figure; instruction counts per ray do not depend on clocks or on the profiler's serialisation."""
import argparse
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload"); ap.add_argument("csv"); ap.add_argument("--log", default="")
    ap.add_argument("--desc", default="")
    ap.add_argument("--rays-depth0", type=int, default=0, help="rays of the first closest-hit launch (= pixels x spp of the capture)")
    args = ap.parse_args()
    rows = [r for r in csv.reader(open(args.csv)) if len(r) > 10]
    h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        d = per.setdefault(int(r[ii]), {"kernel": r[ki]})
        d[r[mi]] = v
    closest = [d for d in per.values() if re.search(r"k_trace<\(bool\)0|k_trace<false|k_trace<0", d["kernel"])]
    shadow = [d for d in per.values() if re.search(r"k_trace<\(bool\)1|k_trace<true|k_trace<1", d["kernel"])]
    if not closest:
        raise SystemExit("no closest-hit launches in " + args.csv)
    # rays per launch: from the render log (closest / shadow totals) when given
    n_closest = n_shadow = None
    if args.log:
        m = re.findall(r"closest (\d+) shadow (\d+) rays", open(args.log).read())
        if m:
            n_closest, n_shadow = int(m[-1][0]), int(m[-1][1])
    if not n_closest:
        raise SystemExit("need --log with the render_once.py output (ray totals)")
    wi = sum(d["smsp__inst_executed.sum"] for d in closest); ti = sum(d["smsp__thread_inst_executed.sum"] for d in closest)
    swi = sum(d["smsp__inst_executed.sum"] for d in shadow); sti = sum(d["smsp__thread_inst_executed.sum"] for d in shadow)
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = bool(subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "shimmer_b200/csrc", "include"], capture_output=True, text=True).stdout.strip())
    ent = {"commit": commit + ("+dirty" if dirty else ""), "capture_desc": args.desc,
           "how": "ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,... -k regex:k_trace over tools/render_once.py; "
                  "sums over the %d closest-hit launches of one render" % len(closest),
           "closest_launches": len(closest), "closest_rays": n_closest,
           "warp_inst_per_closest_ray": wi / n_closest, "thread_inst_per_closest_ray": ti / n_closest,
           "lanes_per_inst_closest": ti / wi,
           "lanes_per_inst_by_launch": [d["smsp__thread_inst_executed.sum"] / d["smsp__inst_executed.sum"] for d in closest],
           "shadow_rays": n_shadow, "warp_inst_per_shadow_ray": (swi / n_shadow) if n_shadow else None,
           "lanes_per_inst_shadow": (sti / swi) if swi else None,
           "rays_depth0_launch": args.rays_depth0 or None,
           "dram_bytes_depth0_launch": closest[0].get("dram__bytes_read.sum", 0.0) + closest[0].get("dram__bytes_write.sum", 0.0)}
    out = os.path.join(ROOT, "profiles", "r02_issue.json")
    try:
        cur = json.load(open(out))
    except Exception:
        cur = {}
    cur[args.workload] = ent
    json.dump(cur, open(out, "w"), indent=1)
    print(json.dumps(ent, indent=1))


if __name__ == "__main__":
    main()
