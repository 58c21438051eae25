#!/usr/bin/env python3
"""Run bench.py with the given args and print a one-line digest (for quick A/B runs under gpurun)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print("BENCH FAILED", out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
d = json.loads(line[-1]); r = d.get("roofline") or {}
print(" ".join(sys.argv[1:]), "| Mpaths/s %.1f Mrays/s %.1f e2e %.1f | closest %.0f Mr/s shadow %.0f Mr/s frac %.3f share %.2f/%.2f | ms/step %.1f launches %d clocks %s" % (
    d["value"], d["mrays_per_s"], d["e2e"]["value"], r.get("closest_mrays_per_s", 0), r.get("shadow_mrays_per_s", 0), r.get("frac", 0),
    r.get("kernel_share_of_step", 0), r.get("shadow_share_of_step", 0), d["ms_per_step"], d["gpu_launches"], d["clocks"]))
if d.get("cpu_baseline"):
    print("   cpu_baseline", d["cpu_baseline"])
