#!/usr/bin/env python3
"""Rewrite BASELINE.md section 5 from measured artefacts: the results table (tools/results_table.py -> results.json),
the bench lines (profiles/*_bench.json) and the multi-GPU runs (profiles/*_scale_*.json).
usage: update_baseline_md.py results.json [tag]"""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = json.load(open(sys.argv[1])); tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
out = ["## 5. Results\n",
       "All rows measured on this pod's B200 (SM clock 1965 MHz, no throttle reasons), round 1.  GPU numbers: one full render of the",
       "configuration with the film resident in HBM (`tools/results_table.py`; CUDA events).  CPU numbers: the C++ oracle (\"reference",
       "restated, not the reference binary\") on the GPU box's host cores, tile-parallel in the reference's RNG mode, on a 96x96 pixel",
       "window of the same scene at full spp (window below the image centre, where paths are longest -- a conservative CPU figure for",
       "C5, whose full frame is mostly background).  Image agreement is computed inside that window on developed RGB",
       "(film.rs:720-738): *same-stream* = GPU vs oracle with identical per-(pixel, sample) random streams (implementation parity);",
       "*independent* = GPU vs oracle in the reference's sequential-RNG mode, i.e. different random numbers, noise-limited at the",
       "config's spp (the converged <= 1 % RMSE bar is tested at 131072 spp in `tests/test_gpu_parity.py`).\n",
       "| config | GPUs | Mrays/s | Mpaths/s | ms / render | CPU threads | CPU Mrays/s | CPU Mpaths/s | RMSE same-stream | rel-lum err same-stream | RMSE independent | rel-lum err independent | traversal roofline fraction (HBM, SURVEY 8d bytes) |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
names = {"cornell": "C1", "mesh1m": "C2", "glass": "C3", "instanced": "C4", "composite": "C5"}
for r in rows:
    out.append("| %s %s %dx%d %d spp | 1 | %.0f | %.1f | %.1f | %d | %.2f | %.2f | %.1e | %.1e | %.3f | %.4f | %.3f |" % (
        names.get(r["config"], "?"), r["config"], r["resolution"][0], r["resolution"][1], r["spp"], r["gpu_mrays"], r["gpu_mpaths"], r["gpu_ms"],
        r["cpu_threads"], r["cpu_mrays"], r["cpu_mpaths"], r["rmse_same"], r["dlum_same"], r["rmse_indep"], r["dlum_indep"], r["trav_frac"]))
out.append("")
out.append("Scene facts: " + "; ".join("%s %s triangles (%s instanced), %s lights, %.1f nodes + %.2f triangles tested per closest-hit ray" % (
    names.get(r["config"], "?"), r["triangles"], r["instanced_triangles"], r["lights"], r["nodes_per_ray"], r["tris_per_ray"]) for r in rows) + ".\n")
sc = []
for f in sorted(glob.glob(os.path.join(ROOT, "profiles", tag + "_scale_*.json"))):
    d = json.loads(open(f).read())
    sc.append((os.path.basename(f), d))
bench = None
bf = os.path.join(ROOT, "profiles", tag + "_bench.json")
if os.path.exists(bf):
    bench = json.loads(open(bf).read().strip().splitlines()[-1])
out.append("### Multi-GPU (one process per GPU, scene replicated, sample-range split, one NCCL film reduce per render; `bench.py` under torchrun)\n")
out.append("| workload | GPUs | spp per GPU | Mpaths/s (whole job) | Mrays/s | ms / step | e2e Mpaths/s (host film) | efficiency vs N=1 |")
out.append("|---|---:|---:|---:|---:|---:|---:|---:|")
base1 = {"mesh1m": bench["value"] if bench else None}
for r in rows:
    if r["config"] == "composite": base1["composite"] = r["gpu_mpaths"]
if bench:
    out.append("| C2 mesh1m (weak) | 1 | %d | %.1f | %.0f | %.1f | %.1f | 1.000 |" % (bench["config"]["spp_per_gpu"], bench["value"], bench["mrays_per_s"], bench["ms_per_step"], bench["e2e"]["value"]))
for name, d in sorted(sc, key=lambda x: (x[1]["config"]["workload"][:4], x[1]["n_gpus"])):
    wl = d["config"]["workload"].split(":")[0]
    b1 = base1.get(wl)
    eff = d["value"] / (b1 * d["n_gpus"]) if b1 else float("nan")
    label = "C2 mesh1m (weak)" if wl == "mesh1m" else "C5 composite 4K, 1024 spp total (strong: %d x %d spp)" % (d["n_gpus"], d["config"]["spp_per_gpu"])
    old_def = d["n_gpus"] > 1 and d["e2e"].get("note", "").startswith("sg_render:")
    out.append("| %s | %d | %d | %.1f | %.0f | %.1f | %.1f%s | %.3f |" % (label, d["n_gpus"], d["config"]["spp_per_gpu"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], " (*)" if old_def else "", eff))
out.append("")
out.append("e2e at N > 1 = `sg_render_device` on every rank + the NCCL film reduce + one D2H of the reduced film into rank 0's pinned host buffer (the whole")
out.append("job's film in host memory).  Rows marked (*) were measured before that definition, as per-rank `sg_render` calls without the reduce.")
out.append("")
if bench:
    r = bench["roofline"]
    out.append("### Headline bench line (`python bench.py`, C2, N=1; `profiles/%s_bench.json`)\n" % tag)
    out.append("value %.1f Mpaths/s (%.0f Mrays/s, %.1f ms/step), e2e %.1f Mpaths/s with the film read back to host memory every step; "
               "closest-hit traversal %.0f Mrays/s = %.0f GB/s algorithmic = **%.3f of the measured HBM peak** (%.0f GB/s), any-hit %.0f Mrays/s; "
               "traversal %.0f %% + %.0f %% of the step; CPU oracle beside it: %.2f Mpaths/s on %d threads.\n" % (
        bench["value"], bench["mrays_per_s"], bench["ms_per_step"], bench["e2e"]["value"], r["closest_mrays_per_s"], r["achieved"], r["frac"], r["peak"],
        r["shadow_mrays_per_s"], 100 * r["kernel_share_of_step"], 100 * r["shadow_share_of_step"],
        (bench.get("cpu_baseline") or {}).get("value", float("nan")), (bench.get("cpu_baseline") or {}).get("cores", 0)))
p = os.path.join(ROOT, "BASELINE.md"); s = open(p).read()
i = s.index("## 5. Results")
j = s.find("\n## ", i + 5)
s = s[:i] + "\n".join(out) + ("\n" + s[j:] if j >= 0 else "\n")
open(p, "w").write(s)
print("BASELINE.md section 5 rewritten (%d config rows, %d scaling rows)" % (len(rows), len(sc)))
