#!/usr/bin/env python3
"""Rewrite BASELINE.md section 5 from measured artefacts: the five-config results table (tools/results_table.py ->
profiles/<tag>_results.json) and the bench lines at N = 1, 2, 4, 8 (profiles/<tag>_bench_n<N>.json, <tag>_bench_reference.json).
usage: update_baseline_md.py [tag]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = lambda name: os.path.join(ROOT, "profiles", "%s_%s" % (tag, name))
last = lambda f: json.loads([l for l in open(f) if l.startswith("{")][-1])
rows = json.load(open(P("results.json")))
r1 = {"cornell": 610.8, "mesh1m": 556.0, "glass": 538.0, "instanced": 188.7, "composite": 1576.0}      # round 1, same table
out = ["## 5. Results\n",
       "All rows measured on this pod's B200s (SM clock 1965 MHz, no throttle reasons), round 2.  GPU numbers: one full render of the",
       "configuration with the film resident in HBM (`tools/results_table.py`; CUDA events).  CPU numbers: the C++ oracle (\"reference",
       "restated, not the reference binary\") on the GPU box's host cores, tile-parallel in the reference's RNG mode, on a 96x96 pixel",
       "window of the same scene at full spp (below the image centre, where paths are longest -- a conservative CPU figure for C5, whose",
       "full frame is mostly background).  Image agreement is computed inside that window on developed RGB (film.rs:720-738):",
       "*same-stream* = GPU vs oracle with identical per-(pixel, sample) random streams (implementation parity; C3's 3e-3 is the reference",
       "algorithm's own one-ulp sensitivity on dielectric paths, DESIGN.md section 7); *independent* = GPU vs oracle in the reference's",
       "sequential-RNG mode, i.e. different random numbers, noise-limited at the config's spp (the converged <= 1 % bars are tested at high",
       "spp in `tests/test_gpu_parity.py` / `tests/test_gpu_configs.py` for C1, C3 and C4).\n",
       "| config | GPUs | Mrays/s | Mpaths/s | round 1 | ms / render | CPU threads | CPU Mrays/s | CPU Mpaths/s | RMSE same-stream | rel-lum err same-stream | RMSE independent | rel-lum err independent | traversal, SURVEY 8d byte-equivalent / HBM copy peak |",
       "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
names = {"cornell": "C1", "mesh1m": "C2", "glass": "C3", "instanced": "C4", "composite": "C5"}
for r in rows:
    out.append("| %s %s %dx%d %d spp | 1 | %.0f | **%.1f** | %.1f | %.1f | %d | %.2f | %.2f | %.1e | %.1e | %.3f | %.4f | %.3f |" % (
        names.get(r["config"], "?"), r["config"], r["resolution"][0], r["resolution"][1], r["spp"], r["gpu_mrays"], r["gpu_mpaths"], r1.get(r["config"], float("nan")),
        r["gpu_ms"], r["cpu_threads"], r["cpu_mrays"], r["cpu_mpaths"], r["rmse_same"], r["dlum_same"], r["rmse_indep"], r["dlum_indep"], r["trav_frac"]))
out.append("")
out.append("Scene facts: " + "; ".join("%s %s triangles (%s instanced), %s lights, %.1f nodes + %.2f triangles tested per closest-hit ray" % (
    names.get(r["config"], "?"), r["triangles"], r["instanced_triangles"], r["lights"], r["nodes_per_ray"], r["tris_per_ray"]) for r in rows) + ".\n")
out.append("### Strong scaling on C5 (`bench.py --gpus N`: 3840x2160, 1024 spp in total, 1024 / N sample indices per GPU; one process per GPU, the split and "
           "ONE in-library `ncclReduce` of the f64 film per step behind the C ABI)\n")
out.append("| GPUs | spp per GPU | Mpaths/s (whole job, device-timed) | Mrays/s | ms / step | e2e Mpaths/s (host film on rank 0) | efficiency vs N=1 | reduce ms (collective / as seen by rank 0 incl. waiting) | film D2H ms |")
out.append("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
b = {}
for n in (1, 2, 4, 8):
    if os.path.exists(P("bench_n%d.json" % n)):
        b[n] = last(P("bench_n%d.json" % n))
for n, d in sorted(b.items()):
    out.append("| %d | %g | %.1f | %.0f | %.1f | %.1f | %.3f | %.2f / %.2f | %.2f |" % (
        n, d["setup"]["spp_per_gpu"], d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], d["value"] / (n * b[1]["value"]) if 1 in b else float("nan"),
        d["multi_gpu"]["reduce_ms_fastest_rank"], d["multi_gpu"]["reduce_ms_root"], d["e2e"]["d2h_ms_root"]))
out.append("")
if 8 in b and 1 in b:
    m8 = b[8]["multi_gpu"]
    out.append("The film is 265 MB (3840 x 2160 x 4 f64); its reduce over NVLink takes %.1f ms (%.1f ms as seen by rank 0, which also waits for the slowest rank) and the D2H "
               "%.1f ms against %.0f ms of rendering at N = 8: the scaling loss (%.1f %% at N = 8) is the slowest rank's wavefront loop (%.0f ms against %.0f ms for an "
               "ideal eighth of the one-GPU step: pipeline fill and drain of 16 batches per rank instead of 128), not the collective.  Builder-run, `--steps 3..5 --warmup 3`.\n" % (
        m8["reduce_ms_fastest_rank"], m8["reduce_ms_root"], b[8]["e2e"].get("d2h_ms_root", float("nan")), b[8]["ms_per_step"],
        100 * (1 - b[8]["value"] / (8 * b[1]["value"])), m8["render_ms_slowest_rank"], b[1]["ms_per_step"] / 8))
if 1 in b:
    d = b[1]; r = d["roofline"]; c2 = d.get("c2"); ref = last(P("bench_reference.json")) if os.path.exists(P("bench_reference.json")) else None
    out.append("### Headline bench line (`python bench.py`, C5, N=1; `profiles/%s_bench_n1.json`)\n" % tag)
    out.append("value %.1f Mpaths/s (%.0f Mrays/s, %.0f ms/step), e2e %.1f Mpaths/s through `sg_render` with the 265 MB film read back to host memory every step; "
               "closest-hit traversal %.0f Mrays/s = %.0f G warp-inst/s = **%.3f of the issue-slot peak** (%.0f G warp-inst/s = 148 SMs x 4 schedulers x 1965 MHz), %.1f of 32 lanes "
               "active per instruction, %.3f by the SURVEY 8d byte-equivalent against the measured HBM copy peak; traversal %.0f %% + %.0f %% (any-hit) of the step; "
               "CPU oracle beside it: %.2f Mpaths/s on %d threads%s.\n" % (
        d["value"], d["mrays_per_s"], d["ms_per_step"], d["e2e"]["value"], r["closest_mrays_per_s"], r["achieved"], r["frac"], r["peak"], r.get("lanes_per_inst", float("nan")),
        r["hbm_equivalent"]["frac"], 100 * r["kernel_share_of_step"], 100 * r["shadow_share_of_step"],
        (d.get("cpu_baseline") or {}).get("value", float("nan")), (d.get("cpu_baseline") or {}).get("cores", 0),
        (" (`--impl reference` arm: %.2f Mpaths/s)" % ref["value"]) if ref else ""))
    if c2:
        rr = c2["roofline"]
        out.append("C2 side measurement of the same line (round 1's headline workload): %.1f Mpaths/s (round 1: 555.4), closest-hit traversal %.0f Mrays/s = **%.3f of the issue-slot "
                   "peak**, %.1f lanes per instruction, %.3f byte-equivalent; traversal %.0f %% + %.0f %% of the step.\n" % (
            c2["value"], rr["closest_mrays_per_s"], rr["frac"], rr.get("lanes_per_inst", float("nan")), rr["hbm_equivalent"]["frac"], 100 * rr["kernel_share_of_step"], 100 * rr["shadow_share_of_step"]))
p = os.path.join(ROOT, "BASELINE.md"); s = open(p).read()
i = s.index("## 5. Results")
j = s.find("\n## ", i + 5)
s = s[:i] + "\n".join(out) + ("\n" + s[j:] if j >= 0 else "\n")
open(p, "w").write(s)
print("BASELINE.md section 5 rewritten (%d config rows, %d scaling rows)" % (len(rows), len(b)))
