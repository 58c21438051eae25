#!/usr/bin/env python3
"""bench.py -- Mpaths/s (and Mrays/s) of the wavefront path tracer on BASELINE.json's C2 workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload mesh1m|cornell|glass]

One "step" = one full pass of the hot path (ImageTileIntegrator::render equivalent) over one batch of
synthetic input: every pixel of the workload x its samples-per-pixel.  N>1: one process per GPU
(torchrun), the scene is replicated, every rank renders its own `spp` samples of every pixel (distinct
sample indices -> weak scaling) and the film is summed onto rank 0 with one NCCL reduce per step.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how roofline/cpu_baseline are defined.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mesh1m", choices=["mesh1m", "cornell", "glass", "instanced", "composite"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (default: the config's)")
    ap.add_argument("--res", type=int, default=0, help="override square resolution (debug only)")
    ap.add_argument("--paths-in-flight", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_scene(args):
    from shimmer_b200 import scenes
    cfg = scenes.CONFIGS[args.workload]
    res = (args.res, args.res) if args.res else cfg["resolution"]
    spp = args.spp or cfg["spp"]
    t0 = time.time()
    sc = cfg["builder"](resolution=res).build()
    return sc, cfg, res, spp, time.time() - t0


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(sc, spp_sample, max_depth, threads, seed=0):
    """Times the CPU oracle (the restated reference: the Rust binary cannot be built in this image) in
    tile-parallel mode on `threads` host threads.  The ONLY place bench.py executes oracle/ code."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    p = orc.make_params(seed=seed, spp=spp_sample, max_depth=max_depth)
    film, st, secs = orc.render(sc, p, n_threads=threads, stream_mode=1)      # stream_mode 1 = reference behaviour
    return dict(secs=secs, paths=int(st.camera_paths), rays=int(st.closest_hit_rays + st.shadow_rays))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, cfg, res, spp, build_s = build_scene(args)
    threads = os.cpu_count() or 1
    # bounded sample of the same workload: all pixels, a quarter of the sample indices per step (~4 s on 16 cores)
    sample_spp = max(1, spp // 4)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
    t_paths = t_rays = 0; t_secs = 0.0
    for _ in range(args.steps):
        r = cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
        t_paths += r["paths"]; t_rays += r["rays"]; t_secs += r["secs"]
    v = t_paths / t_secs / 1e6
    line = {"impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mrays_per_s": t_rays / t_secs / 1e6,
            "config": {"workload": f"{args.workload}: {cfg['desc']}", "resolution": list(res), "spp": spp,
                       "integrator": "path maxdepth=%d, independent sampler, uniform light sampler" % cfg["max_depth"]},
            "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": threads, "kind": "port",
                             "sample": f"all {res[0]}x{res[1]} pixels x {sample_spp} spp per step (of {spp}); tile-parallel 8x8 tiles, "
                                       f"{threads} threads, reference RNG mode; oracle = C++ restatement (reference is Rust, not buildable here)"},
            "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from shimmer_b200 import Options, create_integrator

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sc, cfg, res, spp, build_s = build_scene(args)
    t0 = time.time()
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp}, device=local,
                              max_paths_in_flight=args.paths_in_flight)
    upload_s = time.time() - t0
    opts = Options(seed=0, pixel_samples=spp)
    W, H = integ.width, integ.height
    npix = W * H
    film = torch.zeros((npix, 4), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    from shimmer_b200.distributed import reduce_film, sample_range_for_rank
    my_range = sample_range_for_rank(spp, rank, world, "weak")   # weak scaling: every rank renders spp NEW sample indices

    def step(flags=0):
        film.zero_()
        integ.render_device(opts, film.data_ptr(), sample_range=my_range, stream=stream, flags=flags)
        reduce_film(film, dst=0)                      # one NCCL reduce of the f64 film per step (no-op at N=1)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local); clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    launches = 0; closest = shadow = 0
    for _ in range(args.steps):
        step()
        launches += int(integ.stats.kernel_launches); closest += int(integ.stats.closest_hit_rays); shadow += int(integ.stats.shadow_rays)
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t = torch.tensor([ms, float(closest + shadow), float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0]); rays_total = float(tsum[1]); launches_total = int(tsum[2])
    else:
        rays_total = float(t[1]); launches_total = launches
    paths_total = float(npix) * spp * world * args.steps
    value = paths_total / (ms * 1e-3) / 1e6
    mrays = rays_total / (ms * 1e-3) / 1e6

    # ---- e2e: the user-facing call with HOST buffers (sg_render: params in, film D2H inside the timed region)
    host_film = integ.film
    host_pinned = torch.empty((npix, 4), dtype=torch.float64, pin_memory=True) if (world > 1 and rank == 0) else None

    def e2e_step():
        if world == 1:
            integ.render(opts, sample_range=my_range, flags=4)      # SG_RENDER_OVERWRITE_FILM: film of this step only
            return
        # N > 1: the whole job's film has to land in rank 0's host memory -- device render, the NCCL film reduce, one D2H on rank 0
        film.zero_()
        integ.render_device(opts, film.data_ptr(), sample_range=my_range, stream=stream)
        reduce_film(film, dst=0)
        if rank == 0:
            host_pinned.copy_(film, non_blocking=True)
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):                      # untimed warm-up of the host path too: the first sg_render allocates the
        e2e_step()                                            # pinned staging buffer and first-touches the caller's film pages (~35 ms once)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = paths_total / float(e2e_t[0]) / 1e6

    if rank == 0:
        # ---- roofline of the dominant kernel (closest-hit traversal), measured live with CUDA events
        film.zero_()
        integ.render_device(opts, film.data_ptr(), sample_range=my_range, stream=stream, flags=2)   # per-kernel events
        st_t = integ.stats.as_dict()
        film.zero_()
        integ.render_device(opts, film.data_ptr(), sample_range=my_range, stream=stream, flags=1)   # visit counters
        st_c = integ.stats.as_dict()
        n_closest = st_c["closest_hit_rays"]
        nodes_per_ray = st_c["closest_nodes"] / max(n_closest, 1); tris_per_ray = st_c["closest_tris"] / max(n_closest, 1)
        bytes_per_ray = 32.0 + 32.0 * nodes_per_ray + 48.0 * tris_per_ray + 16.0
        n_launch = max(st_t["closest_launches"], 1)
        avg_launch_ms = st_t["closest_ms"] / n_launch
        alg_bytes_per_launch = bytes_per_ray * n_closest / n_launch
        achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        peak, peak_src = peaks()
        # DRAM traffic + issue-slot figures of the same kernel from the committed `ncu --set full` capture (never measured live:
        # a number taken under a profiler is not a bench value); only attached when the capture was made on this workload
        traffic = issue = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_trace_full_ncu.json")))
            if args.workload == "mesh1m" and not args.spp and not args.res:
                k0 = cap["launches"][0]
                traffic = k0["dram_read_bytes"] + k0["dram_write_bytes"]
                issue = {"issue_active_pct": k0["issue_active_pct"], "lanes_per_inst": k0["lanes_per_inst"], "warp_inst": k0["warp_inst"],
                         "launch": "depth-0 closest-hit launch, 67.1 M rays", "source": "profiles/r01_trace_full_ncu.json (ncu --set full)"}
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_trace<closest-hit>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "traffic_unit": "DRAM bytes read+written by the depth-0 closest-hit launch (ncu); "
                    "its algorithmic bytes are bytes_per_ray x 67.1 M rays = %.3g: the scene is L2-resident" % (bytes_per_ray * 67108864.0),
                    "issue": issue, "peak_source": peak_src,
                    "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                    "rays_per_launch": n_closest / n_launch, "avg_launch_ms": avg_launch_ms, "launches_per_step": n_launch,
                    "kernel_share_of_step": st_t["closest_ms"] / max(st_t["render_ms"], 1e-9),
                    "shadow_share_of_step": st_t["shadow_ms"] / max(st_t["render_ms"], 1e-9),
                    "closest_mrays_per_s": n_closest / max(st_t["closest_ms"], 1e-9) / 1e3,
                    "shadow_mrays_per_s": st_t["shadow_rays"] / max(st_t["shadow_ms"], 1e-9) / 1e3}
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample_spp = max(1, spp // 2)            # ~10 s of CPU work on the 16-core box for C2
            r = cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
            cpu = {"value": r["paths"] / r["secs"] / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port",
                   "mrays_per_s": r["rays"] / r["secs"] / 1e6, "seconds": r["secs"],
                   "sample": f"all {res[0]}x{res[1]} pixels x {sample_spp} spp (of {spp}), tile-parallel oracle, {threads} threads"}
        line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "mrays_per_s": mrays,
                "config": {"workload": f"{args.workload}: {cfg['desc']}", "resolution": list(res), "spp_per_gpu": spp,
                           "triangles": sc.meta["n_triangles"], "bvh_nodes": sc.meta["n_nodes"],
                           "integrator": "path maxdepth=%d, independent sampler, uniform light sampler" % cfg["max_depth"],
                           "parallelism": f"replicated scene, sample-range split x{world}, 1 NCCL film reduce/step" if world > 1 else "single GPU",
                           "l2": "working set (path state %.2f GB + scene) exceeds the 126 MB L2; no explicit flush" % (min(npix * spp, integ.max_paths_in_flight or (1 << 26)) * 276 / 1e9),
                           "scene_build_s": build_s, "scene_upload_s": upload_s},
                "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": C.sizeof(__import__("shimmer_b200").ffi.SgRenderParams),
                        "d2h_bytes_per_step": npix * 32,
                        "note": ("sg_render: host film buffer, scene resident (uploaded once)" if world == 1 else
                                 "sg_render_device on every rank + NCCL film reduce + D2H of the reduced film into rank 0's pinned host buffer; scene resident"),
                        "value_incl_scene_upload": paths_total / (float(e2e_t[0]) + args.steps * upload_s) / 1e6,
                        "scene_upload_bytes": int(sc.meta.get("upload_bytes", 0))},
                "gpu_launches": launches_total, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
