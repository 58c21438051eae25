#!/usr/bin/env python3
"""bench.py -- Mpaths/s (and Mrays/s) of the wavefront path tracer on BASELINE.json's multi-GPU workload.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload composite|mesh1m|cornell|glass|instanced]

Workload (every N): C5, BASELINE.json configs[4] -- the 4K (3840x2160) 1024 spp Cornell + 1M-triangle mesh composite -- the
configuration the metric "Mrays/s & Mpaths/s at 1/2/4/8 B200" is quoted on.  One "step" = one full pass of the hot path
(the `ImageTileIntegrator::render` equivalent, integrator.rs:227-321) over the whole frame: 3840 x 2160 pixels x 1024 samples.
STRONG scaling: the job is the same at every N; with N > 1 (one process per GPU, torchrun) the library splits the 1024
sample indices across the ranks (1024 / N each, SG_RENDER_SPLIT_SAMPLES) and sums the f64 films onto rank 0 with ONE
in-library ncclReduce per step (SG_RENDER_REDUCE_FILM) -- both behind the C ABI (include/shimmer_gpu.h).

At N = 1 the line also carries `c2`: the same measurement on configs[1] (1M-triangle mesh, 1024^2 x 64 spp), the traversal-bound
single-GPU case round 1's headline was quoted on.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how roofline / cpu_baseline are defined.
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

ISSUE_CAPTURE = os.path.join(ROOT, "profiles", "r02_issue.json")      # written by tools/make_issue_capture.py from an ncu run


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="composite", choices=["mesh1m", "cornell", "glass", "instanced", "composite"])
    ap.add_argument("--spp", type=int, default=0, help="override the TOTAL samples per pixel (default: the config's)")
    ap.add_argument("--res", type=int, default=0, help="override square resolution (debug only)")
    ap.add_argument("--paths-in-flight", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2", action="store_true", help="skip the C2 side measurement of the N = 1 line")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer (e2e) loop; default = --steps")
    return ap.parse_args()


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def build_scene(workload, res_override=0):
    from shimmer_b200 import scenes
    cfg = scenes.CONFIGS[workload]
    res = (res_override, res_override) if res_override else cfg["resolution"]
    t0 = time.time()
    sc = cfg["builder"](resolution=res).build()
    return sc, cfg, res, time.time() - t0


def workload_config(workload, cfg, res, spp):
    """The `config` object: identical in both arms (--impl ours / reference)."""
    return {"workload": f"{workload}: {cfg['desc']}", "resolution": list(res), "spp": spp,
            "integrator": "path maxdepth=%d, independent sampler, uniform light sampler" % cfg["max_depth"]}


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
        # median over the samples taken under load (the idle gaps between steps would drag it down)
        busy = [s for s in sm if mx and s >= 0.5 * max(mx)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(sc, spp_sample, max_depth, threads, seed=0):
    """Times the CPU oracle (the restated reference: the Rust binary cannot be built in this image) in
    tile-parallel mode on `threads` host threads.  The ONLY place bench.py executes oracle/ code."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    p = orc.make_params(seed=seed, spp=spp_sample, max_depth=max_depth)
    film, st, secs = orc.render(sc, p, n_threads=threads, stream_mode=1)      # stream_mode 1 = reference behaviour
    return dict(secs=secs, paths=int(st.camera_paths), rays=int(st.closest_hit_rays + st.shadow_rays))


def cpu_sample_spp(npix, spp, budget_paths):
    """Bounded CPU sample: all pixels x as many sample indices as fit `budget_paths` camera paths."""
    return int(max(1, min(spp, budget_paths // max(npix, 1))))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sc, cfg, res, build_s = build_scene(args.workload, args.res)
    spp = args.spp or cfg["spp"]
    threads = os.cpu_count() or 1
    npix = res[0] * res[1]
    # bounded sample of the same workload per step: all pixels x a few sample indices (~4 s of a 16-core host)
    sample_spp = cpu_sample_spp(npix, spp, 17_000_000)
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
    t_paths = t_rays = 0; t_secs = 0.0
    for _ in range(args.steps):
        r = cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
        t_paths += r["paths"]; t_rays += r["rays"]; t_secs += r["secs"]
    v = t_paths / t_secs / 1e6
    line = {"impl": "reference", "metric": "Mpaths/s", "value": v, "unit": "Mpaths/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_secs / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mrays_per_s": t_rays / t_secs / 1e6,
            "config": workload_config(args.workload, cfg, res, spp),
            "cpu_baseline": {"value": v, "unit": "Mpaths/s", "cores": threads, "kind": "port",
                             "sample": f"all {res[0]}x{res[1]} pixels x {sample_spp} spp per step (of {spp}); tile-parallel 8x8 tiles, "
                                       f"{threads} threads, reference RNG mode; oracle = C++ restatement (reference is Rust, not buildable here)"},
            "e2e": {"value": v, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def issue_capture(workload):
    """Instruction counts of the traversal kernel from the committed ncu capture (profiles/r02_issue.json: warp- and
    thread-level instructions per closest-hit ray, DRAM traffic of the depth-0 launch, the commit it was taken at)."""
    try:
        cap = json.load(open(ISSUE_CAPTURE))
        return cap.get(workload)
    except Exception:
        return None


def traversal_roofline(integ, opts, film, stream, sm_clock_mhz, num_sms, workload, sample_range=None):
    """Roofline of the dominant kernel (closest-hit traversal), measured live with CUDA events around every launch.
    `sample_range`: this rank's share of the job (no split / reduce flags: the other ranks are idle meanwhile)."""
    from shimmer_b200 import ffi
    film.zero_()
    integ.render_device(opts, film.data_ptr(), sample_range=sample_range, stream=stream, flags=ffi.SG_RENDER_TIME_KERNELS)
    st_t = integ.stats.as_dict()
    film.zero_()
    integ.render_device(opts, film.data_ptr(), sample_range=sample_range, stream=stream, flags=ffi.SG_RENDER_COUNT_VISITS)
    st_c = integ.stats.as_dict()
    n_closest = st_c["closest_hit_rays"]
    nodes_per_ray = st_c["closest_nodes"] / max(n_closest, 1); tris_per_ray = st_c["closest_tris"] / max(n_closest, 1)
    bytes_per_ray = 32.0 + 32.0 * nodes_per_ray + 48.0 * tris_per_ray + 16.0
    n_launch = max(st_t["closest_launches"], 1)
    avg_launch_ms = st_t["closest_ms"] / n_launch
    closest_s = st_t["closest_ms"] * 1e-3
    hbm_achieved = bytes_per_ray * n_closest / closest_s / 1e9
    hbm_peak, peak_src = peaks()
    clock_hz = (sm_clock_mhz or 1965.0) * 1e6
    issue_peak = num_sms * 4 * clock_hz / 1e9                                          # G warp-instructions / s: 4 schedulers x 1 inst / clk
    cap = issue_capture(workload)
    roof = {"kernel": "k_trace<closest-hit>",
            "rays_per_launch": n_closest / n_launch, "avg_launch_ms": avg_launch_ms, "launches_per_step": n_launch,
            "bytes_per_ray": bytes_per_ray, "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
            "kernel_share_of_step": st_t["closest_ms"] / max(st_t["render_ms"], 1e-9),
            "shadow_share_of_step": st_t["shadow_ms"] / max(st_t["render_ms"], 1e-9),
            "closest_mrays_per_s": n_closest / max(st_t["closest_ms"], 1e-9) / 1e3,
            "shadow_mrays_per_s": st_t["shadow_rays"] / max(st_t["shadow_ms"], 1e-9) / 1e3,
            "hbm_equivalent": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak, "peak_source": peak_src,
                               "note": "SURVEY 8(d) algorithmic bytes per ray x rays / kernel time; the scene is L2-resident, so this is a work-rate "
                                       "figure expressed in bytes, not DRAM traffic (see `traffic`)"}}
    if cap:
        achieved = cap["warp_inst_per_closest_ray"] * n_closest / closest_s / 1e9
        roof.update({"bound": "issue", "achieved": achieved, "peak": issue_peak, "unit": "Gwarp-inst/s", "frac": achieved / issue_peak,
                     "peak_source": "%d SMs x 4 schedulers x %.0f MHz (SM clock sampled during the run)" % (num_sms, clock_hz / 1e6),
                     "warp_execution_efficiency": cap["thread_inst_per_closest_ray"] / cap["warp_inst_per_closest_ray"] / 32.0,
                     "lanes_per_inst": cap["thread_inst_per_closest_ray"] / cap["warp_inst_per_closest_ray"],
                     "warp_inst_per_ray": cap["warp_inst_per_closest_ray"],
                     "traffic": cap.get("dram_bytes_depth0_launch"),
                     "traffic_note": "DRAM bytes read+written by the depth-0 closest-hit launch of the capture (%s); its algorithmic bytes: %.3g"
                                     % (cap.get("capture_desc", ""), bytes_per_ray * cap.get("rays_depth0_launch", 0)),
                     "capture": {"file": "profiles/r02_issue.json", "commit": cap.get("commit"), "how": cap.get("how")}})
    else:       # no instruction capture for this workload: report the byte-equivalent figure as the headline fraction
        roof.update({"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                     "traffic": None, "peak_source": peak_src})
    return roof


def timed_device_loop(step, steps, warmup, torch, dist, world):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return ev0.elapsed_time(ev1)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from shimmer_b200 import Options, create_integrator, ffi
    from shimmer_b200 import distributed as sgd

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: the GPU path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # the data-path collective lives inside the library: torch.distributed only carries the NCCL unique id, barriers and
        # the max-over-ranks of the timings
        sgd.init_process_comm(rank, world, sgd.torch_broadcast_bytes(torch.device("cuda", local)), device=local)
    sc, cfg, res, build_s = build_scene(args.workload, args.res)
    spp = args.spp or cfg["spp"]
    warmup = max(args.warmup, 3)
    t0 = time.time()
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp}, device=local,
                              max_paths_in_flight=args.paths_in_flight)
    upload_s = time.time() - t0
    opts = Options(seed=0, pixel_samples=spp)
    W, H = integ.width, integ.height
    npix = W * H
    film = torch.zeros((npix, 4), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream          # 0 on torch's default stream = the legacy default stream for the library too
    MG = (ffi.SG_RENDER_SPLIT_SAMPLES | ffi.SG_RENDER_REDUCE_FILM) if world > 1 else 0
    acc = dict(launches=0, rays=0, reduce_ms=0.0, render_ms=0.0, n=0)

    def step():
        film.zero_()
        integ.render_device(opts, film.data_ptr(), stream=stream, flags=MG)     # this rank's 1/N of the samples + the in-library film reduce
        st = integ.stats
        acc["launches"] += int(st.kernel_launches); acc["rays"] += int(st.closest_hit_rays + st.shadow_rays)
        acc["reduce_ms"] += float(st.reduce_ms); acc["render_ms"] += float(st.render_ms); acc["n"] += 1

    for _ in range(warmup):
        step()
    for k in acc:
        acc[k] = 0
    clocks = ClockSampler(local); clocks.start()
    ms = timed_device_loop(step, args.steps, 0, torch, dist, world)
    clk = clocks.stop()
    t = torch.tensor([ms, float(acc["rays"]), float(acc["launches"]), acc["reduce_ms"], acc["render_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        ms = float(tmax[0]); rays_total = float(tsum[1]); launches_total = int(tsum[2])
        render_ms_max = float(tmax[4]) / args.steps; reduce_ms_min = float(tmin[3]) / args.steps; reduce_ms_root = float(t[3]) / args.steps
    else:
        rays_total = float(t[1]); launches_total = int(t[2])
        render_ms_max = acc["render_ms"] / args.steps; reduce_ms_min = reduce_ms_root = 0.0
    paths_total = float(npix) * spp * args.steps             # strong scaling: the whole job is npix x spp paths at every N
    value = paths_total / (ms * 1e-3) / 1e6
    mrays = rays_total / (ms * 1e-3) / 1e6

    # ---- e2e: the user-facing call with HOST buffers (sg_render: params in; split, reduce and the film D2H inside the call)
    e2e_steps = args.e2e_steps or args.steps
    e2e_acc = dict(reduce_ms=0.0, d2h_ms=0.0)

    def e2e_step():
        integ.render(opts, flags=ffi.SG_RENDER_OVERWRITE_FILM | MG)      # rank 0's host film receives the whole job's film
        e2e_acc["reduce_ms"] += float(integ.stats.reduce_ms); e2e_acc["d2h_ms"] += float(integ.stats.d2h_ms)

    for _ in range(warmup):                                   # untimed warm-up of the host path too: the first sg_render allocates the
        e2e_step()                                            # pinned staging buffer and first-touches the caller's film pages
    e2e_acc = dict(reduce_ms=0.0, d2h_ms=0.0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = float(npix) * spp * e2e_steps / float(e2e_t[0]) / 1e6

    if rank == 0:
        num_sms = torch.cuda.get_device_properties(local).multi_processor_count
        # roofline of the dominant kernel on THIS rank's share of the job (no collective inside: the other ranks are idle)
        b, e = C.c_int32(), C.c_int32()
        ffi.check(ffi.load_library().sg_sample_range_for_rank(0, spp, 0, world, C.byref(b), C.byref(e)))
        roofline = traversal_roofline(integ, opts, film, stream, clk.get("sm_mhz"), num_sms, args.workload, sample_range=(b.value, e.value))
        cpu = None
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            sample_spp = cpu_sample_spp(npix, spp, 50_000_000)            # ~10-15 s of CPU work on a 16-core host
            r = cpu_reference_run(sc, sample_spp, cfg["max_depth"], threads)
            cpu = {"value": r["paths"] / r["secs"] / 1e6, "unit": "Mpaths/s", "cores": threads, "kind": "port",
                   "mrays_per_s": r["rays"] / r["secs"] / 1e6, "seconds": r["secs"],
                   "sample": f"all {res[0]}x{res[1]} pixels x {sample_spp} spp (of {spp}), tile-parallel oracle (C++ restatement of the Rust reference), {threads} threads"}
        c2 = None
        if world == 1 and args.workload == "composite" and not args.no_c2 and not args.spp and not args.res:
            integ.close()
            del film
            torch.cuda.empty_cache()
            c2 = c2_side_measurement(torch, dist)
        line = {"metric": "Mpaths/s", "value": value, "unit": "Mpaths/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "mrays_per_s": mrays,
                "config": workload_config(args.workload, cfg, res, spp),
                "setup": {"triangles": sc.meta["n_triangles"], "bvh_nodes": sc.meta["n_nodes"], "spp_per_gpu": spp / world,
                          "parallelism": (f"replicated scene, sample-range split x{world} and ONE ncclReduce of the f64 film per step, both inside the "
                                          "library (SG_RENDER_SPLIT_SAMPLES | SG_RENDER_REDUCE_FILM)") if world > 1 else "single GPU",
                          "l2": "working set (path state %.2f GB + scene) exceeds the 126 MB L2; no explicit flush"
                                % (min(npix * spp / world, integ.max_paths_in_flight or (1 << 26)) * 276 / 1e9),
                          "scene_build_s": build_s, "scene_upload_s": upload_s},
                "multi_gpu": {"render_ms_slowest_rank": render_ms_max, "reduce_ms_root": reduce_ms_root, "reduce_ms_fastest_rank": reduce_ms_min,
                              "film_bytes": npix * 32,
                              "note": "per step, device-timed: the slowest rank's wavefront loop; the ncclReduce as seen by rank 0 (includes waiting "
                                      "for the slowest rank) and by the rank that waited least (~ the collective itself)"},
                "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": C.sizeof(ffi.SgRenderParams),
                        "d2h_bytes_per_step": npix * 32, "steps": e2e_steps,
                        "reduce_ms_root": e2e_acc["reduce_ms"] / e2e_steps, "d2h_ms_root": e2e_acc["d2h_ms"] / e2e_steps,
                        "note": "sg_render on every rank: host parameters in, this rank's share of the samples, in-library NCCL film reduce, "
                                "D2H of the reduced film into rank 0's host film; scene resident (uploaded once)",
                        "value_incl_scene_upload": float(npix) * spp * e2e_steps / (float(e2e_t[0]) + e2e_steps * upload_s) / 1e6,
                        "scene_upload_bytes": int(sc.meta.get("upload_bytes", 0))},
                "gpu_launches": launches_total, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu}
        if c2 is not None:
            line["c2"] = c2
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        sgd.destroy_process_comm()
        dist.destroy_process_group()


def c2_side_measurement(torch, dist, steps=5):
    """configs[1] on one GPU, device-resident (round 1's headline measurement, kept for continuity)."""
    from shimmer_b200 import Options, create_integrator
    sc, cfg, res, _ = build_scene("mesh1m")
    spp = cfg["spp"]
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp})
    opts = Options(seed=0, pixel_samples=spp)
    npix = integ.width * integ.height
    film = torch.zeros((npix, 4), dtype=torch.float64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    rays = [0]

    def step():
        film.zero_()
        integ.render_device(opts, film.data_ptr(), stream=stream)
        rays[0] += int(integ.stats.closest_hit_rays + integ.stats.shadow_rays)
    for _ in range(3):
        step()
    rays[0] = 0
    clocks = ClockSampler(torch.cuda.current_device()); clocks.start()
    ms = timed_device_loop(step, steps, 0, torch, dist, 1)
    clk = clocks.stop()
    num_sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    roof = traversal_roofline(integ, opts, film, stream, clk.get("sm_mhz"), num_sms, "mesh1m")
    out = {"config": workload_config("mesh1m", cfg, res, spp), "value": npix * spp * steps / (ms * 1e-3) / 1e6, "unit": "Mpaths/s",
           "mrays_per_s": rays[0] / (ms * 1e-3) / 1e6, "ms_per_step": ms / steps, "steps": steps, "roofline": roof}
    integ.close()
    return out


if __name__ == "__main__":
    main()
