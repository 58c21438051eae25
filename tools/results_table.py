#!/usr/bin/env python3
"""Results table over BASELINE.json's five configs (run on the GPU box): throughput of the CUDA path at the
full configuration, the CPU oracle on a crop window of the same scene, and image agreement on that window.

  python tools/results_table.py [--configs cornell mesh1m glass instanced composite] [--crop 96]
                                [--out gpurun_out/results.json] [--md gpurun_out/results.md]

Per config it reports
  * GPU: Mpaths/s, Mrays/s of ONE full render (all pixels x spp; device-resident film), after one warm-up render;
    traversal roofline fraction (SURVEY 8d bytes/ray x closest-hit rays / closest-hit kernel time / HBM peak);
  * CPU: the oracle (tile-parallel, all host threads, reference RNG mode) on a `crop` x `crop` pixel window at full spp;
  * image agreement inside the window, developed RGB (film.rs:720-738):
      same-stream  : GPU vs oracle with the SAME (pixel, sample) random streams -> RMSE/mean and |dY|/Y (should be ~1e-6)
      independent  : GPU vs oracle in the reference's sequential-RNG mode (different random numbers, what comparing
                     against a real shimmer render would look like) -> noise-limited RMSE and the mean-luminance error.
This is a tool, not product code: it imports tests/orc.py (the oracle) as the checker.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def lum(rgb):
    return 0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]


def agreement(a, b):
    """relative RMSE and mean relative luminance error of developed image a against b."""
    a = a.astype(np.float64); b = b.astype(np.float64)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)) / max(np.mean(b), 1e-30))
    dl = float(abs(lum(a).mean() - lum(b).mean()) / max(lum(b).mean(), 1e-30))
    return rmse, dl


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=["cornell", "mesh1m", "glass", "instanced", "composite"])
    ap.add_argument("--crop", type=int, default=96)
    ap.add_argument("--spp-scale", type=float, default=1.0, help="debug: scale every config's spp")
    ap.add_argument("--out", default="gpurun_out/results.json")
    ap.add_argument("--md", default="gpurun_out/results.md")
    args = ap.parse_args()
    import torch
    import orc
    from shimmer_b200 import Options, create_integrator, scenes
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    threads = os.cpu_count() or 1
    rows = []
    for name in args.configs:
        cfg = scenes.CONFIGS[name]
        W, H = cfg["resolution"]; spp = max(1, int(cfg["spp"] * args.spp_scale))
        t0 = time.time(); sc = cfg["builder"](resolution=(W, H)).build(); build_s = time.time() - t0
        integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp})
        opts = Options(seed=0, pixel_samples=spp)
        film = torch.zeros((W * H, 4), dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        warm = min(spp, ((1 << 26) + W * H - 1) // (W * H) + 1)                                      # fills the wavefront: all buffers allocated
        integ.render_device(opts, film.data_ptr(), sample_range=(0, warm), stream=stream)             # warm-up
        torch.cuda.synchronize(); film.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); integ.render_device(opts, film.data_ptr(), stream=stream); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = integ.stats.as_dict()
        paths = W * H * spp; rays = st["closest_hit_rays"] + st["shadow_rays"]
        gfilm = film.cpu().numpy().copy()
        # roofline of the closest-hit traversal kernel: per-kernel events (flags=2) + visit counters (flags=1) on <=16 spp
        sub = (0, min(spp, 16))
        scratch = torch.zeros_like(film)
        integ.render_device(opts, scratch.data_ptr(), sample_range=sub, stream=stream, flags=2); st_t = integ.stats.as_dict()
        scratch.zero_()
        integ.render_device(opts, scratch.data_ptr(), sample_range=sub, stream=stream, flags=1); st_c = integ.stats.as_dict()
        nc = max(st_c["closest_hit_rays"], 1)
        npr, tpr = st_c["closest_nodes"] / nc, st_c["closest_tris"] / nc
        bpr = 32.0 + 32.0 * npr + 48.0 * tpr + 16.0
        frac = bpr * st_t["closest_hit_rays"] / (st_t["closest_ms"] * 1e-3) / 1e9 / peak
        del scratch
        gdev = integ.develop(gfilm)
        integ.close(); del film
        # ---- oracle on a crop window ----
        c = min(args.crop, W, H)
        x0 = (W - c) // 2; y0 = min(H - c, int(H * 0.55))          # below centre: objects + floor + shadows in every scene
        win = (x0, y0, x0 + c, y0 + c)
        scc = cfg["builder"](resolution=(W, H), crop=win).build()
        p = orc.make_params(seed=0, spp=spp, max_depth=cfg["max_depth"])
        ofilm, _, _ = orc.render(scc, p, n_threads=threads, stream_mode=0)
        rfilm, rst, rsecs = orc.render(scc, p, n_threads=threads, stream_mode=1)
        gwin = gfilm.reshape(H, W, 4)[y0:y0 + c, x0:x0 + c].reshape(-1, 4)
        d_g = gdev[y0:y0 + c, x0:x0 + c].reshape(-1, 3)
        d_o, d_r = orc.develop(scc, ofilm), orc.develop(scc, rfilm)
        same = agreement(d_g, d_o); indep = agreement(d_g, d_r)
        energy = float(abs(gwin[:, :3].sum() - ofilm[:, :3].sum()) / max(ofilm[:, :3].sum(), 1e-30))
        row = dict(config=name, desc=cfg["desc"], resolution=[W, H], spp=spp, triangles=sc.meta.get("n_triangles"),
                   instanced_triangles=sc.meta.get("n_instanced_triangles"), lights=sc.meta.get("n_lights"), scene_build_s=build_s,
                   gpu_ms=ms, gpu_mpaths=paths / ms / 1e3, gpu_mrays=rays / ms / 1e3, launches=st["kernel_launches"],
                   nodes_per_ray=npr, tris_per_ray=tpr, bytes_per_ray=bpr, trav_frac=frac,
                   closest_mrays=st_t["closest_hit_rays"] / max(st_t["closest_ms"], 1e-9) / 1e3,
                   cpu_threads=threads, cpu_window=list(win), cpu_secs=rsecs, cpu_mpaths=rst.camera_paths / rsecs / 1e6,
                   cpu_mrays=(rst.closest_hit_rays + rst.shadow_rays) / rsecs / 1e6,
                   rmse_same=same[0], dlum_same=same[1], rmse_indep=indep[0], dlum_indep=indep[1], film_energy_rel=energy)
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)
    with open(args.md, "w") as f:
        f.write("| config | GPUs | Mrays/s | Mpaths/s | ms/render | CPU threads | CPU Mrays/s | CPU Mpaths/s (crop window) | "
                "RMSE same-stream | rel-lum err same-stream | RMSE vs reference-RNG oracle | rel-lum err vs reference-RNG oracle | "
                "traversal roofline frac (HBM, SURVEY 8d bytes) |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            f.write("| %s %dx%d %d spp | 1 | %.0f | %.1f | %.1f | %d | %.2f | %.2f | %.2e | %.2e | %.3f | %.4f | %.3f |\n" % (
                r["config"], r["resolution"][0], r["resolution"][1], r["spp"], r["gpu_mrays"], r["gpu_mpaths"], r["gpu_ms"],
                r["cpu_threads"], r["cpu_mrays"], r["cpu_mpaths"], r["rmse_same"], r["dlum_same"], r["rmse_indep"], r["dlum_indep"],
                r["trav_frac"]))


if __name__ == "__main__":
    main()
