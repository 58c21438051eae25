#!/usr/bin/env python3
"""A/B timing of the wavefront loop under different runtime knobs (run on the GPU box).

  python tools/perf_ab.py [--workload mesh1m] [--spp 64] CONFIG [CONFIG ...]

CONFIG is a comma-separated list of ENV=VALUE pairs ("base" = no overrides), e.g.
  base SG_PATH_ORDER=0 SG_LEAF_THRESHOLD=12,SG_REFILL_THRESHOLD=8
Knobs are read by sg_scene_create / sg_render_device, so every config re-creates the scene.
The pseudo-knob PIF=<n> sets max_paths_in_flight (the wavefront width).
Prints per-config: render ms, closest/shadow ms per depth (CUDA events), Mpaths/s.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mesh1m")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--res", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("configs", nargs="*", default=["base"])
    args = ap.parse_args()
    import numpy as np
    import torch
    from shimmer_b200 import Options, create_integrator, scenes
    cfg = scenes.CONFIGS[args.workload]
    res = (args.res, args.res) if args.res else cfg["resolution"]
    spp = args.spp or cfg["spp"]
    sc = cfg["builder"](resolution=res).build()
    os.environ["SG_DEBUG_TIMING"] = "1"
    ref_film = None
    for conf in args.configs:
        saved = {}
        pif = 0
        if conf != "base":
            for kv in conf.split(","):
                k, v = kv.split("=")
                if k == "PIF":                       # max_paths_in_flight (a constructor argument, not an environment knob)
                    pif = int(v)
                    continue
                saved[k] = os.environ.get(k)
                os.environ[k] = v
        integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp}, max_paths_in_flight=pif)
        opts = Options(seed=0, pixel_samples=spp)
        film = torch.zeros((integ.width * integ.height, 4), dtype=torch.float64, device="cuda")
        stream = torch.cuda.current_stream().cuda_stream
        best = None
        for r in range(args.reps + 1):
            film.zero_()
            integ.render_device(opts, film.data_ptr(), stream=stream, flags=0)
            torch.cuda.synchronize()
            ms = integ.stats.render_ms
            if r > 0:
                best = ms if best is None else min(best, ms)
        sys.stderr.write("[perf_ab] %s\n" % conf); sys.stderr.flush()
        film.zero_()
        integ.render_device(opts, film.data_ptr(), stream=stream, flags=2)       # per-launch events (prints via SG_DEBUG_TIMING)
        torch.cuda.synchronize()
        st = integ.stats.as_dict()
        f = film.cpu().numpy()
        if ref_film is None:
            ref_film = f
        lum_err = abs(f[:, :3].sum() - ref_film[:, :3].sum()) / max(ref_film[:, :3].sum(), 1e-30)
        print("%-60s best %.2f ms  %.1f Mpaths/s | timed run: total %.2f closest %.2f shadow %.2f other %.2f | energy diff vs first %.2e"
              % (conf, best, integ.width * integ.height * spp / best / 1e3, st["render_ms"], st["closest_ms"], st["shadow_ms"],
                 st["render_ms"] - st["closest_ms"] - st["shadow_ms"], lum_err), flush=True)
        integ.close()
        for k, v in saved.items():
            if v is None:
                del os.environ[k]
            else:
                os.environ[k] = v


if __name__ == "__main__":
    main()
