#!/usr/bin/env python3
"""One render of a BASELINE workload through the C ABI (device-resident film): the command the profilers wrap.

  python tools/render_once.py [--workload composite] [--spp 8] [--reps 1] [--warm 1]

At a reduced --spp the wavefront still covers EVERY pixel (paths are pixel-major), so per-ray instruction counts and
traversal statistics are those of the full-spp job; only the number of batches shrinks."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="composite")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--flags", type=int, default=0)
    args = ap.parse_args()
    import torch
    from shimmer_b200 import Options, create_integrator, scenes
    cfg = scenes.CONFIGS[args.workload]
    spp = args.spp or cfg["spp"]
    sc = cfg["builder"](resolution=cfg["resolution"]).build()
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp})
    opts = Options(seed=0, pixel_samples=spp)
    film = torch.zeros((integ.width * integ.height, 4), dtype=torch.float64, device="cuda")
    for r in range(args.warm + args.reps):
        film.zero_()
        integ.render_device(opts, film.data_ptr(), flags=args.flags)
        torch.cuda.synchronize()
        st = integ.stats
        print("render %d: %.2f ms, %.1f Mpaths/s, closest %d shadow %d rays, %d launches" %
              (r, st.render_ms, integ.width * integ.height * spp / st.render_ms / 1e3, st.closest_hit_rays, st.shadow_rays, st.kernel_launches), flush=True)
    integ.close()


if __name__ == "__main__":
    main()
