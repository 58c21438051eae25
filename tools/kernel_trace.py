#!/usr/bin/env python3
"""Per-kernel device time of ONE render in a normal (unprofiled-clock, unserialised-cache) run: torch.profiler's CUPTI activity trace
sees the kernels libshimmer_gpu.so launches in this process.  Complements the ncu launch lists, whose per-launch times are
cold-cache and serialised.

  python tools/kernel_trace.py [--workload composite] [--spp 16] [--overlap 1]"""
import argparse
import collections
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="composite")
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--overlap", default="1")
    args = ap.parse_args()
    os.environ["SG_OVERLAP"] = args.overlap
    import torch
    from torch.profiler import ProfilerActivity, profile
    from shimmer_b200 import Options, create_integrator, scenes
    cfg = scenes.CONFIGS[args.workload]
    spp = args.spp or cfg["spp"]
    sc = cfg["builder"](resolution=cfg["resolution"]).build()
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": spp})
    opts = Options(seed=0, pixel_samples=spp)
    film = torch.zeros((integ.width * integ.height, 4), dtype=torch.float64, device="cuda")
    integ.render_device(opts, film.data_ptr()); torch.cuda.synchronize()            # warm-up
    film.zero_()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        integ.render_device(opts, film.data_ptr()); torch.cuda.synchronize()
    ms = integ.stats.render_ms
    agg = collections.OrderedDict(); tot = 0.0
    for e in prof.events():
        if e.device_type is None or "DeviceType.CUDA" not in str(e.device_type):
            continue
        name = re.sub(r"\(.*", "", e.name).replace("void ", "").replace("sg::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += e.device_time / 1e3 if hasattr(e, "device_time") else e.cuda_time / 1e3
        tot += a[1] * 0
    tot = sum(v[1] for v in agg.values())
    print("%s %d spp, SG_OVERLAP=%s: render %.2f ms (CUDA events), kernel time summed %.2f ms" % (args.workload, spp, args.overlap, ms, tot))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print("%-60s x%4d %9.3f ms %5.1f %%" % (n[:60], c, t, 100 * t / tot))
    integ.close()


if __name__ == "__main__":
    main()
