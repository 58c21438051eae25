#!/usr/bin/env python3
"""Upper bound of what sorting the ray queue could buy the closest-hit traversal kernel.

Camera rays of a BASELINE workload are traced through `sg_trace_device`; every hit spawns a cosine-distributed bounce ray
(what a diffuse surface would do), which gives the depth-1 ray set in PATH ORDER -- the order the wavefront's ray queue has.
The same set is then traced in other orders (the sort itself is done by torch and is NOT timed):

  path      the order the wavefront produces (pixel-major tiles, compacted)
  octant    global sort by direction octant, then 30-bit Morton code of the origin
  morton    global sort by the Morton code of the origin only
  chunk     octant-bucketed inside chunks of 1024 consecutive queue entries (keeps path-state locality)
  shuffle   random permutation (how bad can it get)

and once more for the depth-2 set.  Time = CUDA events around the kernel, best of --reps.

  python tools/exp_ray_order.py [--workload mesh1m] [--spp 16]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mesh1m")
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import numpy as np
    import torch
    from shimmer_b200 import Options, create_integrator, ffi, scenes
    cfg = scenes.CONFIGS[args.workload]
    res = cfg["resolution"]
    if res[0] * res[1] > 1 << 21:                       # C5: a 1920x1080 quarter frame is plenty
        res = (res[0] // 2, res[1] // 2)
    sc = cfg["builder"](resolution=res).build()
    integ = create_integrator("wavefront", {"maxdepth": cfg["max_depth"]}, sc, {"pixelsamples": args.spp})
    lib, handle = integ._lib, integ._handle
    W, H = integ.width, integ.height
    dev = torch.device("cuda")

    # camera rays, pixel-major in 8x4 tiles, samples of a pixel consecutive (the wavefront's depth-0 layout)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.int32), np.arange(W, dtype=np.int32), indexing="ij")
    tile = ((ys // 4) * ((W + 7) // 8) + xs // 8).ravel()
    inner = ((ys % 4) * 8 + xs % 8).ravel()
    order = np.lexsort((inner, tile))
    px = np.stack([xs.ravel()[order], ys.ravel()[order]], 1)
    pix = np.repeat(px, args.spp, axis=0)
    si = np.tile(np.arange(args.spp, dtype=np.int32), W * H)
    rays, _ = integ.camera_rays(Options(seed=0, pixel_samples=args.spp), pix, si)
    o = torch.from_numpy(rays[:, 0:3].copy()).to(dev); d = torch.from_numpy(rays[:, 3:6].copy()).to(dev)
    del rays

    hit_dt = np.dtype(ffi.SgHit)
    assert hit_dt.itemsize == 32

    def trace(o, d, reps):
        n = o.shape[0]
        o = o.contiguous(); d = d.contiguous()
        tmax = torch.full((n,), float("inf"), device=dev)
        out = torch.empty((n, 8), dtype=torch.float32, device=dev)
        best = 1e30
        for _ in range(reps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            ffi.check(lib.sg_trace_device(handle, n, C.c_void_p(o.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(tmax.data_ptr()), 0,
                                          C.c_void_p(out.data_ptr()), None, None), "sg_trace_device")
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return out, best

    def bounce(o, d, out, gen):
        prim = out[:, 0].view(torch.int32)
        m = prim >= 0
        t = out[m, 1]; ng = out[m, 5:8]
        o = o[m]; d = d[m]
        ng = torch.nn.functional.normalize(ng, dim=1)
        ng = torch.where((ng * d).sum(1, keepdim=True) > 0, -ng, ng)
        p = o + t[:, None] * d
        u = torch.rand((p.shape[0], 2), device=dev, generator=gen)
        r = u[:, 0].sqrt(); phi = 2 * np.pi * u[:, 1]
        lx, ly, lz = r * phi.cos(), r * phi.sin(), (1 - u[:, 0]).clamp_min(0).sqrt()
        a = torch.where(ng[:, 0:1].abs() > 0.9, torch.tensor([0.0, 1.0, 0.0], device=dev), torch.tensor([1.0, 0.0, 0.0], device=dev)).expand_as(ng)
        tx = torch.nn.functional.normalize(torch.linalg.cross(a, ng), dim=1)
        ty = torch.linalg.cross(ng, tx)
        nd = lx[:, None] * tx + ly[:, None] * ty + lz[:, None] * ng
        scale = p.abs().amax().item()
        return p + ng * (1e-4 * scale), torch.nn.functional.normalize(nd, dim=1)

    def morton30(p):
        lo = p.amin(0); hi = p.amax(0)
        q = ((p - lo) / (hi - lo).clamp_min(1e-20) * 1023.0).clamp(0, 1023).to(torch.int64)

        def spread(v):
            v = (v | (v << 16)) & 0x030000FF
            v = (v | (v << 8)) & 0x0300F00F
            v = (v | (v << 4)) & 0x030C30C3
            v = (v | (v << 2)) & 0x09249249
            return v
        return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

    def orders(o, d, gen):
        n = o.shape[0]
        octant = ((d[:, 0] < 0).to(torch.int64) | ((d[:, 1] < 0).to(torch.int64) << 1) | ((d[:, 2] < 0).to(torch.int64) << 2))
        mc = morton30(o)
        idx = torch.arange(n, device=dev, dtype=torch.int64)
        yield "path", idx
        yield "octant", torch.argsort((octant << 30) | mc, stable=True)
        yield "morton", torch.argsort(mc, stable=True)
        yield "chunk", torch.argsort(((idx >> 10) << 3) | octant, stable=True)
        yield "chunk16k", torch.argsort(((idx >> 14) << 3) | octant, stable=True)
        yield "shuffle", torch.randperm(n, device=dev, generator=gen)

    gen = torch.Generator(device=dev); gen.manual_seed(1)
    out0, ms0 = trace(o, d, args.reps)
    print("%s %dx%d x %d spp: depth 0 %d rays %.3f ms %.1f Mrays/s" % (args.workload, W, H, args.spp, o.shape[0], ms0, o.shape[0] / ms0 / 1e3), flush=True)
    for depth in (1, 2, 3):
        o, d = bounce(o, d, out0, gen)
        base = None
        for name, perm in orders(o, d, gen):
            oo, dd = o[perm], d[perm]
            out, ms = trace(oo, dd, args.reps)
            if name == "path":
                base = ms; out0 = out
            print("depth %d %-9s %9d rays %8.3f ms %8.1f Mrays/s  x%.3f vs path" % (depth, name, o.shape[0], ms, o.shape[0] / ms / 1e3, base / ms), flush=True)
            del oo, dd
    integ.close()


if __name__ == "__main__":
    main()
