#!/usr/bin/env python3
"""Where the e2e (sg_render, host film) time goes beyond the device-resident render: wall-clock of sg_render_device + sync,
of sg_render, and of the D2H / host-copy legs in isolation (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from shimmer_b200 import Options, create_integrator, scenes

cfg = scenes.CONFIGS["mesh1m"]
sc = cfg["builder"](resolution=cfg["resolution"]).build()
integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": cfg["spp"], "seed": 0})
opts = Options(seed=0, pixel_samples=cfg["spp"])
npix = integ.width * integ.height
film = torch.zeros((npix, 4), dtype=torch.float64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
def dev():
    film.zero_(); integ.render_device(opts, film.data_ptr(), stream=stream); torch.cuda.synchronize()
def host(flags):
    integ.render(opts, flags=flags)
for name, fn in (("render_device+sync", dev), ("sg_render overwrite", lambda: host(4)), ("sg_render accumulate", lambda: host(0))):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(4): fn()
    print("%-24s %.2f ms/step" % (name, (time.perf_counter() - t0) / 4 * 1e3), flush=True)
pinned = torch.empty((npix, 4), dtype=torch.float64, pin_memory=True)
pageable = np.zeros((npix, 4))
for _ in range(3): pinned.copy_(film); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): pinned.copy_(film, non_blocking=True); torch.cuda.synchronize()
print("D2H pinned 33.5 MB        %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
src = pinned.numpy()
t0 = time.perf_counter()
for _ in range(10): np.copyto(pageable, src)
print("host memcpy 33.5 MB 1 thr %.2f ms" % ((time.perf_counter() - t0) / 10 * 1e3))
