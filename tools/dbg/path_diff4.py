#!/usr/bin/env python3
"""Debug helper (GPU box): single-pixel, single-sample renders at increasing max_depth: radiance + ray counts GPU vs oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes
kind = sys.argv[1]; cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
for px, py, s in cases:
    sc = scenes.tiny_scene(kind, resolution=(32, 32)).build()
    sc.desc.film.pixel_bounds[:] = [px, py, px + 1, py + 1]
    print("pixel", px, py, "sample", s)
    for md in range(0, 6):
        integ = create_integrator("wavefront", {"maxdepth": md}, sc, {"pixelsamples": 16, "seed": 5})
        g = integ.render(Options(), sample_range=(s, s + 1)).copy()
        p = orc.make_params(seed=5, spp=16, sample_range=(s, s + 1), max_depth=md)
        r, rst, _ = orc.render(sc, p)
        rays = orc.path_rays(sc, p, px, py, s)
        print("  md", md, "gpu L %.5f rays %d/%d | orc L %.5f rays %d/%d" % (g[0, :3].sum(), integ.stats.closest_hit_rays, integ.stats.shadow_rays,
                                                                                  r[0, :3].sum(), rst.closest_hit_rays, rst.shadow_rays),
              " ".join(("S" if q[7] else "C") + ("%d" % int(q[8])) for q in rays))
        integ.close()
