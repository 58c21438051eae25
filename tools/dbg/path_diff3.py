#!/usr/bin/env python3
"""Debug helper (GPU box): for differing (pixel, sample) pairs find the smallest max_depth at which GPU and oracle disagree."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes, ffi
kind = sys.argv[1]; res = 32; spp = 16
sc = scenes.tiny_scene(kind, resolution=(res, res)).build()
prims = sc.arrays["prims"]; mats = sc.arrays["materials"]
def desc(pi):
    if pi < 0: return "miss"
    m = int(prims["mesh"][pi]); where = "obj" if pi >= sc.desc.n_top_primitives else "top"
    shape = "sphere" if m == ffi.SG_PRIM_SPHERE else ("patch" if sc.arrays["meshes"][m].flags & 32 else "tri")
    return "%s:%s:k%d" % (where, shape, mats[int(prims["material"][pi])].kind)
integs = {md: create_integrator("wavefront", {"maxdepth": md}, sc, {"pixelsamples": spp, "seed": 5}) for md in range(0, 6)}
n = 0
for s in range(spp):
    films = {}
    for md, integ in integs.items():
        integ.film[:] = 0
        g = integ.render(Options(), sample_range=(s, s + 1)).copy()
        r, _, _ = orc.render(sc, orc.make_params(seed=5, spp=spp, sample_range=(s, s + 1), max_depth=md))
        films[md] = (g[:, :3].sum(axis=1), r[:, :3].sum(axis=1))
    lg, lr = films[5]
    for i in np.nonzero(np.abs(lg - lr) / np.maximum(lr, 1e-2) > 1e-3)[0]:
        px, py = int(i % res), int(i // res)
        first = min(md for md in range(6) if abs(films[md][0][i] - films[md][1][i]) > 1e-3 * max(films[md][1][i], 1e-2))
        rays = orc.path_rays(sc, orc.make_params(seed=5, spp=spp), px, py, s)
        seq = [("S" if q[7] else "C") + ":" + (desc(int(q[8])) if not q[7] else ("occ" if q[8] >= 0 else "free")) for q in rays]
        print("px", px, py, "s", s, "first differing max_depth", first, [("%.4f/%.4f" % (films[md][0][i], films[md][1][i])) for md in range(6)], " | ".join(seq))
        n += 1
        if n >= 10: sys.exit(0)
