#!/usr/bin/env python3
"""Debug helper (GPU box): find the (pixel, sample) pairs whose radiance differs and print the oracle's hit sequence for them."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes, ffi
kind = sys.argv[1]; res = int(sys.argv[2]) if len(sys.argv) > 2 else 32; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sc = scenes.tiny_scene(kind, resolution=(res, res)).build()
integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp, "seed": 5})
prims = sc.arrays["prims"]; mats = sc.arrays["materials"]
def desc(pi):
    if pi < 0: return "miss"
    m = int(prims["mesh"][pi]); where = "obj" if pi >= sc.desc.n_top_primitives else "top"
    shape = "sphere" if m == ffi.SG_PRIM_SPHERE else ("patch" if sc.arrays["meshes"][m].flags & 32 else "tri")
    return "%s:%s:mat%d(kind %d)" % (where, shape, prims["material"][pi], mats[int(prims["material"][pi])].kind)
n = 0
for s in range(spp):
    integ.film[:] = 0
    g = integ.render(Options(), sample_range=(s, s + 1)).copy()
    r, _, _ = orc.render(sc, orc.make_params(seed=5, spp=spp, sample_range=(s, s + 1)))
    lg, lr = g[:, :3].sum(axis=1), r[:, :3].sum(axis=1)
    rel = np.abs(lg - lr) / np.maximum(lr, 1e-3)
    for i in np.nonzero(rel > 1e-3)[0]:
        px, py = int(i % res), int(i // res)
        rays = orc.path_rays(sc, orc.make_params(seed=5, spp=spp), px, py, s)
        seq = [("S" if q[7] else "C") + ":" + (desc(int(q[8])) if not q[7] else ("occ" if q[8] >= 0 else "free")) for q in rays]
        print("px", px, py, "s", s, "gpu", lg[i], "orc", lr[i], " | ".join(seq))
        n += 1
        if n >= 14: sys.exit(0)
