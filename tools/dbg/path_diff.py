#!/usr/bin/env python3
"""Debug helper (GPU box): for the pixels whose film differs, replay every ray the ORACLE's paths trace through sg_trace and report
the rays whose hit differs.  usage: path_diff.py KIND [res] [spp]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes, ffi
kind = sys.argv[1]; res = int(sys.argv[2]) if len(sys.argv) > 2 else 32; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sc = scenes.tiny_scene(kind, resolution=(res, res)).build()
integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp, "seed": 5})
film = integ.render(Options()).copy()
p = orc.make_params(seed=5, spp=spp)
ref, rst, _ = orc.render(sc, p)
lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
rel = np.abs(lg - lr) / np.maximum(lr, 0.05 * lr.mean())
bad = np.nonzero(rel > 2e-3)[0]
print(kind, "bad pixels", len(bad))
prims = sc.arrays["prims"]
nrep = 0
for i in bad[:8]:
    px, py = int(i % res), int(i // res)
    for s in range(spp):
        rays = orc.path_rays(sc, p, px, py, s)
        for k, r in enumerate(rays):
            o, d, tm, anyh, prim, t = r[0:3], r[3:6], r[6], r[7] != 0, int(r[8]), r[9]
            got = integ.trace(o[None], d[None], np.array([tm], np.float32), any_hit=bool(anyh))[0]
            gp = int(got["prim"]); exp = (0 if prim >= 0 else -1) if anyh else prim
            if gp != exp or (not anyh and prim >= 0 and got["t"] != t):
                def desc(pi):
                    if pi < 0: return "miss"
                    m = int(prims["mesh"][pi]); where = "obj" if pi >= sc.desc.n_top_primitives else "top"
                    return "%s:%s" % (where, "sphere" if m == ffi.SG_PRIM_SPHERE else ("inst" if m == ffi.SG_PRIM_INSTANCE else ("patch" if sc.arrays["meshes"][m].flags & 32 else "tri")))
                print(" px", px, py, "s", s, "ray", k, "any" if anyh else "closest", "tmax", tm, "oracle", prim, desc(prim) if not anyh else "", t, "gpu", gp, desc(gp) if not anyh else "", float(got["t"]), "o", o, "d", d)
                nrep += 1
                break
    if nrep > 12: break
print("done", nrep)
