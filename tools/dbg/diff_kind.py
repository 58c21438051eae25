#!/usr/bin/env python3
"""Debug helper (GPU box): per-pixel film differences GPU vs oracle for one tiny-scene kind.  usage: diff_kind.py KIND [res] [spp]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes
kind = sys.argv[1]; res = int(sys.argv[2]) if len(sys.argv) > 2 else 16; spp = int(sys.argv[3]) if len(sys.argv) > 3 else 4; md = int(sys.argv[4]) if len(sys.argv) > 4 else 5
sc = scenes.tiny_scene(kind, resolution=(res, res)).build()
integ = create_integrator("wavefront", {"maxdepth": md}, sc, {"pixelsamples": spp, "seed": 5})
film = integ.render(Options()).copy()
ref, rst, _ = orc.render(sc, orc.make_params(seed=5, spp=spp, max_depth=md))
lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
rel = np.abs(lg - lr) / np.maximum(lr, 0.05 * lr.mean())
bad = np.nonzero(rel > 2e-3)[0]
print(kind, "pixels", len(lr), "bad", len(bad), "frac ok", 1 - len(bad) / len(lr), "rays", integ.stats.closest_hit_rays, rst.closest_hit_rays, integ.stats.shadow_rays, rst.shadow_rays)
print("rel quantiles", np.quantile(rel, [0.5, 0.9, 0.99, 0.999, 1.0]))
for i in bad[:12]:
    print(" px", i % res, i // res, "gpu", film[i, :3], "ref", ref[i, :3], "rel", rel[i])
