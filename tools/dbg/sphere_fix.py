#!/usr/bin/env python3
"""Debug helper (GPU box): differing-pixel counts of the sphere scenes in the literal and the fixed transform mode."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes
for kind in ("spheres", "spherestex", "spherelight"):
    for fix in (False, True):
        sc = scenes.sphere_tiny_scene(kind, resolution=(32, 32), fix=fix).build()
        integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 16, "seed": 5})
        film = integ.render(Options()).copy()
        ref, rst, _ = orc.render(sc, orc.make_params(seed=5, spp=16))
        lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
        rel = np.abs(lg - lr) / np.maximum(lr, 0.05 * lr.mean())
        print(kind, "fix" if fix else "literal", "bad", int((rel > 2e-3).sum()), "rays", integ.stats.closest_hit_rays, rst.closest_hit_rays, integ.stats.shadow_rays, rst.shadow_rays)
        integ.close()
