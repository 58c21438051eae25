#!/usr/bin/env python3
"""Debug helper (GPU box): isolate which instanced shape / material combination makes GPU and oracle films differ."""
import sys, os, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from shimmer_b200 import Options, create_integrator, scenes
from shimmer_b200.host import SceneBuilder, Transform, named_spectrum

def build(shapes, mats, fix, top_sphere):
    b = SceneBuilder(); b.fix_instancing = fix
    b.set_camera(pos=(0.0, 1.4, -4.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=(32, 32))
    white = b.diffuse(scenes._white())
    if mats == "diffuse":
        mat = metal = glass = b.diffuse(scenes._green())
    else:
        mat = b.diffuse(scenes._green()); metal = b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.15)
        glass = b.dielectric(("const", 1.5))
    sph_xf = Transform.translate((0.0, 0.1, 0.0)) * Transform.rotate(-70.0, (1, 0, 0)); patch_xf = Transform.translate((0.0, 0.0, 0.5))
    tw = np.array([[-0.5, -0.3, -0.2], [0.5, -0.3, -0.4], [-0.4, 0.5, 0.3], [0.6, 0.4, -0.1]], np.float32)
    TP = np.array([[-0.4, -0.35, -0.3], [0.4, -0.35, -0.3], [0.0, -0.35, 0.4], [0.0, 0.2, 0.0]], np.float32)
    TI = np.array([[0, 1, 3], [1, 2, 3], [2, 0, 3]], np.uint32)
    xfs = [Transform.translate((-1.2, 0.6, 0.3)) * Transform.rotate(35.0, (0, 1, 0.3)) * Transform.scale(1.0, 1.4, 0.8),
           Transform.translate((0.0, 0.7, 0.0)) * Transform.scale(1.3, 1.3, 1.3),
           Transform.translate((1.2, 0.55, -0.2)) * Transform.rotate(-50.0, (1, 0.2, 0))]
    group = b.begin_object()
    if "s" in shapes: b.add_sphere(0.35, mat, z_min=-0.2, z_max=0.3, phi_max=300.0, object=group, object_from_world=sph_xf)
    if "p" in shapes: b.add_bilinear_mesh(tw, [[0, 1, 2, 3]], metal, object=group, object_from_world=patch_xf)
    if "t" in shapes: b.add_mesh(TP, TI, glass, object=group)
    for xf in xfs: b.add_instance(group, xf)
    if top_sphere: b.add_sphere(0.25, glass, object_from_world=Transform.translate((0.0, 0.25, -1.2)))
    gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3)); b.add_mesh(gp, gi, white)
    lp, li = scenes._quad((-0.6, 2.8, -0.6), (0.6, 2.8, -0.6), (0.6, 2.8, 0.6), (-0.6, 2.8, 0.6))
    b.add_mesh(lp, li, white, area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b.build()

for shapes, mats, fix, top in [("s", "diffuse", True, False), ("p", "diffuse", True, False), ("t", "diffuse", True, False), ("t", "orig", True, False),
                               ("sp", "diffuse", True, False), ("st", "diffuse", True, False), ("spt", "diffuse", True, False), ("spt", "orig", True, False),
                               ("t", "orig", False, False), ("spt", "diffuse", False, False)]:
    sc = build(shapes, mats, fix, top)
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 16, "seed": 5})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=5, spp=16))
    lg, lr = film[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
    rel = np.abs(lg - lr) / np.maximum(lr, 0.05 * lr.mean())
    print(shapes, mats, "fix" if fix else "lit", "bad", int((rel > 2e-3).sum()), "rays", integ.stats.closest_hit_rays, rst.closest_hit_rays, integ.stats.shadow_rays, rst.shadow_rays)
    integ.close()
