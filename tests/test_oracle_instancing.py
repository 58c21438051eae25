"""Oracle checks for object instancing (primitive.rs:136-176, transform.rs:573-609,701-723).
The reference has no unit test for TransformedPrimitive, so the pins are structural: with SG_SCENE_FIX_INSTANCING
(pbrt semantics) an instanced scene must reproduce the same scene with the instances baked into meshes, and the
reference-literal mode must differ from it exactly where the reference's quirks say it does."""
import json
import os

import numpy as np
import pytest

import orc
from shimmer_b200 import scenes


def _rays(n, seed, sc=None):
    rng = np.random.default_rng(seed)
    root = sc.arrays["nodes"][0]                       # render space = camera-world: sample inside the scene bounds
    lo, hi = np.array(root["bmin"], np.float32), np.array(root["bmax"], np.float32)
    o = (lo + (hi - lo) * rng.random((n, 3))).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def test_fixed_mode_equals_flattened_scene():
    inst = scenes.instanced_tiny_scene("instfix", (24, 24)).build()
    flat = scenes.instanced_tiny_scene("instfix", (24, 24), flatten=True).build()
    assert inst.meta["n_instanced_triangles"] == flat.meta["n_triangles"] and inst.meta["n_triangles"] < flat.meta["n_triangles"]
    o, d = _rays(20000, 3, flat)
    tmax = np.full(len(o), np.inf, np.float32)
    a, _ = orc.trace(inst, o, d, tmax); b, _ = orc.trace(flat, o, d, tmax)
    assert np.array_equal(a["prim"] >= 0, b["prim"] >= 0) or ((a["prim"] >= 0) != (b["prim"] >= 0)).mean() < 1e-3
    both = (a["prim"] >= 0) & (b["prim"] >= 0)
    assert both.mean() > 0.3
    assert np.allclose(a["t"][both], b["t"][both], rtol=2e-4, atol=1e-5)          # instance-space vs baked-vertex rounding
    sa, _ = orc.trace(inst, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    sb, _ = orc.trace(flat, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    assert (sa["prim"] != sb["prim"]).mean() < 1e-3
    fa, sta, _ = orc.render(inst, orc.make_params(seed=1, spp=8)); fb, stb, _ = orc.render(flat, orc.make_params(seed=1, spp=8))
    la, lb = fa[:, :3].sum(axis=1), fb[:, :3].sum(axis=1)
    assert abs(la.sum() - lb.sum()) / lb.sum() < 5e-3
    assert abs(int(sta.closest_hit_rays) - int(stb.closest_hit_rays)) < 2e-3 * stb.closest_hit_rays


def test_reference_mode_closest_hits_are_right_for_translations_but_shadows_are_not():
    """primitive.rs:159-163 vs :172-175: closest-hit uses the inverse transform, the predicate the forward one."""
    inst = scenes.instanced_tiny_scene("inst", (24, 24)).build()
    flat = scenes.instanced_tiny_scene("inst", (24, 24), flatten=True).build()
    o, d = _rays(20000, 4, flat)
    tmax = np.full(len(o), np.inf, np.float32)
    a, _ = orc.trace(inst, o, d, tmax); b, _ = orc.trace(flat, o, d, tmax)
    both = (a["prim"] >= 0) & (b["prim"] >= 0)
    assert ((a["prim"] >= 0) != (b["prim"] >= 0)).mean() < 1e-3 and np.allclose(a["t"][both], b["t"][both], rtol=2e-4, atol=1e-5)
    sa, _ = orc.trace(inst, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    sb, _ = orc.trace(flat, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    assert (sa["prim"] != sb["prim"]).mean() > 0.01                                # the forward-transform quirk is visible


def test_shapes_inside_object_definitions_equal_the_flattened_scene_in_fixed_mode():
    """Object definitions may hold spheres and bilinear patches (scene.rs:814-866 takes any Shape): with SG_SCENE_FIX_INSTANCING the
    instanced scene must trace like the same shapes placed at the top level with composed transforms."""
    inst = scenes.instanced_shapes_tiny_scene("instshapesfix", (24, 24)).build()
    flat = scenes.instanced_shapes_tiny_scene("instshapesfix", (24, 24), flatten=True).build()
    assert inst.desc.n_objects == 2 and inst.arrays["objects"][1].n_nodes == 0 and inst.desc.n_spheres == 3 and flat.desc.n_spheres == 6
    o, d = _rays(20000, 5, flat)
    tmax = np.full(len(o), np.inf, np.float32)
    a, _ = orc.trace(inst, o, d, tmax); b, _ = orc.trace(flat, o, d, tmax)
    assert ((a["prim"] >= 0) != (b["prim"] >= 0)).mean() < 2e-3
    both = (a["prim"] >= 0) & (b["prim"] >= 0)
    assert both.mean() > 0.3 and np.allclose(a["t"][both], b["t"][both], rtol=5e-4, atol=2e-5)
    sa, _ = orc.trace(inst, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    sb, _ = orc.trace(flat, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    assert (sa["prim"] != sb["prim"]).mean() < 2e-3
    fa, _, _ = orc.render(inst, orc.make_params(seed=1, spp=64)); fb, _, _ = orc.render(flat, orc.make_params(seed=1, spp=64))
    assert abs(fa[:, :3].sum() - fb[:, :3].sum()) / fb[:, :3].sum() < 0.02
    # the literal mode differs (forward transform on shadow rays, inverse on interaction vectors)
    lit = scenes.instanced_shapes_tiny_scene("instshapes", (24, 24)).build()
    sl, _ = orc.trace(lit, o, d * np.float32(2.5), np.full(len(o), 0.9999, np.float32), any_hit=True)
    assert (sl["prim"] != sb["prim"]).mean() > 0.01


def test_single_primitive_object_has_no_aggregate():
    sc = scenes.instanced_tiny_scene("inst", (8, 8)).build()
    objs = sc.arrays["objects"]
    assert objs[0].n_nodes > 0 and objs[1].n_nodes == 0 and objs[1].n_prims == 1   # scene.rs:821-833
    assert sc.desc.n_top_primitives == 4 + 6 and sc.desc.n_instances == 6


@pytest.mark.parametrize("kind", scenes.INSTANCED_KINDS + scenes.INSTANCED_SHAPE_KINDS)
def test_instanced_scene_golden_film(kind):
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tiny_films.json")))[kind]
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    film, st, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    assert st.closest_hit_rays == gold["closest_hit_rays"] and st.shadow_rays == gold["shadow_rays"]
    assert np.allclose(film.sum(axis=0), gold["film_sum"], rtol=1e-9)
    assert np.isfinite(film).all()
