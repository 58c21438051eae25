"""Sphere shape in the oracle, pinned by the reference's own BVH-over-spheres tests
(aggregate.rs:574-702): the only golden vectors the reference holds for traversal."""
import numpy as np

import orc
from shimmer_b200 import scenes
from shimmer_b200.host import SceneBuilder, Transform


def _spheres(mults):
    b = SceneBuilder(); b.set_camera((0, 0, -20), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    m = b.diffuse(("const", 0.5))
    for mu in mults:
        b.add_sphere(1.0, m, object_from_world=Transform.translate((mu, 0, 0)))
    return b


def _render_space(b, p):
    return b.render_from_world.apply_points_f32(np.array([p], np.float32))


def test_reference_single_sphere_bvh_intersection():
    """aggregate.rs:604-641 `single_primitive_bvh_intersetion`: ray from (-5,0,0) along +x hits the unit sphere at
    t = 4 (assert_approx_eq default tolerance), p = (-1,0,0) +- 1e-6, shading normal = -x."""
    b = _spheres([0.0]); sc = b.build()
    assert sc.meta["n_nodes"] == 1                                         # single_primitive_bvh (:574-601): one leaf, bounds of the primitive
    root = sc.arrays["nodes"][0]
    assert np.array_equal(root["bmin"], _render_space(b, [-1, -1, -1])[0]) and np.array_equal(root["bmax"], _render_space(b, [1, 1, 1])[0])
    h, _ = orc.trace(sc, _render_space(b, [-5, 0, 0]), [[1, 0, 0]], [np.inf])
    assert h["prim"][0] == 0 and abs(h["t"][0] - 4.0) <= 4 * np.spacing(np.float32(4.0))
    assert abs(h["b0"][0] + 1.0) <= 1e-6 and h["b1"][0] == 0.0 and h["b2"][0] == 0.0          # p_obj == hit point (identity transform)
    assert np.dot(h["ng"][0], [-1, 0, 0]) == 1.0


def test_reference_set_of_spheres():
    """aggregate.rs:643-702 `set_of_spheres`: centres x = -3.5, 0, 5; ray from (-10,0,0) along +x hits at t = 5.5 +- 1e-5,
    p.x = -4.5 +- 1e-5, normal -x; predicate true; the ray offset by z = 1.001 misses (closest and predicate)."""
    b = _spheres([-3.5, 0.0, 5.0]); sc = b.build()
    o = _render_space(b, [-10, 0, 0]); d = [[1, 0, 0]]
    h, _ = orc.trace(sc, o, d, [np.inf])
    sph = sc.arrays["spheres"][sc.arrays["prims"]["tri"][h["prim"][0]]]
    assert abs(h["t"][0] - 5.5) <= 1e-5 and sph.render_from_object[3] == -3.5
    p_render_x = h["b0"][0] + sph.render_from_object[3]
    assert abs(p_render_x + 4.5) <= 1e-5 and h["b1"][0] == 0.0
    assert np.dot(h["ng"][0], [-1, 0, 0]) == 1.0
    assert orc.trace(sc, o, d, [np.inf], any_hit=True)[0]["prim"][0] == 0
    o2 = _render_space(b, [-10, 0, 1.001])
    assert orc.trace(sc, o2, d, [np.inf])[0]["prim"][0] == -1
    assert orc.trace(sc, o2, d, [np.inf], any_hit=True)[0]["prim"][0] == -1


def test_sphere_hits_against_closed_form():
    """Hits of random rays on transformed full spheres agree with the analytic ray/ellipsoid intersection in f64
    (t within 1e-4 relative), t_max and origin-inside cases included; clipped spheres only ever lose hits."""
    b = scenes.sphere_tiny_scene("spheres"); sc = b.build()
    rng = np.random.default_rng(3)
    n = 20000
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.05
    centres = np.array([[-1.3, 0.6, 0.4], [0.1, 0.55, -0.6], [1.4, 0.75, 0.5]], np.float32)
    d = (centres[rng.integers(0, 3, n)] + rng.uniform(-0.7, 0.7, (n, 3)).astype(np.float32) - o).astype(np.float32)   # unnormalised, like shadow rays
    o[: n // 10] = centres[rng.integers(0, 3, n // 10)] + rng.uniform(-0.2, 0.2, (n // 10, 3)).astype(np.float32)   # origins inside a sphere
    o = b.render_from_world.apply_points_f32(o)
    h, _ = orc.trace(sc, o, d, np.full(n, np.inf, np.float32))
    prims = sc.arrays["prims"]
    is_sph = (h["prim"] >= 0) & (prims["mesh"][np.maximum(h["prim"], 0)] == 0xfffffffe)
    assert is_sph.sum() > 2000
    checked = 0
    for i in np.nonzero(is_sph)[0][:3000]:
        S = sc.arrays["spheres"][prims["tri"][h["prim"][i]]]
        Mi = np.array(S.object_from_render, np.float64).reshape(4, 4)
        oo = Mi[:3, :3] @ o[i].astype(np.float64) + Mi[:3, 3]; dd = Mi[:3, :3] @ d[i].astype(np.float64)
        a, bq, c = dd @ dd, 2 * oo @ dd, oo @ oo - S.radius ** 2
        disc = bq * bq - 4 * a * c
        assert disc > -1e-4
        r = np.sort([(-bq - np.sqrt(max(disc, 0))) / (2 * a), (-bq + np.sqrt(max(disc, 0))) / (2 * a)])
        assert min(abs(h["t"][i] - r[0]), abs(h["t"][i] - r[1])) <= 2e-4 * max(1.0, abs(h["t"][i]))
        p_obj = np.array([h["b0"][i], h["b1"][i], h["b2"][i]])
        assert abs(np.linalg.norm(p_obj) - S.radius) <= 1e-5 * S.radius
        if S.z_min > -S.radius: assert p_obj[2] >= S.z_min
        if S.z_max < S.radius: assert p_obj[2] <= S.z_max
        checked += 1
    assert checked > 1000


def test_sphere_scene_renders_and_modes_agree():
    """Film sanity: energy arrives, both RNG modes converge to the same mean, uv/texture variant runs."""
    for kind in scenes.SPHERE_KINDS:
        sc = scenes.sphere_tiny_scene(kind, resolution=(24, 24)).build()
        a, st, _ = orc.render(sc, orc.make_params(seed=1, spp=64), stream_mode=0)
        b, _, _ = orc.render(sc, orc.make_params(seed=1, spp=64), stream_mode=1)
        assert np.isfinite(a).all() and a[:, :3].sum() > 0 and st.closest_hit_rays > 24 * 24 * 64
        assert abs(a[:, :3].sum() - b[:, :3].sum()) / b[:, :3].sum() < 0.05
