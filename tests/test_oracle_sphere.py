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
        b, _, _ = orc.render(sc, orc.make_params(seed=1, spp=64), stream_mode=1, n_threads=1)     # one generator, one tile order: deterministic
        assert np.isfinite(a).all() and a[:, :3].sum() > 0 and st.closest_hit_rays > 24 * 24 * 64
        assert abs(a[:, :3].sum() - b[:, :3].sum()) / b[:, :3].sum() < 0.05


def _unit_sphere_world(z_min=None, z_max=None):
    """Sphere::new(Transform::default(), Transform::default(), false, 1, zmin, zmax, 360) exactly: rendering space
    "world" makes render_from_object the identity like in the reference's shape tests."""
    b = SceneBuilder(rendering_space="world"); b.set_camera((0, 0, -20), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    b.add_sphere(1.0, b.diffuse(("const", 0.5)), z_min=z_min, z_max=z_max)
    assert np.array_equal(b.render_from_world.m, np.eye(4))
    return b.build()


REF_SPHERE_PREDICATES = [   # (zmin, zmax, origin, direction, expected)  -- shape/shape.rs:299-342
    (None, None, (0, 0, -2), (0, 0, 1), True), (None, None, (0, 0, -2), (0, 0, -1), False), (None, None, (0, 1.0001, -2), (0, 0, 1), False),
    (-0.5, 0.5, (0, -2, 0), (0, 1, 0), True), (-0.5, 0.5, (0, -2, 0), (0, -1, 0), False),
    (-0.5, 0.5, (0, 0, 0.5001), (0, 1, 0), False), (-0.5, 0.5, (0, 0, -0.5001), (0, 1, 0), False)]


def test_reference_sphere_basic_and_partial_predicates():
    """shape.rs `sphere_basic` / `sphere_partial_basic`: intersect_predicate of the unit sphere (full, and clipped to
    |z| <= 0.5) for the seven rays the reference asserts."""
    for zmin, zmax, o, d, want in REF_SPHERE_PREDICATES:
        sc = _unit_sphere_world(zmin, zmax)
        h, _ = orc.trace(sc, np.array([o], np.float32), np.array([d], np.float32), [np.inf], any_hit=True)
        assert (h["prim"][0] >= 0) == want, (zmin, zmax, o, d)


def test_sphere_light_sampling_is_self_consistent():
    """Sphere::sample_with_context / pdf_with_context (sphere.rs:339-456) through the light interface.  Outside the sphere:
    the sample lies on the sphere, inside the subtended cone, with the cone pdf 1/(2 pi (1 - cos theta_max)); pdf_li returns
    that pdf times 2/2.90 (the reference's constant, :455); the cone pdf integrates to 1 (Monte Carlo).  Inside the sphere
    (the big shell): area sampling, pdf = dist^2 / (area |cos|); pdf_li is ((1/area)/|cos|)/dist^2 as written (:436-438)."""
    import ctypes as C
    b = scenes.sphere_tiny_scene("spherelight"); sc = b.build()
    L = orc.lib(); rng = np.random.default_rng(2)
    lights = sc.arrays["lights"]; spheres = sc.arrays["spheres"]
    lam = orc.fa([450, 520, 600, 680]); out = np.zeros(14, np.float32)
    small = [i for i in range(sc.meta["n_lights"]) if spheres[lights[i].tri].radius < 1.0 and spheres[lights[i].tri].phi_max > 6.0][0]
    shell = [i for i in range(sc.meta["n_lights"]) if spheres[lights[i].tri].radius > 5.0][0]
    S = spheres[lights[small].tri]; centre = np.array([S.render_from_object[3], S.render_from_object[7], S.render_from_object[11]], np.float64)
    p = (centre + np.array([1.1, -1.7, 0.6])).astype(np.float32); n = orc.fa([0, 1, 0])
    dist = np.linalg.norm(p - centre); cos_max = np.sqrt(1 - (S.radius / dist) ** 2); cone_pdf = 1 / (2 * np.pi * (1 - cos_max))
    for _ in range(200):
        u = orc.fa(rng.random(2))
        assert L.orc_light_sample(sc.ptr(), small, p.ctypes.data, n.ctypes.data, n.ctypes.data, u.ctypes.data, lam.ctypes.data, out.ctypes.data) == 1
        wi, pdf, pl = out[4:7].astype(np.float64), out[7], out[8:11].astype(np.float64)
        assert abs(np.linalg.norm(pl - centre) - S.radius) < 1e-5 and abs(pdf - cone_pdf) < 1e-4 * cone_pdf
        assert np.dot(wi, (centre - p) / dist) >= cos_max - 1e-5                       # inside the cone
        p2 = L.orc_light_pdf(sc.ptr(), small, p.ctypes.data, n.ctypes.data, n.ctypes.data, orc.fa(wi).ctypes.data)
        assert abs(p2 - pdf * 2.0 / 2.90) < 1e-4 * pdf
    # inside the shell
    S = spheres[lights[shell].tri]; centre = np.array([S.render_from_object[3], S.render_from_object[7], S.render_from_object[11]], np.float64)
    area = S.phi_max * S.radius * (S.z_max - S.z_min)
    assert abs(area - 4 * np.pi * S.radius ** 2) < 1e-3 * area and abs(lights[shell].area - area) < 1e-3 * area
    p = (centre + np.array([0.5, -0.3, 1.0])).astype(np.float32)
    acc = 0.0; cnt = 0
    for _ in range(2000):
        u = orc.fa(rng.random(2))
        if L.orc_light_sample(sc.ptr(), shell, p.ctypes.data, n.ctypes.data, n.ctypes.data, u.ctypes.data, lam.ctypes.data, out.ctypes.data) != 1:
            continue
        wi, pdf, pl, nl = out[4:7].astype(np.float64), out[7], out[8:11].astype(np.float64), out[11:14].astype(np.float64)
        d2 = np.sum((pl - p) ** 2); cosl = abs(np.dot(nl, -wi))
        assert abs(np.linalg.norm(pl - centre) - S.radius) < 1e-4 and abs(pdf - d2 / (area * cosl)) < 2e-3 * pdf
        p2 = L.orc_light_pdf(sc.ptr(), shell, p.ctypes.data, n.ctypes.data, n.ctypes.data, orc.fa(wi).ctypes.data)
        assert abs(p2 - (1 / area) / cosl / d2) < 5e-3 * p2
        acc += 1.0 / pdf; cnt += 1
    assert cnt > 1900 and abs(acc / cnt - 4 * np.pi) < 0.05 * 4 * np.pi             # solid-angle pdf integrates over the full sphere of directions


def test_sphere_scenes_are_sensitive_to_one_ulp_and_triangle_scenes_are_not():
    """Why the GPU film tests allow ~1-3 % differing pixels on scenes with rotated / scaled spheres: the reference algorithm itself
    is discontinuous at the ulp level there.  Scaling the x component of every first-bounce ray direction by (1 + 2^-23) in the ORACLE
    changes several per cent of the pixels of the sphere scene by more than 2e-3, and none of the triangle scene -- Sphere::
    basic_intersect decides hits with interval bounds (sphere.rs:95-186: `discrim.lower_bound() < 0`, `t0.lower_bound() <= 0`), so
    a last-bit change of a ray moves paths across those tests.  The CUDA path differs from the oracle by libm last bits
    (sin / cos / atan2 in the samplers), i.e. by exactly such perturbations; it stays below the oracle's own sensitivity."""
    import ctypes as C
    L = orc.lib()
    p = orc.make_params(seed=5, spp=16)
    def changed(kind):
        sc = scenes.tiny_scene(kind, resolution=(32, 32)).build()
        base, _, _ = orc.render(sc, p)
        L.orc_set_debug_perturb(C.c_float(1.1920929e-07))
        try:
            f, _, _ = orc.render(sc, p)
        finally:
            L.orc_set_debug_perturb(C.c_float(0.0))
        lg, lr = f[:, :3].sum(axis=1), base[:, :3].sum(axis=1)
        return int((np.abs(lg - lr) / np.maximum(lr, 0.05 * lr.mean()) > 2e-3).sum())
    assert changed("diffuse") == 0 and changed("conductor") == 0
    assert changed("spheres") >= 10          # observed: 45 of 1024 pixels (the CUDA path differs from the oracle in 10)
