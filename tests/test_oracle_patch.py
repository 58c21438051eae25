"""Bilinear patches in the oracle (bilinear_patch.rs:144-425).  The reference has no unit test for this
shape, so the restatement is checked against closed forms: the hit point of a (u, v, t) triple lies on the ray and on the
bilinear surface, a planar patch agrees with its two triangles, normals match the analytic surface normal."""
import numpy as np

import orc
from shimmer_b200 import scenes
from shimmer_b200.host import SceneBuilder


def _single(P, uv=None, n=None):
    b = SceneBuilder(rendering_space="world"); b.set_camera((0, 0, -20), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    b.add_bilinear_mesh(np.asarray(P, np.float32), [[0, 1, 2, 3]], b.diffuse(("const", 0.5)), uv=uv, n=n)
    return b.build()


def test_patch_hits_lie_on_ray_and_surface():
    P = np.array([[0, 0, 0], [1, 0, 0.2], [0, 1, -0.1], [1.2, 1.1, 0.6]], np.float64)
    sc = _single(P)
    rng = np.random.default_rng(5); n = 20000
    o = np.stack([rng.uniform(-0.3, 1.5, n), rng.uniform(-0.3, 1.5, n), np.full(n, -3.0)], 1).astype(np.float32)
    d = (np.stack([rng.uniform(-0.2, 1.4, n), rng.uniform(-0.2, 1.4, n), rng.uniform(0.0, 0.5, n)], 1) - o).astype(np.float32)
    h, _ = orc.trace(sc, o, d, np.full(n, np.inf, np.float32))
    hit = h["prim"] >= 0
    assert 0.3 < hit.mean() < 0.95
    u, v, t = h["b0"][hit].astype(np.float64), h["b1"][hit].astype(np.float64), h["t"][hit].astype(np.float64)
    assert ((u >= 0) & (u <= 1) & (v >= 0) & (v <= 1) & (t > 0)).all()
    on_ray = o[hit].astype(np.float64) + t[:, None] * d[hit].astype(np.float64)
    surf = ((1 - u) * (1 - v))[:, None] * P[0] + (u * (1 - v))[:, None] * P[1] + ((1 - u) * v)[:, None] * P[2] + (u * v)[:, None] * P[3]
    assert np.abs(on_ray - surf).max() < 5e-5
    # geometric normal = normalised dp/du x dp/dv of the bilinear surface
    dpdu = (1 - v)[:, None] * (P[1] - P[0]) + v[:, None] * (P[3] - P[2]); dpdv = (1 - u)[:, None] * (P[2] - P[0]) + u[:, None] * (P[3] - P[1])
    nn = np.cross(dpdu, dpdv); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    assert np.abs(h["ng"][hit] - nn).max() < 1e-4
    any_h, _ = orc.trace(sc, o, d, np.full(n, np.inf, np.float32), any_hit=True)
    assert np.array_equal(any_h["prim"] >= 0, hit)


def test_planar_patch_matches_its_two_triangles():
    quad = np.array([[-1, -1, 0.5], [1, -1, 0.5], [-1, 1, 0.5], [1, 1, 0.5]], np.float32)
    sc_p = _single(quad)
    b = SceneBuilder(rendering_space="world"); b.set_camera((0, 0, -20), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    b.add_mesh(quad, [[0, 1, 3], [0, 3, 2]], b.diffuse(("const", 0.5)))
    sc_t = b.build()
    rng = np.random.default_rng(8); n = 10000
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32); o[:, 2] = -2.0
    d = rng.standard_normal((n, 3)).astype(np.float32); d[:, 2] = np.abs(d[:, 2]) + 0.2
    hp, _ = orc.trace(sc_p, o, d, np.full(n, np.inf, np.float32)); ht, _ = orc.trace(sc_t, o, d, np.full(n, np.inf, np.float32))
    both = (hp["prim"] >= 0) & (ht["prim"] >= 0)
    assert ((hp["prim"] >= 0) != (ht["prim"] >= 0)).mean() < 2e-3 and both.mean() > 0.05         # only edge-grazing rays may differ
    assert np.allclose(hp["t"][both], ht["t"][both], rtol=2e-5)
    assert np.allclose(np.abs(hp["ng"][both]), np.abs(ht["ng"][both]), atol=1e-6)


def test_patch_scene_renders():
    for kind in scenes.PATCH_KINDS:
        sc = scenes.patch_tiny_scene(kind, resolution=(24, 24)).build()
        assert sc.meta["n_patches"] >= 38
        a, st, _ = orc.render(sc, orc.make_params(seed=1, spp=32), stream_mode=0)
        b, _, _ = orc.render(sc, orc.make_params(seed=1, spp=32), stream_mode=1, n_threads=1)     # deterministic tile order
        assert np.isfinite(a).all() and a[:, :3].sum() > 0
        assert abs(a[:, :3].sum() - b[:, :3].sum()) / b[:, :3].sum() < 0.08


def _rect_light_scene(side=1.2, height=3.0):
    b = SceneBuilder(rendering_space="world"); b.set_camera((0, 0.5, -4), (0, 0.5, 0), (0, 1, 0), 40.0, (8, 8))
    h = side / 2
    lq = np.array([[-h, height, -h], [h, height, -h], [-h, height, h], [h, height, h]], np.float32)
    b.add_bilinear_mesh(lq, [[0, 1, 2, 3]], b.diffuse(("const", 0.5)), area_light=dict(L=("const", 1.0), scale=1.0, two_sided=False))
    gp = np.array([[-3, 0, -3], [-3, 0, 3], [3, 0, 3], [3, 0, -3]], np.float32)
    b.add_mesh(gp, [[0, 1, 2], [0, 2, 3]], b.diffuse(("const", 0.5)))
    return b.build()


def test_rectangular_patch_light_is_sampled_by_solid_angle():
    """BilinearPatch::sample_with_context / pdf_with_context for a rectangle (bilinear_patch.rs:666-736, 770-782): without a shading
    normal in the context the pdf is 1 / solid angle; on the axis of a square of side a at distance d the solid angle is
    4 asin(a^2 / (a^2 + 4 d^2)).  Samples land on the rectangle and the pdf of their direction is the pdf they reported."""
    a, d = 1.2, 3.0
    sc = _rect_light_scene(a, d)
    L = orc.lib()
    z = np.zeros(3, np.float32); lam = np.array([450, 520, 600, 680], np.float32); out = np.zeros(14, np.float32)
    omega = 4.0 * np.arcsin(a * a / (a * a + 4.0 * d * d))
    rng = np.random.default_rng(3)
    for _ in range(300):
        u = rng.random(2).astype(np.float32)
        assert L.orc_light_sample(sc.ptr(), 0, z.ctypes.data, z.ctypes.data, z.ctypes.data, u.ctypes.data, lam.ctypes.data, out.ctypes.data)
        assert abs(out[7] * omega - 1.0) < 2e-3
        p = out[8:11]
        assert abs(p[1] - d) < 1e-4 and abs(p[0]) <= a / 2 + 1e-4 and abs(p[2]) <= a / 2 + 1e-4
        pdf = L.orc_light_pdf(sc.ptr(), 0, z.ctypes.data, z.ctypes.data, z.ctypes.data, out[4:7].copy().ctypes.data)
        assert abs(pdf - out[7]) < 1e-4 * out[7]
    # with a shading normal the bilinear warp enters both (sample: sample_bilinear + bilinear_pdf; pdf: invert_spherical_rectangle_sample)
    ns = np.array([0.3, 0.9, 0.1], np.float32); ns /= np.linalg.norm(ns)
    ref = np.array([0.4, 0.0, -0.2], np.float32)
    bad = 0
    for _ in range(300):
        u = rng.random(2).astype(np.float32)
        assert L.orc_light_sample(sc.ptr(), 0, ref.ctypes.data, ns.ctypes.data, ns.ctypes.data, u.ctypes.data, lam.ctypes.data, out.ctypes.data)
        pdf = L.orc_light_pdf(sc.ptr(), 0, ref.ctypes.data, ns.ctypes.data, ns.ctypes.data, out[4:7].copy().ctypes.data)
        bad += abs(pdf - out[7]) > 2e-2 * out[7]
    assert bad <= 6            # "this (rarely) goes differently than sample" (sampling.rs:689)
    # directions that miss the patch have zero density
    miss = np.array([1.0, 0.0, 0.0], np.float32)
    assert L.orc_light_pdf(sc.ptr(), 0, z.ctypes.data, z.ctypes.data, z.ctypes.data, miss.ctypes.data) == 0.0


def test_rectangular_patch_light_nee_is_unbiased():
    """Path integrator with next-event estimation on the emissive rectangle vs SimplePath without light sampling (pure BSDF
    sampling finds the emitter by chance): same mean."""
    sc = scenes.patch_tiny_scene("patchlight", resolution=(12, 12)).build()
    a, _, _ = orc.render(sc, orc.make_params(seed=0, spp=512))
    b, _, _ = orc.render(sc, orc.make_params(seed=0, spp=4096, integrator="simplepath", sample_lights=False))
    ma, mb = a[:, :3].sum() / a[:, 3].sum(), b[:, :3].sum() / b[:, 3].sum()
    assert abs(ma - mb) / mb < 0.03, (ma, mb)


def test_patch_helpers_closed_form():
    from shimmer_b200 import host
    sq = np.array([[0, 0, 0], [2, 0, 0], [0, 3, 0], [2, 3, 0]], np.float32)
    assert host.bilinear_patch_is_rectangle(sq) and abs(host.bilinear_patch_area(sq) - 6.0) < 1e-6
    tw = np.array([[0, 0, 0], [2, 0, 0], [0, 3, 0], [2, 3, 1]], np.float32)
    assert not host.bilinear_patch_is_rectangle(tw) and 6.0 < host.bilinear_patch_area(tw) < 6.6
    par = np.array([[0, 0, 0], [2, 0, 0], [1, 3, 0], [3, 3, 0]], np.float32)       # planar parallelogram: not a rectangle
    assert not host.bilinear_patch_is_rectangle(par)
