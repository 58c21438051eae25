"""SURVEY 8f next-4 on the CPU oracle: image-infinite light (light.rs:805-981), CoatedConductor (material.rs:995-1283),
normal maps (:1453-1474), non-UV texture mappings (texture.rs:938-1035), MixMaterial (material.rs:1286-1330) and the
SimplePath / RandomWalk integrators (integrator.rs:458-728).  The reference holds no test vectors for any of these, so the
checks are closed-form values and consistency properties."""
import ctypes as C

import numpy as np
import pytest

import orc
from shimmer_b200 import ffi, host, scenes
from shimmer_b200.host import SceneBuilder, Transform, named_spectrum

LAM = np.array([450.0, 520.0, 600.0, 680.0], np.float32)
Z3 = np.zeros(3, np.float32)


def _sq2sph(p):
    p = np.asarray(p, np.float32); w = np.zeros(3, np.float32)
    orc.lib().orc_equal_area_square_to_sphere(p.ctypes.data, w.ctypes.data)
    return w


def _sph2sq(d):
    d = np.asarray(d, np.float32); q = np.zeros(2, np.float32)
    orc.lib().orc_equal_area_sphere_to_square(d.ctypes.data, q.ctypes.data)
    return q


def test_equal_area_maps_closed_form():
    """math.rs:453-530.  Centre of the square is the +z pole, the corners are the -z pole; the inverse (which is pbrt's,
    unmodified) maps the axes to the edge midpoints of the inner diamond."""
    assert np.allclose(_sq2sph([0.5, 0.5]), [0, 0, 1], atol=1e-7)
    for c in ([0, 0], [1, 0], [0, 1], [1, 1]):
        assert np.allclose(_sq2sph(c), [0, 0, -1], atol=1e-6)
    assert np.allclose(_sph2sq([0, 0, 1]), [0.5, 0.5], atol=1e-7)
    assert np.allclose(_sph2sq([1, 0, 0]), [1.0, 0.5], atol=1e-5)
    assert np.allclose(_sph2sq([0, 1, 0]), [0.5, 1.0], atol=1e-5)
    assert np.allclose(_sph2sq([-1, 0, 0]), [0.0, 0.5], atol=1e-5)     # 6th-degree minimax atan: ~2e-6
    rng = np.random.default_rng(0)
    for _ in range(500):
        w = _sq2sph(rng.random(2))
        assert abs(np.linalg.norm(w) - 1.0) < 1e-6                 # always a unit vector, also with the `vp - up / r + 1` slip
    # the slip (math.rs:472) only changes the azimuth inside a quadrant: z and the quadrant survive a round trip
    for _ in range(500):
        p = rng.random(2).astype(np.float32)
        w = _sq2sph(p); q = _sph2sq(w)
        assert np.sign(q[0] - 0.5) == np.sign(p[0] - 0.5) or abs(p[0] - 0.5) < 1e-3
        assert abs(_sq2sph(q)[2] - w[2]) < 1e-5
    # pbrt's forward map agrees wherever up / r == up, i.e. on the diamond |u| + |v| = 1 (r = 1)
    assert np.allclose(_sq2sph([0.75, 0.25]), _pbrt_sq2sph(0.75, 0.25), atol=1e-6)


def _pbrt_sq2sph(px, py):
    u, v = 2 * px - 1, 2 * py - 1
    up, vp = abs(u), abs(v)
    sd = 1 - (up + vp); r = 1 - abs(sd)
    phi = (1.0 if r == 0 else (vp - up) / r + 1) * np.pi / 4
    z = np.copysign(1 - r * r, sd)
    return np.array([np.copysign(np.cos(phi), u) * r * np.sqrt(2 - r * r), np.copysign(np.sin(phi), v) * r * np.sqrt(2 - r * r), z])


def test_piecewise_constant_2d_construction_and_sampling():
    """sampling.rs:11-179: cdf rows end at 1, marginal integral = mean of the function, sample() returns a point whose
    pdf() is the value sample() reported, and samples are distributed like the function."""
    rng = np.random.default_rng(1)
    f = rng.random((8, 16)).astype(np.float32) ** 3
    f[2, :] = 0.0                                                  # an all-zero row: cdf falls back to i / n (sampling.rs:44-48)
    fn, cdf, mf, mcdf, mint = host.piecewise_constant_2d(f)
    assert np.allclose(cdf[:, -1], 1.0, atol=1e-6) and abs(mcdf[-1] - 1.0) < 1e-6
    assert np.allclose(cdf[2], np.arange(17) / 16.0)
    assert abs(mint - f.mean()) < 1e-6
    sc = scenes.tiny_scene("envmap", (8, 8)).build()
    L = orc.lib()
    out = np.zeros(14, np.float32)
    img = scenes.procedural_envmap(32)
    d = img.mean(axis=2); comp = np.maximum(d - d.mean(), 0)
    for _ in range(4000):
        u = rng.random(2).astype(np.float32)
        assert L.orc_light_sample(sc.ptr(), 0, Z3.ctypes.data, Z3.ctypes.data, Z3.ctypes.data, u.ctypes.data, LAM.ctypes.data, out.ctypes.data)
        assert out[7] > 0 and np.isfinite(out[:8]).all()
    # pdf over the sphere integrates to one for both distributions
    dirs = rng.standard_normal((20000, 3)).astype(np.float32); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    full = np.mean([L.orc_light_pdf_complete(sc.ptr(), 0, w.ctypes.data) for w in dirs]) * 4 * np.pi
    compd = np.mean([L.orc_light_pdf(sc.ptr(), 0, Z3.ctypes.data, Z3.ctypes.data, Z3.ctypes.data, w.ctypes.data) for w in dirs]) * 4 * np.pi
    assert abs(full - 1) < 0.03 and abs(compd - 1) < 0.06
    # the compensated pdf is zero wherever the image is below its mean (light.rs:941-948)
    zero_frac = np.mean([L.orc_light_pdf(sc.ptr(), 0, Z3.ctypes.data, Z3.ctypes.data, Z3.ctypes.data, w.ctypes.data) == 0 for w in dirs[:4000]])
    assert abs(zero_frac - (comp == 0).mean()) < 0.05


def test_constant_environment_map_equals_uniform_infinite_light():
    """A grey map c goes through RgbIlluminantSpectrum (spectrum.rs:566-606): 2c * sigmoid(0) * D65 = c * D65, the
    compensated distribution degenerates to uniform (light.rs:944-947), so the image matches the uniform infinite
    light with L = D65 up to noise."""
    def build(kind):
        b = SceneBuilder()
        b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=(12, 12))
        if kind == "image":
            b.add_image_infinite_light(np.full((8, 8, 3), 0.7, np.float32), scale=1.0)
        else:
            b.add_uniform_infinite_light(named_spectrum("stdillum-D65"), scale=0.7)
        P, I, Nn, UV = scenes.uv_sphere(8, 10, center=(0.0, 0.6, 0.0), radius=0.6)
        b.add_mesh(P, I, b.diffuse(scenes._white()), n=Nn, uv=UV)
        gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, b.diffuse(scenes._green()))
        return b.build()
    a, _, _ = orc.render(build("image"), orc.make_params(seed=1, spp=256))
    u, _, _ = orc.render(build("uniform"), orc.make_params(seed=1, spp=256))
    ma, mu = a[:, :3].sum(axis=0), u[:, :3].sum(axis=0)
    assert np.allclose(ma, mu, rtol=0.03), (ma, mu)
    # escaped camera rays see exactly c * D65 in both (no sampling involved): compare a background pixel
    bg = np.argmax(u[:, 1])
    assert np.allclose(a[bg, :3], u[bg, :3], rtol=0.02)


def _mean_rgb(sc, spp=128, **kw):
    f, st, _ = orc.render(sc, orc.make_params(seed=2, spp=spp, **kw))
    assert np.isfinite(f).all()
    return f[:, :3].sum(axis=0) / f[:, 3].sum()


def _one_material_scene(make_mat, res=(12, 12), textured=False):
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=res)
    mat = make_mat(b)
    P, I, Nn, UV = scenes.uv_sphere(10, 14, center=(0.0, 0.6, 0.0), radius=0.6)
    b.add_mesh(P, I, mat, n=Nn, uv=UV)
    gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, b.diffuse(scenes._white()), uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    lp, li = scenes._quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
    b.add_mesh(lp, li, b.diffuse(scenes._white()), area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b.build()


def test_coated_conductor_with_index_matched_coat_is_the_conductor():
    """interface eta = 1: the top DielectricBxDF is a pure specular pass-through (bxdf.rs:778-790), albedo = 0 and a thin layer
    -> the LayeredBxDF random walk reproduces the bottom ConductorBxDF.  remaproughness = false so that the conductor roughness is
    the one given (with remapping the reference derives it from the interface roughness, material.rs:1237-1241)."""
    cu = (named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"))
    coated = _one_material_scene(lambda b: b.coated_conductor(*cu, interface_eta=("const", 1.0), interface_roughness=0.0, conductor_roughness=0.3,
                                                              thickness=1e-4, remap=False))
    plain = _one_material_scene(lambda b: b.conductor(*cu, roughness=0.3, remap=False))
    a, p = _mean_rgb(coated, 256), _mean_rgb(plain, 256)
    assert np.allclose(a, p, rtol=0.03), (a, p)


def test_coated_conductor_reflectance_form_and_roughness_quirk():
    D = _one_material_scene(lambda b: b.coated_conductor(reflectance=scenes._red(), interface_roughness=0.3, conductor_roughness=0.9))
    E = _one_material_scene(lambda b: b.coated_conductor(reflectance=scenes._red(), interface_roughness=0.3, conductor_roughness=0.0))
    fd, _, _ = orc.render(D, orc.make_params(seed=3, spp=8)); fe, _, _ = orc.render(E, orc.make_params(seed=3, spp=8))
    assert fd.tobytes() == fe.tobytes()          # with remaproughness the conductor roughness parameter is ignored (material.rs:1237-1241)
    F = _one_material_scene(lambda b: b.coated_conductor(reflectance=scenes._red(), interface_roughness=0.3, conductor_roughness=0.9, remap=False))
    ff, _, _ = orc.render(F, orc.make_params(seed=3, spp=8))
    assert ff.tobytes() != fd.tobytes() and np.isfinite(ff).all()


def test_mix_material_limits_and_proportions():
    """amount <= 0 -> first material, amount >= 1 -> second (no random number drawn, material.rs:1313-1322); in between the
    second material is chosen with probability `amount`."""
    green = lambda b: b.diffuse(scenes._green())
    cu = lambda b: b.conductor(named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"), roughness=0.05)
    fa, _, _ = orc.render(_one_material_scene(green), orc.make_params(seed=4, spp=8))
    fb, _, _ = orc.render(_one_material_scene(cu), orc.make_params(seed=4, spp=8))
    f0, _, _ = orc.render(_one_material_scene(lambda b: b.mix(green(b), cu(b), amount=0.0)), orc.make_params(seed=4, spp=8))
    f1, _, _ = orc.render(_one_material_scene(lambda b: b.mix(green(b), cu(b), amount=1.0)), orc.make_params(seed=4, spp=8))
    fn, _, _ = orc.render(_one_material_scene(lambda b: b.mix(b.mix(cu(b), green(b), amount=2.0), cu(b), amount=-1.0)), orc.make_params(seed=4, spp=8))
    assert f0.tobytes() == fa.tobytes() and f1.tobytes() == fb.tobytes() and fn.tobytes() == fa.tobytes()
    # a mix of an emitter-free black diffuse and a white diffuse: first-bounce radiance scales with `amount`
    black = lambda b: b.diffuse(("const", 0.0)); white = lambda b: b.diffuse(("const", 0.8))
    mw = _mean_rgb(_one_material_scene(white), 128, max_depth=1)
    mk = _mean_rgb(_one_material_scene(black), 128, max_depth=1)
    mm = _mean_rgb(_one_material_scene(lambda b: b.mix(black(b), white(b), amount=0.25)), 128, max_depth=1)
    assert np.allclose(mm, 0.75 * mk + 0.25 * mw, rtol=0.04), (mm, mk, mw)


def test_flat_normal_map_changes_nothing_but_rounding():
    flat = np.zeros((4, 4, 3), np.float32); flat[:, :, :2] = 0.5; flat[:, :, 2] = 1.0
    cu = (named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k"))
    with_map = _one_material_scene(lambda b: b.conductor(*cu, roughness=0.2, normal_map=b.image_texture(flat)))
    without = _one_material_scene(lambda b: b.conductor(*cu, roughness=0.2))
    a, p = _mean_rgb(with_map, 64), _mean_rgb(without, 64)
    assert np.allclose(a, p, rtol=0.02), (a, p)
    bumpy = _one_material_scene(lambda b: b.conductor(*cu, roughness=0.2, normal_map=b.image_texture(scenes.procedural_normal_map(32))))
    fb, _, _ = orc.render(bumpy, orc.make_params(seed=2, spp=8)); fw, _, _ = orc.render(without, orc.make_params(seed=2, spp=8))
    assert fb.tobytes() != fw.tobytes()


def test_planar_mapping_reproduces_the_uv_mapping_of_an_axis_aligned_quad():
    """The ground quad's uv is ((x + 3) / 6, (z + 3) / 6): a planar mapping with v1 = (1/6, 0, 0), v2 = (0, 0, 1/6) and deltas
    0.5 (texture.rs:1003-1035) gives the same (s, t); only the footprint estimate differs (ds/dp . dp/dx instead of su du/dx)."""
    img = scenes.procedural_image(64, 3)

    def build(mapped):
        b = SceneBuilder()
        b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=(16, 16))
        mp = b.texture_mapping("planar", v1=(1 / 6, 0, 0), v2=(0, 0, 1 / 6), udelta=0.5, vdelta=0.5) if mapped else None
        ground = b.diffuse(scenes._white(), reflectance_tex=b.image_texture(img, filter="bilinear", mapping=mp))
        gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
        b.add_mesh(gp, gi, ground, uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
        lp, li = scenes._quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
        b.add_mesh(lp, li, b.diffuse(scenes._white()), area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
        return b.build()
    a = _mean_rgb(build(True), 64); p = _mean_rgb(build(False), 64)
    assert np.allclose(a, p, rtol=0.02), (a, p)


def test_spherical_and_cylindrical_mappings_as_written():
    """texture.rs:943-1001: st of the spherical mapping is (theta / pi, theta / 2pi) with theta = asin(z) (safe_acos calls asin),
    the cylindrical s is pi + atan2(y, x) / 2pi.  A one-texel-wide stripe image makes the lookup position observable."""
    n = 64
    img = np.zeros((n, n), np.float32); img[:, n // 4: n // 2] = 1.0     # bright for s in [0.25, 0.5)
    b = SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    ts = b.image_texture(img, filter="point", mapping=b.texture_mapping("spherical"))
    tc = b.image_texture(img, filter="point", mapping=b.texture_mapping("cylindrical"))
    b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), b.diffuse(("const", 0.5), reflectance_tex=ts))
    sc = b.build()
    rfw = b.render_from_world
    rng = np.random.default_rng(5)
    pw = rng.standard_normal((256, 3)).astype(np.float32)
    pr = rfw.apply_points_f32(pw)
    got_s = orc.texture_eval_p(sc, ts, pr, as_float=True)[:, 0]
    got_c = orc.texture_eval_p(sc, tc, pr, as_float=True)[:, 0]
    unit = pw / np.linalg.norm(pw, axis=1, keepdims=True)
    s_sph = np.arcsin(np.clip(unit[:, 2], -1, 1)) / np.pi
    exp_s = ((np.mod(s_sph, 1.0) >= 0.25) & (np.mod(s_sph, 1.0) < 0.5)).astype(np.float32)
    s_cyl = np.pi + np.arctan2(pw[:, 1], pw[:, 0]) / (2 * np.pi)
    exp_c = ((np.mod(s_cyl, 1.0) >= 0.25) & (np.mod(s_cyl, 1.0) < 0.5)).astype(np.float32)
    assert (got_s == exp_s).mean() > 0.97 and (got_c == exp_c).mean() > 0.97      # texel-boundary cases may round the other way


@pytest.mark.parametrize("integ,sl,sb", [("simplepath", True, True), ("simplepath", False, True), ("randomwalk", True, True)])
def test_other_integrators_converge_to_the_path_integrator(integ, sl, sb):
    """SimplePath with BSDF sampling and RandomWalk estimate the same integral as PathIntegrator (integrator.rs:458-728)."""
    sc = scenes.cornell_box(resolution=(16, 16)).build()
    ref = _mean_rgb(sc, 512)
    got = _mean_rgb(sc, 1024 if integ == "randomwalk" else 512, integrator=integ, sample_lights=sl, sample_bsdf=sb)
    assert np.allclose(got, ref, rtol=0.06 if integ == "randomwalk" else 0.03), (got, ref)


def test_simplepath_uniform_sampling_uses_the_reference_pdf_constants():
    """samplebsdf = false: hemisphere sampling around +z of RENDER space with pdf 1 / (4 pi) (sampling.rs:295-308) -- as written,
    so the image is brighter than the path integrator's; it must still be finite, deterministic and draw no light samples
    when samplelights = false."""
    sc = scenes.cornell_box(resolution=(16, 16)).build()
    f1, s1, _ = orc.render(sc, orc.make_params(seed=9, spp=8, integrator="simplepath", sample_lights=False, sample_bsdf=False))
    f2, s2, _ = orc.render(sc, orc.make_params(seed=9, spp=8, integrator="simplepath", sample_lights=False, sample_bsdf=False), n_threads=1)
    assert f1.tobytes() == f2.tobytes() and s1.shadow_rays == 0 and np.isfinite(f1).all()
    f3, s3, _ = orc.render(sc, orc.make_params(seed=9, spp=8, integrator="simplepath", sample_lights=True, sample_bsdf=False))
    assert s3.shadow_rays > 0


@pytest.mark.parametrize("kind", scenes.VARIETY_KINDS)
def test_variety_scene_golden_film(kind):
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_films.json")))[kind]
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    film, st, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    assert st.closest_hit_rays == gold["closest_hit_rays"] and st.shadow_rays == gold["shadow_rays"]
    assert np.allclose(film.sum(axis=0), gold["film_sum"], rtol=1e-9) and np.isfinite(film).all()


def test_generate_pyramid_oracle_properties():
    """Image::generate_pyramid (image.rs:699-787): power-of-two images keep level 0, every level is the 2x2 mean of the one below
    (odd sizes never occur after the resize), the last level is 1x1; resample_weights as written gives four equal taps, so the
    resized image is a 4x4 box average: a constant image stays constant and the mean is preserved to rounding."""
    img = scenes.procedural_image(32, 3)
    lv = orc.generate_pyramid(img)
    assert len(lv) == 6 and np.array_equal(lv[0], img) and lv[-1].shape == (1, 1, 3)
    assert np.allclose(lv[-1].ravel(), img.reshape(-1, 3).mean(axis=0), rtol=1e-5)
    odd = np.full((37, 50, 1), 0.3, np.float32)
    lo = orc.generate_pyramid(odd, "clamp")
    assert [l.shape for l in lo][:2] == [(64, 64, 1), (32, 32, 1)] and np.allclose(lo[0], 0.3, rtol=1e-6)
    rng = np.random.default_rng(2)
    r = rng.random((20, 9, 3)).astype(np.float32)
    lr = orc.generate_pyramid(r, "repeat")
    assert lr[0].shape == (32, 16, 3) and (lr[0] >= 0).all() and abs(lr[0].mean() - r.mean()) < 0.03
    from shimmer_b200.host import SceneBuilder
    b = SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    t = b.image_texture(None, levels=lr)
    assert b.textures[t]["n_channels"] == 3 and len(b.textures[t]["levels"]) == 6


# ---- the non-image members of FloatTexture / SpectrumTexture (texture.rs:180-310, :537-826) --------------------------------
f32 = np.float32


def composite_texture_scene():
    """One scene holding every non-image texture kind, float and spectrum typed; shared with the GPU lookup-parity test."""
    b = SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    T = {}
    T["mono"] = b.image_texture(scenes.procedural_image(32, 1), filter="bilinear", su=2.0, sv=3.0)
    T["mono2"] = b.image_texture(scenes.procedural_image(16, 1), filter="point", wrap="clamp")
    T["rgb"] = b.image_texture(scenes.procedural_image(64, 3), filter="bilinear")
    T["c03"], T["c0"], T["c1"], T["cinf"] = b.constant_texture(0.3), b.constant_texture(0.0), b.constant_texture(1.0), b.constant_texture(float("inf"))
    T["cspec"] = b.constant_texture(spectrum=b.spectrum(named_spectrum("metal-Cu-k")))
    T["f_scaled"] = b.scaled_texture(T["mono"], T["c03"])
    T["f_mix"] = b.mix_texture(T["mono"], T["mono2"], T["c03"])
    T["f_mix_tex_amount"] = b.mix_texture(T["c03"], T["mono"], T["mono2"])
    T["f_dir"] = b.direction_mix_texture(T["mono"], T["mono2"], dir=(0.2, 0.9, -0.4))
    T["f_deep"] = b.scaled_texture(b.mix_texture(b.scaled_texture(T["mono"], T["mono2"]), T["c03"], T["mono"]), T["c03"])   # depth 3
    T["s_scaled"] = b.scaled_texture(T["rgb"], T["mono"])
    T["s_mix"] = b.mix_texture(T["rgb"], T["cspec"], T["mono2"])
    T["s_dir"] = b.direction_mix_texture(T["cspec"], T["rgb"], dir=(0.0, 1.0, 0.0))
    T["s_mono_operand"] = b.mix_texture(T["rgb"], T["mono"], T["c03"])           # a one-channel row as a spectrum operand (texture.rs:803-807)
    # short circuits: the skipped operand is +inf, so evaluating it anyway would give NaN (inf * 0)
    T["sc_scale0"] = b.scaled_texture(T["cinf"], T["c0"])
    T["sc_mix1"] = b.mix_texture(T["cinf"], T["mono"], T["c1"])
    T["sc_mix0"] = b.mix_texture(T["mono"], T["cinf"], T["c0"])
    T["sc_dir"] = b.direction_mix_texture(T["cinf"], T["mono"], dir=(0.0, 1.0, 0.0))   # skipped where dot(n, dir) == 0
    T["sc_dir1"] = b.direction_mix_texture(T["mono"], T["cinf"], dir=(0.0, 1.0, 0.0))  # skipped where dot(n, dir) == 1
    m = b.diffuse(("const", 0.5), reflectance_tex=T["s_mix"], displacement_tex=T["f_scaled"])
    b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), m)
    return b, T


def composite_texture_queries(n=2048, seed=21):
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 6), np.float32); q[:, :2] = rng.random((n, 2)) * 3.0 - 1.0
    q[:, 2:] = (rng.random((n, 4)) - 0.5) * (10.0 ** rng.uniform(-4, -1.0, (n, 1)))
    nrm = rng.standard_normal((n, 3)).astype(np.float32); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True).astype(np.float32)
    nrm[: n // 8] = [1.0, 0.0, 0.0]                       # dot(n, +y) == 0
    nrm[n // 8: n // 4] = [0.0, 1.0, 0.0]                 # dot(n, +y) == 1
    lam = rng.uniform(360.0, 830.0, (n, 4)).astype(np.float32)
    return q, nrm, lam


def test_composite_textures_follow_the_reference_formulas():
    b, T = composite_texture_scene()
    sc = b.build()
    q, nrm, lam = composite_texture_queries()
    n = len(q)
    f = lambda name: orc.texture_eval_ctx(sc, T[name], q, nrm, lambda4=lam, as_float=True)[:, 0]
    s = lambda name: orc.texture_eval_ctx(sc, T[name], q, nrm, lambda4=lam, as_float=False)
    mono, mono2, rgb, cspec = f("mono"), f("mono2"), s("rgb"), s("cspec")
    one = f32(1.0)
    assert np.array_equal(f("c03"), np.full(n, f32(0.3))) and np.array_equal(s("c03"), np.full((n, 4), f32(0.3)))
    assert np.ptp(cspec, axis=0).max() > 0.0                                      # a real spectrum, sampled per wavelength
    assert np.array_equal(f("f_scaled"), np.where(f32(0.3) == 0, 0, mono * f32(0.3)))                          # texture.rs:206-213
    assert np.array_equal(f("f_mix"), mono * (one - f32(0.3)) + mono2 * f32(0.3))                              # :246-261
    amt = mono2
    t1 = np.where(amt != 1.0, f32(0.3), f32(0.0)); t2 = np.where(amt != 0.0, mono, f32(0.0))
    assert np.array_equal(f("f_mix_tex_amount"), t1 * (one - amt) + t2 * amt)
    d = np.array([0.2, 0.9, -0.4], np.float32)
    amt = (nrm.astype(np.float64) @ d.astype(np.float64)).astype(np.float32)      # dot3 is fma-based: compare with a tolerance
    assert np.allclose(f("f_dir"), amt * mono + (one - amt) * mono2, rtol=1e-5, atol=1e-6)                     # :295-310 (amt * t1: reversed w.r.t. Mix)
    a_in = np.where(mono2 == 0, f32(0.0), mono * mono2)                          # scaled(mono, mono2)
    inner = np.where(mono != 1, a_in, f32(0.0)) * (one - mono) + np.where(mono != 0, f32(0.3), f32(0.0)) * mono   # mix(., 0.3, amount = mono)
    assert np.array_equal(f("f_deep"), inner * f32(0.3))
    assert np.array_equal(s("s_scaled"), np.where(mono[:, None] == 0, 0, rgb * mono[:, None]))                 # :567-583
    a2 = mono2[:, None]
    assert np.array_equal(s("s_mix"), np.where(a2 != 1, rgb, 0) * (one - a2) + np.where(a2 != 0, cspec, 0) * a2)   # :631-651
    ay = nrm[:, 1:2]                                                               # dot(n, (0,1,0)) == n.y exactly
    assert np.array_equal(s("s_dir"), ay * np.where(ay != 0, cspec, 0) + (one - ay) * np.where(ay != 1, rgb, 0))    # :810-826
    assert np.array_equal(s("s_mono_operand"), rgb * (one - f32(0.3)) + mono[:, None] * f32(0.3))
    # short circuits exactly as written: the +inf operand is never touched where the reference skips it
    assert np.array_equal(f("sc_scale0"), np.zeros(n, np.float32))
    assert np.array_equal(f("sc_mix1"), mono) and np.array_equal(f("sc_mix0"), mono)
    perp, par = slice(0, n // 8), slice(n // 8, n // 4)
    assert np.array_equal(f("sc_dir")[perp], mono[perp]) and np.isinf(f("sc_dir")[par]).all()
    assert np.array_equal(f("sc_dir1")[par], mono[par]) and np.isinf(f("sc_dir1")[perp]).all()


def test_composite_texture_scene_renders_and_bumps():
    """The textree scene uses a direction-mix reflectance and a scaled displacement: both change the film."""
    base = scenes.tiny_scene("textree", resolution=(16, 16))
    sc = base.build()
    film, st, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    assert np.isfinite(film).all() and film[:, :3].sum() > 0
    flat = scenes.tiny_scene("textree", resolution=(16, 16))
    for m in flat.materials:
        m["tex_displacement"] = -1
    film2, _, _ = orc.render(flat.build(), orc.make_params(seed=5, spp=4))
    assert not np.array_equal(film, film2)


def _param_scene(kind, via_textures):
    """One sphere + ground + light; the sphere material's parameters given as plain constants or as constant textures."""
    b = SceneBuilder(); b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=(16, 16))
    cu_eta, cu_k, au_eta, au_k = (named_spectrum(n) for n in ("metal-Cu-eta", "metal-Cu-k", "metal-Au-eta", "metal-Au-k"))
    ct = lambda v: b.constant_texture(v)
    cs = lambda spec: b.constant_texture(spectrum=b.spectrum(spec))
    if kind == "conductor":
        mat = b.conductor(*((au_eta, au_k) if not via_textures else (cu_eta, cu_k)), roughness=0.25 if not via_textures else 0.9)
        if via_textures:
            b.set_material_textures(mat, u_roughness=ct(0.25), v_roughness=ct(0.25), spec_a=cs(au_eta), spec_b=cs(au_k))
    elif kind == "dielectric":
        mat = b.dielectric(("const", 1.5), roughness=0.3 if not via_textures else 0.0)
        if via_textures:
            b.set_material_textures(mat, u_roughness=ct(0.3), v_roughness=ct(0.3))
    elif kind == "coated":
        args = dict(roughness=0.2, thickness=0.05, albedo=("const", 0.3), g=-0.4) if not via_textures else dict(roughness=0.7, thickness=0.5, albedo=("const", 0.0), g=0.0)
        mat = b.coated_diffuse(scenes._red(), **args)
        if via_textures:
            b.set_material_textures(mat, u_roughness=ct(0.2), v_roughness=ct(0.2), thickness=ct(0.05), spec_b=ct(0.3), g=ct(-0.4))
    else:
        args = dict(conductor_eta=au_eta, conductor_k=au_k, interface_roughness=0.1, conductor_roughness=0.3, thickness=0.02, albedo=("const", 0.2), g=0.3, remap=False)
        if via_textures:
            args = dict(conductor_eta=cu_eta, conductor_k=cu_k, interface_roughness=0.6, conductor_roughness=0.0, thickness=0.3, albedo=("const", 0.0), g=0.0, remap=False)
        mat = b.coated_conductor(**args)
        if via_textures:
            b.set_material_textures(mat, u_roughness=ct(0.1), v_roughness=ct(0.1), u_roughness2=ct(0.3), v_roughness2=ct(0.3), thickness=ct(0.02),
                                    spec_b=ct(0.2), g=ct(0.3), spec_a=cs(au_eta), spec_d=cs(au_k))
    P, I, Nn, UV = scenes.uv_sphere(10, 14, center=(0.0, 0.6, 0.0), radius=0.6)
    b.add_mesh(P, I, mat, n=Nn, uv=UV)
    gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, b.diffuse(scenes._white()))
    lp, li = scenes._quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
    b.add_mesh(lp, li, b.diffuse(scenes._white()), area_light=dict(L=named_spectrum("stdillum-D65"), scale=30.0, two_sided=False))
    return b.build()


@pytest.mark.parametrize("kind", ["conductor", "dielectric", "coated", "coatedconductor"])
def test_constant_parameter_textures_equal_the_plain_constants(kind):
    """SgMaterialTextures: a *ConstantTexture given through the texture table renders what the same constant in SgMaterial renders
    (the reference makes no distinction: both are `tex_eval.evaluate_*` of a constant texture, material.rs:456-499, 603-635, 917-963,
    1188-1260).  The texture variant starts from deliberately different SgMaterial values, so every override must take effect."""
    plain, _, _ = orc.render(_param_scene(kind, False), orc.make_params(seed=7, spp=8))
    textured, st, _ = orc.render(_param_scene(kind, True), orc.make_params(seed=7, spp=8))
    assert np.isfinite(plain).all() and plain[:, :3].sum() > 0
    assert np.allclose(textured, plain, rtol=1e-6, atol=1e-9)
