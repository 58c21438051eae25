"""GPU parity for SURVEY 8f next-4: image-infinite lights, CoatedConductor, normal maps, spherical / cylindrical / planar
texture mappings, MixMaterial and the SimplePath / RandomWalk integrators -- the CUDA path against the CPU oracle on the
same (pixel, sample) random streams, through the C ABI."""
import json
import os

import numpy as np
import pytest

import orc
from shimmer_b200 import Options, create_integrator, ffi, scenes
from shimmer_b200.host import SceneBuilder, Transform, named_spectrum as host_named
from test_gpu_parity import _film_close

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("kind", list(scenes.VARIETY_KINDS))
def test_variety_scene_films(kind):
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 4, "seed": 5})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    _film_close(film, ref, frac=1.0, rtol=1e-4)        # every pixel (tools/film_agreement.py: all variety kinds reach it on the B200)
    gold = json.load(open(os.path.join(GOLDEN, "tiny_films.json")))[kind]
    assert abs(int(integ.stats.closest_hit_rays) - gold["closest_hit_rays"]) <= 2
    assert abs(int(integ.stats.shadow_rays) - int(rst.shadow_rays)) <= 2
    assert np.allclose(film.sum(axis=0), gold["film_sum"], rtol=5e-3)
    integ.close()


@pytest.mark.parametrize("kind", ["envmap", "mix", "coatedcond"])
def test_variety_scene_films_more_samples(kind):
    """32x32 pixels x 16 spp: enough paths that every branch of the new code (compensated-distribution sampling, MIS
    against escaped rays, nested layered walks, the mix resolver's queue appends) runs thousands of times."""
    sc = scenes.tiny_scene(kind, resolution=(32, 32)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 16, "seed": 1})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=1, spp=16))
    _film_close(film, ref, frac=0.995)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 1e-4 * rst.closest_hit_rays + 2
    img_g = integ.develop(film).reshape(-1, 3); img_r = orc.develop(sc, ref)
    assert np.sqrt(np.mean((img_g - img_r) ** 2)) / np.mean(img_r) < 0.01
    integ.close()


INTEGRATOR_CASES = [("simplepath", True, True), ("simplepath", False, True), ("simplepath", True, False), ("simplepath", False, False),
                    ("randomwalk", True, True)]


@pytest.mark.parametrize("integ_name,sl,sb", INTEGRATOR_CASES)
@pytest.mark.parametrize("kind", ["cornell", "mirror", "envmap", "glass", "texewa", "instfix", "coated", "mixtex", "spherelight", "patchlightbent"])
def test_simplepath_and_randomwalk_film_parity(kind, integ_name, sl, sb):
    """SimplePathIntegrator / RandomWalkIntegrator (integrator.rs:458-728) on the device vs the oracle, same random streams."""
    sc = (scenes.cornell_box(resolution=(16, 16)) if kind == "cornell" else scenes.tiny_scene(kind, resolution=(16, 16))).build()
    integ = create_integrator("wavefront", {"maxdepth": 4, "integrator": integ_name, "samplelights": sl, "samplebsdf": sb}, sc,
                              {"pixelsamples": 4, "seed": 3})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=3, spp=4, max_depth=4, integrator=integ_name, sample_lights=sl, sample_bsdf=sb))
    _film_close(film, ref, frac=0.99)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 2
    assert abs(int(integ.stats.shadow_rays) - int(rst.shadow_rays)) <= 2
    if integ_name == "randomwalk" or not sl:
        assert integ.stats.shadow_rays == 0
    integ.close()


def test_mapped_texture_lookup_parity():
    """sg_texture_eval_p vs the oracle: spherical / cylindrical / planar mappings (texture.rs:938-1035) with random positions
    and footprints, every filter."""
    rgb_img, mono = scenes.procedural_image(64, 3), scenes.procedural_image(32, 1)
    b = SceneBuilder(); b.set_camera((0, 0.5, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    maps = [b.texture_mapping("spherical", texture_from_world=Transform.translate((0.1, -0.2, 0.3))),
            b.texture_mapping("cylindrical", texture_from_world=Transform.rotate(40.0, (1, 0.2, 0)) * Transform.scale(0.5, 0.5, 2.0)),
            b.texture_mapping("planar", v1=(0.7, 0.0, 0.2), v2=(0.0, 0.3, 0.9), udelta=0.1, vdelta=0.25)]
    ids = []
    for mp in maps:
        for filt in ("point", "bilinear", "trilinear", "ewa"):
            ids.append((b.image_texture(rgb_img, filter=filt, mapping=mp, scale=0.9), False))
            ids.append((b.image_texture(mono, filter=filt, wrap="clamp", mapping=mp), True))
    m = b.diffuse(("const", 0.5), reflectance_tex=ids[0][0])
    b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), m)
    sc = b.build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 1})
    rng = np.random.default_rng(8)
    n = 4096
    p = (rng.standard_normal((n, 3)) * 2.0).astype(np.float32)
    mag = (10.0 ** rng.uniform(-4, -0.5, (n, 1))).astype(np.float32)
    dpdx = (rng.standard_normal((n, 3)) * mag).astype(np.float32); dpdy = (rng.standard_normal((n, 3)) * mag).astype(np.float32)
    dpdx[: n // 8] = 0.0; dpdy[: n // 8] = 0.0
    lam = rng.uniform(360.0, 830.0, (n, 4)).astype(np.float32)
    for tex, as_float in ids:
        got = integ.texture_eval_p(tex, p, dpdx=dpdx, dpdy=dpdy, lambda4=lam, as_float=as_float)
        exp = orc.texture_eval_p(sc, tex, p, dpdx=dpdx, dpdy=dpdy, lambda4=lam, as_float=as_float)
        close = np.isclose(got, exp, rtol=5e-5, atol=5e-6).all(axis=1)
        assert close.mean() > 0.99, (tex, as_float, close.mean())      # asin / atan2 / log2 come from different libms: texel-boundary flips
    integ.close()


def test_environment_light_converged_image_matches_reference_rng_mode():
    """Different random numbers on both sides (the oracle in the reference's sequential-RNG mode): mean radiance of the
    environment-lit scene agrees to 1 %."""
    sc = scenes.tiny_scene("envonly", resolution=(16, 16)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 4096, "seed": 11})
    film = integ.render(Options()).copy()
    ref, _, _ = orc.render(sc, orc.make_params(seed=0, spp=1024), stream_mode=1)
    mg = film[:, :3].sum() / film[:, 3].sum(); mr = ref[:, :3].sum() / ref[:, 3].sum()
    assert abs(mg - mr) / mr < 0.01, (mg, mr)
    integ.close()


def test_invalid_variety_inputs_are_rejected():
    from shimmer_b200 import ffi, ShimmerGpuError
    sc = scenes.tiny_scene("mix", resolution=(8, 8)).build()
    mats = sc.arrays["materials"]
    mix_ids = [i for i in range(sc.desc.n_materials) if mats[i].kind == ffi.SG_MATERIAL_MIX]
    mats[mix_ids[0]].mix_materials[0] = mix_ids[0]                 # a cycle
    with pytest.raises(ShimmerGpuError, match="cyclic|nesting"):
        create_integrator("wavefront", {}, sc)
    sc = scenes.tiny_scene("envmap", resolution=(8, 8)).build()
    sc.arrays["env_maps"][0].res = 1 << 20                         # image would lie outside the texel pool
    with pytest.raises(ShimmerGpuError, match="texel pool"):
        create_integrator("wavefront", {}, sc)
    sc = scenes.tiny_scene("diffuse", resolution=(8, 8)).build()
    with pytest.raises(ShimmerGpuError, match="Unknown integrator"):
        create_integrator("wavefront", {"integrator": "bdpt"}, sc)
    sc = scenes.tiny_scene("diffuse", resolution=(8, 8)).build()
    spectra = sc.arrays["spectra"]
    pl = [i for i in range(sc.desc.n_spectra) if spectra[i].kind == ffi.SG_SPECTRUM_PIECEWISE_LINEAR]
    spectra[pl[0]].off_b = sc.desc.n_pool - 1                      # the values of a piecewise-linear spectrum would run past spectrum_pool
    with pytest.raises(ShimmerGpuError, match="spectrum_pool"):
        create_integrator("wavefront", {}, sc)


@pytest.mark.parametrize("shape,wrap", [((64, 64, 3), "repeat"), ((32, 128, 1), "clamp"), ((37, 50, 3), "repeat"), ((100, 37, 1), "clamp"),
                                        ((5, 3, 3), "repeat"), ((1, 1, 3), "repeat"), ((300, 500, 3), "clamp")])
def test_device_mip_pyramid_matches_generate_pyramid(shape, wrap):
    """sg_image_generate_pyramid vs the oracle's Image::generate_pyramid (image.rs:699-787) + float_resize_up (:1007-1111).  The 2x2
    box-filter levels are pure IEEE adds: bit-exact given the same level 0; the resize weights go through sinf (different libms)."""
    from shimmer_b200 import generate_pyramid
    rng = np.random.default_rng(7)
    img = rng.random(shape).astype(np.float32) ** 2
    got = generate_pyramid(img, wrap)
    exp = orc.generate_pyramid(img, wrap)
    assert [g.shape for g in got] == [e.shape for e in exp] and got[-1].shape[:2] == (1, 1)
    pow2 = all((s & (s - 1)) == 0 for s in shape[:2])
    if pow2:
        assert all(np.array_equal(g, e) for g, e in zip(got, exp))
        assert np.array_equal(got[0], img.reshape(got[0].shape))
    else:
        for g, e in zip(got, exp):
            assert np.allclose(g, e, rtol=2e-6, atol=2e-7)
        # the box-filter chain itself is exact: rebuild it from the device's own level 0
        for l in range(len(got) - 1):
            a = got[l]; h, w = a.shape[:2]
            y0 = np.arange(0, h, 2); x0 = np.arange(0, w, 2); y1 = y0 + (1 if h > 1 else 0); x1 = x0 + (1 if w > 1 else 0)
            nxt = np.float32(0.25) * (((a[y0][:, x0] + a[y0][:, x1]) + a[y1][:, x0]) + a[y1][:, x1])
            assert np.array_equal(nxt.astype(np.float32), got[l + 1])


def test_device_pyramid_feeds_a_texture_and_rejects_what_the_reference_asserts_on():
    from shimmer_b200 import generate_pyramid, ShimmerGpuError
    img = scenes.procedural_image(64, 3)[:50, :37].copy()                   # 37 x 50: not a power of two
    levels = generate_pyramid(img, "repeat")
    b = SceneBuilder(); b.set_camera((0, 1.0, -3), (0, 0.5, 0), (0, 1, 0), 45.0, (16, 16))
    tex = b.image_texture(None, filter="trilinear", levels=levels)
    gp, gi = scenes._quad((-3, 0.0, -3), (-3, 0.0, 3), (3, 0.0, 3), (3, 0.0, -3))
    b.add_mesh(gp, gi, b.diffuse(scenes._white(), reflectance_tex=tex), uv=np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.float32))
    lp, li = scenes._quad((-0.5, 2.5, -0.5), (0.5, 2.5, -0.5), (0.5, 2.5, 0.5), (-0.5, 2.5, 0.5))
    b.add_mesh(lp, li, b.diffuse(scenes._white()), area_light=dict(L=host_named("stdillum-D65"), scale=30.0, two_sided=False))
    sc = b.build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 4, "seed": 2})
    film = integ.render(Options()).copy()
    ref, _, _ = orc.render(sc, orc.make_params(seed=2, spp=4))
    _film_close(film, ref, frac=0.99)
    integ.close()
    with pytest.raises(ShimmerGpuError, match="both dimensions"):
        generate_pyramid(np.zeros((50, 64, 3), np.float32))                 # 64 is already a power of two: image.rs:1009 asserts
    with pytest.raises(ShimmerGpuError, match="repeat and clamp"):
        generate_pyramid(np.zeros((5, 3, 3), np.float32), "black")


def test_composite_texture_lookup_parity():
    """sg_texture_eval_ctx vs the oracle for the non-image textures (texture.rs:180-310,:537-826): constant, scaled, mix and
    direction-mix, float and spectrum typed, nested three deep, including the short circuits that keep a +inf operand from
    turning the result into NaN.  Leaves are point / bilinear images, whose lookups are exact f32 arithmetic up to the level
    choice (log2f), hence the small allowance; everything above the leaves must then agree bit for bit."""
    from test_oracle_variety import composite_texture_scene, composite_texture_queries
    b, T = composite_texture_scene()
    sc = b.build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 1})
    q, nrm, lam = composite_texture_queries(n=4096)
    for name, tex in T.items():
        for as_float in ([True] if sc.arrays["textures"][tex].n_channels == 1 else []) + [False]:
            got = integ.texture_eval_ctx(tex, q, nrm, lambda4=lam, as_float=as_float)
            exp = orc.texture_eval_ctx(sc, tex, q, nrm, lambda4=lam, as_float=as_float)
            same = ((got == exp) | (np.isnan(got) & np.isnan(exp))).all(axis=1)
            close = same | np.isclose(got, exp, rtol=2e-5, atol=2e-6).all(axis=1)
            assert close.mean() > 0.998 and same.mean() > 0.98, (name, as_float, same.mean(), close.mean())
            if name.startswith("c") or name.startswith("sc_scale0"):
                assert same.all(), name
    integ.close()


def test_malformed_texture_trees_are_rejected():
    from shimmer_b200 import ffi, ShimmerGpuError
    def scene_with(build):
        b = SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
        mono = b.image_texture(scenes.procedural_image(8, 1)); rgb = b.image_texture(scenes.procedural_image(8, 3))
        t = build(b, mono, rgb)
        m = b.diffuse(("const", 0.5), reflectance_tex=t)
        b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), m)
        return b.build()
    def too_deep(b, mono, rgb):
        t = mono
        for _ in range(ffi.SG_MAX_TEXTURE_DEPTH + 1):
            t = b.scaled_texture(t, mono)
        return t
    with pytest.raises(ShimmerGpuError, match="SG_MAX_TEXTURE_DEPTH"):
        create_integrator("wavefront", {}, scene_with(too_deep))
    def ok_depth(b, mono, rgb):
        t = rgb
        for _ in range(ffi.SG_MAX_TEXTURE_DEPTH):
            t = b.scaled_texture(t, mono)
        return t
    create_integrator("wavefront", {}, scene_with(ok_depth)).close()
    sc = scene_with(lambda b, mono, rgb: b.scaled_texture(rgb, mono))
    sc.arrays["texture_nodes"][0].tex1 = sc.desc.n_textures - 1                # its own row: a cycle
    with pytest.raises(ShimmerGpuError, match="cycle"):
        create_integrator("wavefront", {}, sc)
    sc = scene_with(lambda b, mono, rgb: b.scaled_texture(rgb, mono))
    sc.arrays["texture_nodes"][0].tex2 = 1                                      # `scale` must be a float texture
    with pytest.raises(ShimmerGpuError, match="operand"):
        create_integrator("wavefront", {}, sc)


@pytest.mark.parametrize("kind", ["conductor", "dielectric", "coated", "coatedconductor"])
def test_constant_parameter_textures_equal_the_plain_constants_on_gpu(kind):
    """SgMaterialTextures on the device: constant textures given through the table render what the same constants in SgMaterial
    render, and both match the oracle (material.rs:456-499, 603-635, 917-963, 1188-1260)."""
    from test_oracle_variety import _param_scene
    films = []
    for via in (False, True):
        sc = _param_scene(kind, via)
        integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 8, "seed": 7})
        films.append(integ.render(Options()).copy())
        integ.close()
    assert np.allclose(films[0], films[1], rtol=1e-6, atol=1e-9)
    ref, _, _ = orc.render(_param_scene(kind, True), orc.make_params(seed=7, spp=8))
    _film_close(films[1], ref, frac=0.99)


@pytest.mark.parametrize("integ_name", ["simplepath", "randomwalk"])
@pytest.mark.parametrize("max_depth", [1, 2, 3])
def test_last_vertex_leaves_the_wavelengths_alone(integ_name, max_depth):
    """SimplePath / RandomWalk return at `depth == max_depth` BEFORE get_bsdf (integrator.rs:520-523, :628-631), so a
    dispersive dielectric at the LAST vertex does not terminate the secondary wavelengths (PathIntegrator::li does: it builds the BSDF
    first).  Terminating there would rescale everything the path gathered before (the film divides L by the final wavelength pdfs):
    floor (light sample) -> BK7 sphere as the last vertex is common in this scene at small max_depth.  Every pixel must match."""
    sc = scenes.tiny_scene("glass", resolution=(24, 24)).build()
    integ = create_integrator("wavefront", {"maxdepth": max_depth, "integrator": integ_name}, sc, {"pixelsamples": 8, "seed": 11})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=11, spp=8, max_depth=max_depth, integrator=integ_name))
    _film_close(film, ref, frac=1.0, rtol=1e-4)
    assert int(integ.stats.closest_hit_rays) == int(rst.closest_hit_rays)
    integ.close()


def test_triangle_emitter_validation():
    """an emissive triangle must point at an SG_LIGHT_DIFFUSE_AREA light over that very triangle, and may not sit inside
    an object definition (light sampling reads the emitter's render-space vertices) -- checked at the C ABI, not only in host.py."""
    import ctypes as C
    sc = scenes.cornell_box(resolution=(8, 8)).build()
    lib = ffi.load_library(); h = C.c_void_p()
    prims = sc.arrays["prims"]
    emissive = [i for i in range(len(prims)) if prims["light"][i] >= 0]
    assert len(emissive) >= 2
    # (1) two emitters swap their lights: each light row now describes another triangle
    bad = prims.copy(); bad["light"][emissive[0]], bad["light"][emissive[1]] = prims["light"][emissive[1]], prims["light"][emissive[0]]
    d = ffi.SgSceneDesc.from_buffer_copy(sc.desc); d.primitives = bad.ctypes.data_as(C.POINTER(ffi.SgPrimitive))
    assert lib.sg_scene_create(C.byref(d), C.byref(h)) == -1 and b"emissive triangle" in lib.sg_last_error()
    # (2) the light row is of another kind
    lights = np.frombuffer(bytes(C.string_at(sc.desc.lights, sc.desc.n_lights * C.sizeof(ffi.SgLight))), dtype=np.dtype(ffi.SgLight)).copy()
    lights["kind"][prims["light"][emissive[0]]] = ffi.SG_LIGHT_POINT
    d = ffi.SgSceneDesc.from_buffer_copy(sc.desc); d.lights = lights.ctypes.data_as(C.POINTER(ffi.SgLight))
    assert lib.sg_scene_create(C.byref(d), C.byref(h)) == -1 and b"emissive triangle" in lib.sg_last_error()
    # (3) an emitter inside an object definition
    inst = scenes.tiny_scene("inst", resolution=(8, 8)).build()
    ip = inst.arrays["prims"].copy()
    n_top = inst.desc.n_top_primitives
    assert n_top and n_top < len(ip)
    ip["light"][n_top] = 0
    d = ffi.SgSceneDesc.from_buffer_copy(inst.desc); d.primitives = ip.ctypes.data_as(C.POINTER(ffi.SgPrimitive))
    assert lib.sg_scene_create(C.byref(d), C.byref(h)) == -4 and b"object definitions" in lib.sg_last_error()     # SG_ERR_UNSUPPORTED
    # (4) null table with a non-zero count is an error, not a crash
    d = ffi.SgSceneDesc.from_buffer_copy(sc.desc); d.materials = None
    assert lib.sg_scene_create(C.byref(d), C.byref(h)) == -1


def test_sg_init_can_be_called_again():
    """sg_init twice (same device) keeps working; scenes created before and after both render."""
    lib = ffi.load_library()
    sc = scenes.cornell_box(resolution=(8, 8)).build()
    a = create_integrator("wavefront", {}, sc, {"pixelsamples": 2})
    fa = a.render(Options(seed=2)).copy()
    assert lib.sg_init(0) == 0 and lib.sg_device_count() == 1
    b = create_integrator("wavefront", {}, sc, {"pixelsamples": 2})
    fb = b.render(Options(seed=2)).copy()
    a.film[:] = 0
    assert np.array_equal(fa, fb) and np.array_equal(a.render(Options(seed=2)), fa)
    a.close(); b.close()
