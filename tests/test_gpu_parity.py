"""GPU parity tests proper: every call goes through the C ABI (libshimmer_gpu.so) and is compared with the
CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star):
  * bit-exact: RNG streams; first-hit primitive index and hit/miss on the ray-cast sets
  * hit t, barycentrics, geometric normal: <= 1e-5 relative (observed: bit-exact)
  * films: per-sample deterministic streams are shared, so films agree to float round-off except for the
    rare path whose control flow flips on a libm last-bit difference (atanh/cosh/sin/cos/atan2/asin)
"""
import json
import os

import numpy as np
import pytest

import orc
from shimmer_b200 import Options, create_integrator, ffi, scenes
from shimmer_b200.integrator import sampler_fill

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def cornell_gpu(cornell64):
    integ = create_integrator("wavefront", {"maxdepth": 5}, cornell64, {"pixelsamples": 16})
    yield integ
    integ.close()


def test_native_library_is_the_one_running():
    import ctypes
    lib = ffi.load_library()
    assert isinstance(lib, ctypes.CDLL) and os.path.samefile(lib._name, ffi.LIB_PATH)


# ---- bit-exact: sampler / RNG ---------------------------------------------------------------------
def test_rng_stream_bit_exact():
    ref = np.zeros(4096, np.float32)
    orc.lib().orc_sampler_fill(0, 1, 0, 0, len(ref), ref.ctypes.data)
    got = sampler_fill(0, len(ref), raw=True)
    assert got.tobytes() == ref.tobytes()
    assert got[:4].tolist() == [0.3245752453804016, 0.38223928213119507, 0.3596171736717224, 0.0114554762840271]
    for seed, pix, smp in ((0, 0, 0), (7, 123456, 3), (2 ** 63 + 5, 4095 * 4096, 1023)):
        orc.lib().orc_sampler_fill(seed, 0, pix, smp, 256, ref.ctypes.data)
        assert sampler_fill(seed, 256, pix, smp).tobytes() == ref[:256].tobytes()


def test_camera_rays_and_wavelengths(cornell_gpu, cornell64):
    rng = np.random.default_rng(0)
    n = 2000
    xy = rng.integers(0, 64, size=(n, 2)).astype(np.int32); si = rng.integers(0, 64, size=n).astype(np.int32)
    opts = Options(seed=9)
    rays, lam = cornell_gpu.camera_rays(opts, xy, si)
    p = orc.make_params(seed=9, spp=16)
    r_ref, l_ref = orc.camera_rays(cornell64, p, xy, si)
    assert np.array_equal(rays, r_ref)                      # pure IEEE arithmetic: bit-exact
    assert np.allclose(lam, l_ref, rtol=2e-6, atol=0)       # atanh / cosh come from different libms
    opts2 = Options(seed=9, disable_pixel_jitter=True, disable_wavelength_jitter=True)
    rays2, lam2 = cornell_gpu.camera_rays(opts2, xy, si)
    p2 = orc.make_params(seed=9, spp=16, flags=ffi.SG_OPT_DISABLE_PIXEL_JITTER | ffi.SG_OPT_DISABLE_WAVELENGTH_JITTER)
    r2, l2 = orc.camera_rays(cornell64, p2, xy, si)
    assert np.array_equal(rays2, r2) and np.allclose(lam2, l2, rtol=2e-6)


# ---- bit-exact: first hit ------------------------------------------------------------------------------
def _ray_set(sc, n, seed):
    """camera-like rays + uniformly random origin/direction pairs inside the scene bounds + adversarial rays
    (axis-parallel directions with zero components -> +-inf inv_dir, rays through shared edges/vertices)."""
    rng = np.random.default_rng(seed)
    root = sc.arrays["nodes"][0]
    lo, hi = np.array(root["bmin"]), np.array(root["bmax"])
    o = (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)
    d = rng.standard_normal((n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 8
    axes = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1.0, 1.0], (k, 1)).astype(np.float32)
    d[:k] = axes
    # aim a batch of rays exactly at mesh vertices and edge midpoints
    P = sc.arrays["p"]; I = sc.arrays["idx"]
    tri = I[rng.integers(0, len(I), k)]
    base = np.array([m.first_vertex for m in sc.arrays["meshes"]])
    mesh_of_tri = np.searchsorted(np.cumsum([m.n_triangles for m in sc.arrays["meshes"]]), rng.integers(0, len(I), k), side="right")
    tri_ids = rng.integers(0, len(I), k)
    first_tri = np.concatenate([[0], np.cumsum([m.n_triangles for m in sc.arrays["meshes"]])])
    mesh_of_tri = np.searchsorted(first_tri, tri_ids, side="right") - 1
    v = I[tri_ids] + base[mesh_of_tri][:, None]
    target = np.where(rng.random((k, 1)) < 0.5, P[v[:, 0]], (P[v[:, 0]] + P[v[:, 1]]) * np.float32(0.5)).astype(np.float32)
    d[k:2 * k] = target - o[k:2 * k]
    return o, d.astype(np.float32)


def _assert_hits_equal(got, ref):
    assert np.array_equal(got["prim"], ref["prim"])                       # bit-exact hit/miss + primitive index
    hit = ref["prim"] >= 0
    for f in ("t", "b0", "b1", "b2"):
        assert np.allclose(got[f][hit], ref[f][hit], rtol=1e-5, atol=0)
    assert np.allclose(got["ng"][hit], ref["ng"][hit], rtol=1e-5, atol=1e-7)
    exact = sum(np.array_equal(got[f], ref[f]) for f in ("t", "b0", "b1", "b2"))
    return exact


def test_raycast_golden_fixture(cornell_gpu):
    g = np.load(os.path.join(GOLDEN, "cornell_raycast.npz"))
    got, st = cornell_gpu.trace(g["o"], g["d"], np.full(len(g["o"]), np.inf, np.float32), want_stats=True)
    assert np.array_equal(got["prim"], g["prim"])
    hit = g["prim"] >= 0
    assert np.array_equal(got["t"][hit], g["t"][hit])
    assert np.array_equal(np.stack([got["b0"], got["b1"], got["b2"]], 1)[hit], g["b"][hit])
    assert np.array_equal(got["ng"][hit], g["ng"][hit])
    # traversal order is the reference's: identical visit counts
    assert st.nodes_visited == int(g["nodes"]) and st.tris_tested == int(g["tris"])


@pytest.mark.parametrize("name", ["cornell", "tiny", "mesh", "inst", "instrot", "instfix"])
def test_raycast_parity_closest_and_any(name, cornell64):
    sc = {"cornell": lambda: cornell64, "tiny": lambda: scenes.tiny_scene("glass").build(),
          "mesh": lambda: scenes.mesh_scene(n_theta=120, n_phi=120, resolution=(32, 32)).build(),
          "inst": lambda: scenes.tiny_scene("inst").build(), "instrot": lambda: scenes.tiny_scene("instrot").build(),
          "instfix": lambda: scenes.tiny_scene("instfix").build()}[name]()
    integ = create_integrator("wavefront", {}, sc)
    n = 1 << 17
    o, d = _ray_set(sc, n, seed=1)
    tmax = np.full(n, np.inf, np.float32)
    got, gst = integ.trace(o, d, tmax, want_stats=True)
    ref, rst = orc.trace(sc, o, d, tmax)
    assert _assert_hits_equal(got, ref) == 4                                # observed: all four fields bit-exact
    assert gst.nodes_visited == rst.nodes_visited and gst.tris_tested == rst.tris_tested
    assert (ref["prim"] >= 0).mean() > 0.3
    # any-hit with the shadow-ray convention: unnormalised direction, t_max = 1 - 1e-4
    d2 = (d * np.float32(3.0)).astype(np.float32)
    t2 = np.full(n, np.float32(0.9999), np.float32)
    got2 = integ.trace(o, d2, t2, any_hit=True)
    ref2, _ = orc.trace(sc, o, d2, t2, any_hit=True)
    assert np.array_equal(got2["prim"], ref2["prim"])
    integ.close()


def test_trace_edge_cases(cornell_gpu):
    e = np.zeros((0, 3), np.float32)
    assert len(cornell_gpu.trace(e, e, np.zeros(0, np.float32))) == 0      # empty batch
    o = np.array([[0, 0, 1080]] * 3, np.float32)
    d = np.array([[0, 0, 1], [0, 0, 1], [np.nan, 0, 1]], np.float32)
    t = np.array([np.inf, 1e-3, np.inf], np.float32)
    got = cornell_gpu.trace(o, d, t)
    assert got["prim"][0] >= 0 and got["prim"][1] == -1                     # t_max clips the hit


# ---- films ---------------------------------------------------------------------------------------------
def _film_close(got, ref, frac=0.999, rtol=2e-3):
    assert np.array_equal(got[:, 3], ref[:, 3])                             # weight sums are exact
    lum_g, lum_r = got[:, :3].sum(axis=1), ref[:, :3].sum(axis=1)
    scale = max(lum_r.mean(), 1e-12)
    ok = np.abs(lum_g - lum_r) <= rtol * np.maximum(lum_r, 0.05 * scale)
    assert ok.mean() >= frac, f"only {ok.mean():.5f} of pixels within {rtol}"
    assert abs(lum_g.sum() - lum_r.sum()) / lum_r.sum() < 2e-3
    return ok.mean()


def test_cornell_film_and_ray_counts(cornell_gpu, cornell64):
    cornell_gpu.film[:] = 0
    film = cornell_gpu.render(Options(seed=0)).copy()
    ref, rst, _ = orc.render(cornell64, orc.make_params(seed=0, spp=16))
    _film_close(film, ref)
    gst = cornell_gpu.stats
    assert gst.camera_paths == rst.camera_paths == 64 * 64 * 16
    assert abs(int(gst.closest_hit_rays) - int(rst.closest_hit_rays)) <= 1e-4 * rst.closest_hit_rays
    assert abs(int(gst.shadow_rays) - int(rst.shadow_rays)) <= 1e-4 * rst.shadow_rays
    assert gst.kernel_launches > 0
    # developed image: RMSE and mean relative luminance error <= 1 % (north-star image bar)
    img_g = cornell_gpu.develop(film).reshape(-1, 3); img_r = orc.develop(cornell64, ref)
    rmse = np.sqrt(np.mean((img_g - img_r) ** 2)) / np.mean(img_r)
    assert rmse < 0.01


@pytest.mark.parametrize("kind", ["diffuse", "conductor", "mirror", "glass", "roughglass", "coated", "coatedrough", "ortho", "thinglass"] + list(scenes.TEXTURED_KINDS)
                         + list(scenes.INSTANCED_KINDS))
def test_tiny_scene_films(kind):
    """Round 2 bar (VERDICT r01 "film bars are loose"): EVERY pixel within 1e-4 relative -- what tools/film_agreement.py measures on
    the B200 for all of these kinds (remaining differences are float round-off of libm calls; > 99 % of pixels agree to 1e-5)."""
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 4, "seed": 5})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    _film_close(film, ref, frac=1.0, rtol=1e-4)
    gold = json.load(open(os.path.join(GOLDEN, "tiny_films.json")))[kind]
    assert abs(int(integ.stats.closest_hit_rays) - gold["closest_hit_rays"]) <= 2
    assert np.allclose(film.sum(axis=0), gold["film_sum"], rtol=5e-3)
    integ.close()


def test_texture_lookup_parity():
    """sg_texture_eval vs the oracle for every filter / wrap mode, RGB and one-channel, random footprints (incl. zero
    and strongly anisotropic ones).  Filtering is pure f32 arithmetic; only log2f (level selection) may differ in the
    last bit, so a handful of lookups that sit exactly on a level boundary are allowed to disagree."""
    rgb_img, mono = scenes.procedural_image(64, 3), scenes.procedural_image(32, 1)
    b = scenes.SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    ids = []
    for filt in ("point", "bilinear", "trilinear", "ewa"):
        for wrap in ("repeat", "clamp", "black"):
            ids.append((b.image_texture(rgb_img, filter=filt, wrap=wrap, su=1.7, sv=0.8, du=0.1, dv=-0.2, scale=0.9,
                                        spectrum_type="unbounded" if wrap == "clamp" else "albedo", invert=(wrap == "black")), False))
            ids.append((b.image_texture(mono, filter=filt, wrap=wrap, max_anisotropy=4.0), True))
    m = b.diffuse(("const", 0.5), reflectance_tex=ids[0][0])
    b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2]], np.uint32), m)
    sc = b.build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 1})
    rng = np.random.default_rng(4)
    n = 4096
    q = np.zeros((n, 6), np.float32)
    q[:, :2] = rng.random((n, 2)) * 4.0 - 1.5
    q[:, 2:] = (rng.random((n, 4)) - 0.5) * (10.0 ** rng.uniform(-4, -0.3, (n, 1)))
    q[: n // 8, 2:] = 0.0                                                   # zero footprint
    q[n // 8: n // 4, 4:] *= 0.02                                           # anisotropic
    lam = rng.uniform(360.0, 830.0, (n, 4)).astype(np.float32)
    for tex, as_float in ids:
        got = integ.texture_eval(tex, q, lam, as_float=as_float)
        exp = orc.texture_eval(sc, tex, q, lam, as_float=as_float)
        close = np.isclose(got, exp, rtol=2e-5, atol=2e-6).all(axis=1)
        assert close.mean() > 0.998, (tex, as_float, close.mean())
        assert np.isfinite(got[close]).all()
    integ.close()


def test_sample_range_split_and_batching(cornell_gpu, cornell64):
    """(a) sample ranges add up (the multi-GPU decomposition); (b) the result does not depend on the
    wavefront width (ragged last batch, tiny batches)."""
    opts = Options(seed=2, pixel_samples=6)
    cornell_gpu.film[:] = 0                      # sg_render ADDS into the caller's film
    full = cornell_gpu.render(opts).copy()
    cornell_gpu.film[:] = 0
    a = cornell_gpu.render(opts, sample_range=(0, 2)).copy(); cornell_gpu.film[:] = 0
    b = cornell_gpu.render(opts, sample_range=(2, 6)).copy(); cornell_gpu.film[:] = 0
    assert np.allclose(a + b, full, rtol=1e-9)
    small = create_integrator("wavefront", {}, cornell64, {"pixelsamples": 6}, max_paths_in_flight=1000)
    c = small.render(opts).copy()
    assert np.allclose(c, full, rtol=1e-9)
    small.close()


def test_options_and_depth_limits(cornell_gpu, cornell64):
    integ = create_integrator("wavefront", {"maxdepth": 0}, cornell64, {"pixelsamples": 2})
    f = integ.render(Options()).copy()
    assert integ.stats.shadow_rays == 0 and integ.stats.closest_hit_rays == 64 * 64 * 2
    ref, _, _ = orc.render(cornell64, orc.make_params(seed=0, spp=2, max_depth=0))
    assert np.allclose(f, ref, rtol=1e-5, atol=1e-12)
    integ.close()
    sp = create_integrator("wavefront", {"integrator": "simplepath"}, cornell64, {"pixelsamples": 2})
    with pytest.raises(ffi.ShimmerGpuError, match="path integrator only"):
        sp.render(Options(force_diffuse=True))                              # SimplePath / RandomWalk + force_diffuse: error, not fallback
    sp.close()


def test_scene_validation_errors(cornell64):
    import copy
    import ctypes as C
    lib = ffi.load_library()
    bad = ffi.SgSceneDesc.from_buffer_copy(cornell64.desc)
    bad.abi_version = 99
    h = C.c_void_p()
    assert lib.sg_scene_create(C.byref(bad), C.byref(h)) == -1 and b"abi_version" in lib.sg_last_error()
    bad = ffi.SgSceneDesc.from_buffer_copy(cornell64.desc)
    prims = cornell64.arrays["prims"].copy(); prims["material"][0] = 1000
    bad.primitives = prims.ctypes.data_as(C.POINTER(ffi.SgPrimitive))
    assert lib.sg_scene_create(C.byref(bad), C.byref(h)) == -1


def test_film_output_stage(cornell_gpu, cornell64, tmp_path):
    """sg_film_get_image vs the oracle, bit for bit: random films incl. zero weights, NaN/inf sums, values around the
    f16 overflow threshold and in the f16 subnormal range; then the PFM file layout of image.rs:1333-1377."""
    W, H = cornell_gpu.width, cornell_gpu.height
    rng = np.random.default_rng(9)
    film = np.zeros((W * H, 4))
    film[:, :3] = rng.random((W * H, 3)) * (10.0 ** rng.uniform(-9, 6, (W * H, 1)))
    film[:, 3] = rng.integers(0, 5, W * H)
    film[::97, 0] = np.nan; film[5::101, 1] = np.inf; film[7::89, 2] = -1.0
    for fp16 in (True, False):
        for flip in (False, True):
            got = cornell_gpu.get_image(film, write_fp16=fp16, bottom_up=flip)
            exp = orc.film_get_image(cornell64, film, W, H, fp16=fp16, bottom_up=flip)
            assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), (fp16, flip)
    path = tmp_path / "out.pfm"
    cornell_gpu.write_image(str(path), film)
    raw = open(path, "rb").read()
    hdr = b"PF\n%d %d\n-1\n" % (W, H)
    assert raw.startswith(hdr) and len(raw) == len(hdr) + W * H * 12
    body = np.frombuffer(raw[len(hdr):], "<f4").reshape(H, W, 3)
    assert np.array_equal(body[::-1].view(np.uint32), orc.film_get_image(cornell64, film, W, H).view(np.uint32))
    with pytest.raises(ffi.ShimmerGpuError):
        cornell_gpu.write_image(str(tmp_path / "out.png"), film)


def test_converged_image_against_reference_rng_mode():
    """North-star level 3: a CONVERGED image (131072 spp on a 12x12 window of the C1 Cornell box, below the tall box
    where direct light, shadow and colour bleeding meet) rendered by the CUDA path agrees with the oracle run in the
    reference's own RNG mode (stream_mode=1: one sequential generator per worker thread, integrator.rs:250-263 -- random
    numbers completely different from the GPU's per-(pixel, sample) streams) to <= 1 % RMSE and <= 1 % mean luminance.
    Calibration (oracle vs oracle, two RNG modes, 16x16 window): RMSE 1.28 % at 32768 spp, 0.65 % at 131072 spp."""
    win = (250, 302, 262, 314)
    sc = scenes.cornell_box(resolution=(512, 512), crop=win).build()
    spp = 131072
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp})
    film = integ.render(Options(seed=0, pixel_samples=spp)).copy()
    img_g = integ.develop(film).reshape(-1, 3)
    integ.close()
    ref, _, _ = orc.render(sc, orc.make_params(seed=0, spp=spp), stream_mode=1)
    img_r = orc.develop(sc, ref)
    rmse = np.sqrt(np.mean((img_g.astype(np.float64) - img_r) ** 2)) / np.mean(img_r)
    lum = lambda a: (0.2126 * a[:, 0] + 0.7152 * a[:, 1] + 0.0722 * a[:, 2]).mean()
    assert rmse <= 0.01, rmse
    assert abs(lum(img_g) - lum(img_r)) / lum(img_r) <= 0.01
    assert np.all(film[:, 3] == spp)


# ---- spheres ------------------------------------------------------------------------------
def _sphere_only(mults):
    from shimmer_b200.host import SceneBuilder, Transform
    b = SceneBuilder(); b.set_camera((0, 0, -20), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    m = b.diffuse(("const", 0.5))
    for mu in mults:
        b.add_sphere(1.0, m, object_from_world=Transform.translate((mu, 0, 0)))
    return b


def test_reference_bvh_sphere_vectors_on_gpu():
    """The reference's own traversal tests (aggregate.rs:604-702) through sg_trace: t = 4 on the single unit sphere,
    t = 5.5 +- 1e-5 on the set of spheres, normal -x, predicate true; the ray offset by z = 1.001 misses."""
    b = _sphere_only([0.0]); sc = b.build()
    rs = lambda p: b.render_from_world.apply_points_f32(np.array([p], np.float32))
    integ = create_integrator("wavefront", {}, sc)
    h = integ.trace(rs([-5, 0, 0]), [[1, 0, 0]], [np.inf])
    assert h["prim"][0] == 0 and abs(h["t"][0] - 4.0) <= 4 * np.spacing(np.float32(4.0)) and abs(h["b0"][0] + 1.0) <= 1e-6
    assert np.dot(h["ng"][0], [-1, 0, 0]) == 1.0
    integ.close()
    b = _sphere_only([-3.5, 0.0, 5.0]); sc = b.build()
    integ = create_integrator("wavefront", {}, sc)
    o = rs([-10, 0, 0]); d = [[1, 0, 0]]
    h = integ.trace(o, d, [np.inf])
    assert abs(h["t"][0] - 5.5) <= 1e-5 and np.dot(h["ng"][0], [-1, 0, 0]) == 1.0
    assert integ.trace(o, d, [np.inf], any_hit=True)["prim"][0] == 0
    o2 = rs([-10, 0, 1.001])
    assert integ.trace(o2, d, [np.inf])["prim"][0] == -1 and integ.trace(o2, d, [np.inf], any_hit=True)["prim"][0] == -1
    ref, _ = orc.trace(sc, o, d, [np.inf])
    assert h["t"][0] == ref["t"][0] and h["prim"][0] == ref["prim"][0]
    integ.close()


def test_sphere_raycast_parity():
    """Random + aimed rays on the sphere scene (full, scaled, clipped/rotated spheres + triangles): primitive, t and p_obj
    bit-exact (pure interval arithmetic); rays whose hit sits on a phimax clip boundary may flip on atan2f's last bit."""
    b = scenes.sphere_tiny_scene("spheres"); sc = b.build()
    integ = create_integrator("wavefront", {}, sc)
    rng = np.random.default_rng(6)
    n = 1 << 16
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.05
    centres = np.array([[-1.3, 0.6, 0.4], [0.1, 0.55, -0.6], [1.4, 0.75, 0.5]], np.float32)
    d = (centres[rng.integers(0, 3, n)] + rng.uniform(-0.8, 0.8, (n, 3)).astype(np.float32) - o).astype(np.float32)
    o[: n // 10] = centres[rng.integers(0, 3, n // 10)] + rng.uniform(-0.2, 0.2, (n // 10, 3)).astype(np.float32)
    d[n // 2:] = rng.standard_normal((n - n // 2, 3)).astype(np.float32)
    o = b.render_from_world.apply_points_f32(o)
    for tmax in (np.inf, 2.5):
        t = np.full(n, tmax, np.float32)
        got, gst = integ.trace(o, d, t, want_stats=True)
        ref, rst = orc.trace(sc, o, d, t)
        same = got["prim"] == ref["prim"]
        assert same.mean() > 0.9999
        hit = same & (ref["prim"] >= 0)
        for f in ("t", "b0", "b1", "b2"):
            assert np.array_equal(got[f][hit], ref[f][hit]), f
        assert np.allclose(got["ng"][hit], ref["ng"][hit], rtol=1e-5, atol=1e-6)
        assert (sc.arrays["prims"]["mesh"][ref["prim"][hit]] == ffi.SG_PRIM_SPHERE).mean() > 0.3
        any_g = integ.trace(o, d, np.minimum(t, np.float32(0.9999)), any_hit=True)
        any_r, _ = orc.trace(sc, o, d, np.minimum(t, np.float32(0.9999)), any_hit=True)
        assert (any_g["prim"] == any_r["prim"]).mean() > 0.9999
    integ.close()


@pytest.mark.parametrize("kind", list(scenes.INSTANCED_SHAPE_KINDS))
def test_shapes_inside_instances_raycast_parity(kind):
    """Spheres and bilinear patches inside object definitions: the two-level traversal tests them with the instance-space ray
    (inverse transform for closest hits, the reference's FORWARD transform for shadow rays unless SG_SCENE_FIX_INSTANCING).
    Primitive, instance-space t and the hit record bit-exact against the oracle."""
    b = scenes.instanced_shapes_tiny_scene(kind); sc = b.build()
    integ = create_integrator("wavefront", {}, sc)
    rng = np.random.default_rng(12)
    n = 1 << 16
    o = rng.uniform(-2.5, 2.5, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.05
    centres = np.array([[-1.2, 0.6, 0.3], [0.0, 0.7, 0.0], [1.2, 0.55, -0.2], [0.6, 1.6, 0.4], [-0.7, 1.5, -0.4]], np.float32)
    d = (centres[rng.integers(0, 5, n)] + rng.uniform(-0.6, 0.6, (n, 3)).astype(np.float32) - o).astype(np.float32)
    d[n // 2:] = rng.standard_normal((n - n // 2, 3)).astype(np.float32)
    o = b.render_from_world.apply_points_f32(o)
    t = np.full(n, np.inf, np.float32)
    got = integ.trace(o, d, t); ref, _ = orc.trace(sc, o, d, t)
    same = got["prim"] == ref["prim"]
    assert same.mean() > 0.9999
    hit = same & (ref["prim"] >= 0)
    for f in ("t", "b0", "b1", "b2"):
        assert np.array_equal(got[f][hit], ref[f][hit]), f
    kinds = sc.arrays["prims"]["mesh"][ref["prim"][hit]]
    in_object = ref["prim"][hit] >= sc.desc.n_top_primitives
    assert (in_object & (kinds == ffi.SG_PRIM_SPHERE)).mean() > 0.05 and in_object.mean() > 0.2
    any_g = integ.trace(o, d, np.full(n, 0.9999, np.float32), any_hit=True)
    any_r, _ = orc.trace(sc, o, d, np.full(n, 0.9999, np.float32), any_hit=True)
    assert (any_g["prim"] == any_r["prim"]).mean() > 0.9999
    integ.close()


@pytest.mark.parametrize("kind", list(scenes.INSTANCED_SHAPE_KINDS))
def test_shapes_inside_instances_films(kind):
    """Films of the scenes with spheres / patches inside object definitions.  Like the top-level sphere scenes they carry the
    interval-arithmetic sphere test, which is ulp-chaotic in the reference itself (tests/test_oracle_sphere.py::
    test_sphere_scenes_are_sensitive_to_one_ulp...): a per-mille of paths takes another branch than the oracle's (same rate as `spheres`:
    ~1 % of pixels at 16 spp); the first three path vertices agree everywhere (max_depth <= 2: zero differing pixels)."""
    sc = scenes.tiny_scene(kind, resolution=(24, 24)).build()
    for md, frac in ((2, 0.999), (5, 0.97)):      # instshapestex adds EWA footprints that react to last-bit uv differences
        integ = create_integrator("wavefront", {"maxdepth": md}, sc, {"pixelsamples": 8, "seed": 3})
        film = integ.render(Options()).copy()
        ref, rst, _ = orc.render(sc, orc.make_params(seed=3, spp=8, max_depth=md))
        _film_close(film, ref, frac=frac, rtol=2e-3 if md == 2 else 5e-3)
        assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= (1e-4 if md == 2 else 3e-3) * rst.closest_hit_rays + 1
        integ.close()


@pytest.mark.parametrize("kind", list(scenes.SPHERE_KINDS))
def test_sphere_scene_films(kind):
    sc = scenes.tiny_scene(kind, resolution=(24, 24)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 8, "seed": 3})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=3, spp=8))
    _film_close(film, ref, frac=0.99)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 1e-3 * rst.closest_hit_rays
    integ.close()


def test_sphere_validation_errors():
    import ctypes as C
    b = _sphere_only([0.0]); sc = b.build()
    lib = ffi.load_library(); h = C.c_void_p()
    prims = sc.arrays["prims"].copy(); prims["light"][0] = 0                      # no such light
    bad = ffi.SgSceneDesc.from_buffer_copy(sc.desc); bad.primitives = prims.ctypes.data_as(C.POINTER(ffi.SgPrimitive))
    assert lib.sg_scene_create(C.byref(bad), C.byref(h)) != 0 and b"sphere" in lib.sg_last_error()
    prims = sc.arrays["prims"].copy(); prims["tri"][0] = 7
    bad = ffi.SgSceneDesc.from_buffer_copy(sc.desc); bad.primitives = prims.ctypes.data_as(C.POINTER(ffi.SgPrimitive))
    assert lib.sg_scene_create(C.byref(bad), C.byref(h)) != 0


def test_reference_sphere_predicates_on_gpu():
    """shape.rs:299-342 (`sphere_basic`, `sphere_partial_basic`) through sg_trace(any_hit)."""
    from test_oracle_sphere import REF_SPHERE_PREDICATES, _unit_sphere_world
    for zmin, zmax, o, d, want in REF_SPHERE_PREDICATES:
        integ = create_integrator("wavefront", {}, _unit_sphere_world(zmin, zmax))
        h = integ.trace(np.array([o], np.float32), np.array([d], np.float32), [np.inf], any_hit=True)
        assert (h["prim"][0] >= 0) == want, (zmin, zmax, o, d)
        integ.close()


# ---- BASELINE.json's full sizes ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["mesh1m", "instanced"])
def test_fullsize_raycast_parity(name):
    """SURVEY 8d ray-cast parity set at the FULL C2 (1.0 M triangles) and C4 (10 M instanced triangles, fixed-instancing
    mode) scenes: 2^20 rays, first-hit primitive / instance, t, barycentrics bit-exact; visit counters equal the oracle's."""
    sc = scenes.CONFIGS[name]["builder"](resolution=(64, 64)).build()
    integ = create_integrator("wavefront", {}, sc)
    n = 1 << 20
    o, d = _ray_set(sc, n, seed=7)
    tmax = np.full(n, np.inf, np.float32)
    got, gst = integ.trace(o, d, tmax, want_stats=True)
    ref, rst = orc.trace(sc, o, d, tmax)
    assert _assert_hits_equal(got, ref) == 4
    assert gst.nodes_visited == rst.nodes_visited and gst.tris_tested == rst.tris_tested
    assert (ref["prim"] >= 0).mean() > 0.2
    any_g = integ.trace(o, (d * np.float32(5.0)).astype(np.float32), np.full(n, np.float32(0.9999)), any_hit=True)
    any_r, _ = orc.trace(sc, o, (d * np.float32(5.0)).astype(np.float32), np.full(n, np.float32(0.9999)), any_hit=True)
    assert np.array_equal(any_g["prim"], any_r["prim"])
    integ.close()


def test_fullsize_film_properties():
    """C2 at its full configuration (1024x1024, 64 spp): size-independent film properties -- every pixel's weight sum is
    exactly spp, the two halves of the sample range add up to the full render (the multi-GPU decomposition), and a 48x48
    window agrees with the oracle rendered on that window at full spp."""
    import torch
    cfg = scenes.CONFIGS["mesh1m"]; W, H = cfg["resolution"]; spp = cfg["spp"]
    sc = cfg["builder"](resolution=(W, H)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": spp})
    opts = Options(seed=0, pixel_samples=spp)
    full = torch.zeros((W * H, 4), dtype=torch.float64, device="cuda"); halves = torch.zeros_like(full)
    integ.render_device(opts, full.data_ptr())
    integ.render_device(opts, halves.data_ptr(), sample_range=(0, spp // 2)); integ.render_device(opts, halves.data_ptr(), sample_range=(spp // 2, spp))
    torch.cuda.synchronize()
    assert bool((full[:, 3] == spp).all())
    assert torch.allclose(full, halves, rtol=1e-9, atol=1e-12)
    x0, y0, c = 488, 560, 48
    scc = cfg["builder"](resolution=(W, H), crop=(x0, y0, x0 + c, y0 + c)).build()
    ref, _, _ = orc.render(scc, orc.make_params(seed=0, spp=spp))
    win = full.cpu().numpy().reshape(H, W, 4)[y0:y0 + c, x0:x0 + c].reshape(-1, 4)
    _film_close(win, ref, frac=0.995)
    integ.close()


# ---- bilinear patches ------------------------------------------------------------------------
def test_patch_raycast_parity():
    """Random + aimed rays on the patch scene (wavy patch grid, twisted patch, planar quad, triangles): primitive, (u, v) and
    t bit-exact against the oracle (quadratic + 3x3 determinants are pure f32 with the reference's explicit FMAs)."""
    b = scenes.patch_tiny_scene("patches"); sc = b.build()
    integ = create_integrator("wavefront", {}, sc)
    rng = np.random.default_rng(12)
    n = 1 << 16
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32); o[:, 1] = np.abs(o[:, 1]) + 0.8
    tgt = np.stack([rng.uniform(-1.5, 2.0, n), rng.uniform(0.0, 1.2, n), rng.uniform(-1.2, 1.9, n)], 1).astype(np.float32)
    d = (tgt - o).astype(np.float32)
    d[n // 2:] = rng.standard_normal((n - n // 2, 3)).astype(np.float32)
    o = b.render_from_world.apply_points_f32(o)
    for tmax in (np.inf, 3.0):
        t = np.full(n, tmax, np.float32)
        got, gst = integ.trace(o, d, t, want_stats=True)
        ref, rst = orc.trace(sc, o, d, t)
        assert _assert_hits_equal(got, ref) == 4
        assert gst.nodes_visited == rst.nodes_visited and gst.tris_tested == rst.tris_tested
        hit = ref["prim"] >= 0
        is_patch = np.array([m.flags & ffi.SG_MESH_BILINEAR for m in sc.arrays["meshes"]], bool)[sc.arrays["prims"]["mesh"][ref["prim"][hit]]]
        assert is_patch.mean() > 0.3
        any_g = integ.trace(o, d, np.minimum(t, np.float32(0.9999)), any_hit=True)
        any_r, _ = orc.trace(sc, o, d, np.minimum(t, np.float32(0.9999)), any_hit=True)
        assert np.array_equal(any_g["prim"], any_r["prim"])
    integ.close()


@pytest.mark.parametrize("kind", list(scenes.PATCH_KINDS))
def test_patch_scene_films(kind):
    sc = scenes.tiny_scene(kind, resolution=(24, 24)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 8, "seed": 4})
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=4, spp=8))
    _film_close(film, ref, frac=0.99)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 1e-3 * rst.closest_hit_rays
    integ.close()


def test_orthographic_camera_rays_bit_exact():
    sc = scenes.tiny_scene("ortho", resolution=(16, 16)).build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 4})
    rng = np.random.default_rng(1)
    xy = rng.integers(0, 16, (512, 2)).astype(np.int32); si = rng.integers(0, 4, 512).astype(np.int32)
    opts = Options(seed=9, pixel_samples=4)
    rays, lam = integ.camera_rays(opts, xy, si)
    r2, l2 = orc.camera_rays(sc, orc.make_params(seed=9, spp=4), xy, si)
    assert np.array_equal(rays, r2) and np.allclose(lam, l2, rtol=2e-6)
    integ.close()


@pytest.mark.parametrize("kind", ["tex", "texewa", "diffuse"])
def test_thin_lens_camera_parity(kind):
    """lens_radius > 0 (camera.rs:1026-1038, auxiliary rays :1044-1068): camera rays and the film of a textured scene (whose
    footprints come from the lens-aware auxiliary rays) against the oracle."""
    from test_oracle_render import _lens_scene
    sc = _lens_scene(kind, lens_radius=0.15, focal_distance=3.0)
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 8, "seed": 6})
    rng = np.random.default_rng(2)
    xy = rng.integers(0, 16, (1024, 2)).astype(np.int32); si = rng.integers(0, 8, 1024).astype(np.int32)
    rays, lam = integ.camera_rays(Options(seed=6, pixel_samples=8), xy, si)
    r2, l2 = orc.camera_rays(sc, orc.make_params(seed=6, spp=8), xy, si)
    assert np.allclose(rays, r2, rtol=2e-6, atol=2e-7) and np.allclose(lam, l2, rtol=2e-6)   # the concentric disk map calls sin / cos (different libms)
    assert (rays == r2).mean() > 0.5
    film = integ.render(Options()).copy()
    ref, rst, _ = orc.render(sc, orc.make_params(seed=6, spp=8))
    _film_close(film, ref, frac=0.99)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 2
    integ.close()


@pytest.mark.parametrize("kind", ["diffuse", "conductor", "mirror", "glass", "roughglass", "coated", "thinglass", "tex", "texbump", "coatedcond", "mix", "texparams"])
def test_force_diffuse_films(kind):
    """Options::force_diffuse (interaction.rs:258-273): every BSDF replaced by DiffuseBxDF(rho_hd(wo, one sample)) on the same frame,
    three extra draws from the path's generator per hit -- the CUDA path (its own kernel instantiations) against the oracle on the
    same streams."""
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    integ = create_integrator("wavefront", {"maxdepth": 5}, sc, {"pixelsamples": 8, "seed": 5})
    film = integ.render(Options(force_diffuse=True)).copy()
    p = orc.make_params(seed=5, spp=8); p.option_flags = ffi.SG_OPT_FORCE_DIFFUSE
    ref, rst, _ = orc.render(sc, p)
    _film_close(film, ref, frac=0.99)
    assert abs(int(integ.stats.closest_hit_rays) - int(rst.closest_hit_rays)) <= 2
    assert abs(int(integ.stats.shadow_rays) - int(rst.shadow_rays)) <= 2
    plain = integ.render(Options()).copy()
    integ.film[:] = 0
    assert not np.array_equal(plain, film)
    integ.close()


def test_degenerate_triangles_are_skipped_like_the_reference():
    """triangle.rs:181: a triangle whose cross product has zero length never reports a hit.  The traversal kernels read that
    answer from a per-triangle flag set at scene upload (kDegenerateBit) instead of evaluating it per test: a mesh with repeated
    and collinear vertices in front of a regular one must trace exactly like the oracle, visit counts included."""
    from shimmer_b200.host import SceneBuilder
    b = SceneBuilder()
    b.set_camera(pos=(0.0, 1.0, -3.0), look=(0.0, 0.5, 0.0), up=(0, 1, 0), fov=45.0, resolution=(16, 16))
    white = b.diffuse(("const", 0.6))
    gp = np.array([[-2, 0, -2], [-2, 0, 2], [2, 0, 2], [2, 0, -2]], np.float32)
    b.add_mesh(gp, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), white)
    rng = np.random.default_rng(5)
    n = 40
    P = (rng.random((3 * n, 3)).astype(np.float32) - np.float32(0.5)) * np.float32(1.5) + np.array([0, 0.9, 0], np.float32)
    I = np.arange(3 * n, dtype=np.uint32).reshape(n, 3)
    I[0::4, 2] = I[0::4, 1]                                               # repeated vertex
    P[3 * 1 + 2::12] = (P[3 * 1::12] + P[3 * 1 + 1::12]) * np.float32(0.5)  # exactly representable midpoints: collinear
    I[2::4] = I[2::4][:, [0, 0, 0]]                                       # a point
    b.add_mesh(P, I, b.diffuse(("const", 0.3)))
    lp = np.array([[-0.5, 2.5, -0.5], [0.5, 2.5, -0.5], [0.5, 2.5, 0.5], [-0.5, 2.5, 0.5]], np.float32)
    b.add_mesh(lp, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), white, area_light=dict(L=("const", 1.0), scale=20.0, two_sided=False))
    sc = b.build()
    integ = create_integrator("wavefront", {}, sc, {"pixelsamples": 16})
    m = 1 << 15
    o, d = _ray_set(sc, m, seed=2)
    tmax = np.full(m, np.inf, np.float32)
    got, gst = integ.trace(o, d, tmax, want_stats=True)
    ref, rst = orc.trace(sc, o, d, tmax)
    assert _assert_hits_equal(got, ref) == 4
    assert gst.nodes_visited == rst.nodes_visited and gst.tris_tested == rst.tris_tested
    got2 = integ.trace(o, (d * np.float32(3.0)).astype(np.float32), np.full(m, np.float32(0.9999), np.float32), any_hit=True)
    ref2, _ = orc.trace(sc, o, (d * np.float32(3.0)).astype(np.float32), np.full(m, np.float32(0.9999), np.float32), any_hit=True)
    assert np.array_equal(got2["prim"], ref2["prim"])
    film = integ.render(Options(seed=0, pixel_samples=16)).copy()          # the wavefront's own traversal kernels (k_trace)
    rfilm, rs, _ = orc.render(sc, orc.make_params(seed=0, spp=16))
    _film_close(film, rfilm)
    assert integ.stats.closest_hit_rays == rs.closest_hit_rays and integ.stats.shadow_rays == rs.shadow_rays
    integ.close()
