"""Oracle checks for image textures, MIP filtering and screen-space differentials.
Pins: the reference's own unit tests where they exist (transform.rs:916-940 rotate_from_to, math.rs:569-575 poly
coefficient order, color.rs:1151-1181 RGB -> albedo spectrum -> RGB round trip) and independent numpy
restatements of the filtering formulas (image.rs:619-646, mipmap.rs:121-201)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import orc
from shimmer_b200 import host, rgb2spec, scenes

f32 = np.float32


def _rot(frm, to):
    out = np.zeros(9, np.float32)
    orc.lib().orc_rotate_from_to(orc.fa(frm).ctypes.data, orc.fa(to).ctypes.data, out.ctypes.data)
    return out.reshape(3, 3)


def test_rotate_from_to_reference_vectors():                               # transform.rs:916-940
    Z, X, Y = np.array([0, 0, 1], f32), np.array([1, 0, 0], f32), np.array([0, 1, 0], f32)
    for to in (Z, X, Y):
        assert np.array_equal(_rot(Z, to) @ Z, to)                        # assert_eq! in the reference: exact
    a = np.array([0.1, 0.2, 0.3], f32); a /= np.linalg.norm(a)
    b = np.array([0.4, 0.5, 0.6], f32); b /= np.linalg.norm(b)
    assert np.allclose(_rot(a, b) @ a, b, atol=1e-6)


def test_sigmoid_polynomial_order_and_limits():                            # math.rs:569-575 + color.rs:352-383
    L = orc.lib()
    get = lambda c, lam: L.orc_sigmoid_poly_get(orc.fa(c).ctypes.data, C.c_float(lam))
    # poly(x, [c2, c1, c0]) = c2 + c1 x + c0 x^2: with c = (3, 2, 1) and x = 2 the polynomial is 17 (the reference test)
    x = 17.0
    assert abs(get([3.0, 2.0, 1.0], 2.0) - (0.5 + x / (2.0 * np.sqrt(1.0 + x * x)))) < 1e-7
    assert get([0.0, 0.0, 0.0], 500.0) == 0.5
    assert get([0.0, 0.0, np.inf], 500.0) == 1.0 and get([0.0, 0.0, -np.inf], 500.0) == 0.0


@pytest.fixture(scope="module")
def tex_scene():
    return scenes.tiny_scene("tex", resolution=(16, 16)).build()


def test_rgb_albedo_round_trip(tex_scene):                                 # color.rs:1151-1181
    """RGB -> sigmoid spectrum -> (x D65) XYZ -> RGB comes back within 0.01 (the reference's own bar; our stand-in
    table is 16^3 instead of 64^3, so the bar is 0.03 here)."""
    T = host.tables()
    lam = np.arange(360, 831, dtype=np.float64)
    illum = np.asarray(host.spectrum_dense(host.named_spectrum("stdillum-D65")), np.float64)
    X, Y, Z = (np.asarray(T["CIE_" + c], np.float64) for c in "XYZ")
    M = np.array([[3.240479, -1.537150, -0.498535], [-0.969256, 1.875991, 0.041556], [0.055648, -0.204043, 1.057311]])
    rng = np.random.default_rng(0)
    L = orc.lib()
    worst = 0.0
    for _ in range(100):
        rgb = rng.random(3).astype(f32)
        c = np.zeros(3, f32)
        L.orc_rgb2spec_fetch(tex_scene.ptr(), rgb.ctypes.data, c.ctypes.data)
        x = c[0] * lam * lam + c[1] * lam + c[2]
        s = 0.5 + x / (2.0 * np.sqrt(1.0 + x * x))                         # RgbSigmoidPolynomial::get, pinned above
        k = np.sum(illum * Y)
        xyz = np.array([np.sum(s * illum * X), np.sum(s * illum * Y), np.sum(s * illum * Z)]) / k
        worst = max(worst, np.abs(M @ xyz - rgb).max())
    assert worst < 0.03, worst


def _np_channel(level, x, y, c, wrap):
    H, W = level.shape[:2]
    if wrap == "repeat":
        x %= W; y %= H
    elif wrap == "clamp":
        x = min(max(x, 0), W - 1); y = min(max(y, 0), H - 1)
    elif not (0 <= x < W and 0 <= y < H):
        return f32(0.0)
    return level[y, x, c]


def _np_bilerp(level, st, c, wrap):                                        # image.rs:619-646
    H, W = level.shape[:2]
    x = f32(st[0]) * f32(W) - f32(0.5); y = f32(st[1]) * f32(H) - f32(0.5)
    xi, yi = int(np.floor(x)), int(np.floor(y))
    dx, dy = f32(x - f32(xi)), f32(y - f32(yi))
    v = [_np_channel(level, xi + a, yi + b, c, wrap) for b in (0, 1) for a in (0, 1)]
    one = f32(1.0)
    return (one - dx) * (one - dy) * v[0] + dx * (one - dy) * v[1] + (one - dx) * dy * v[2] + dx * dy * v[3]


def test_bilinear_and_level_selection_match_numpy():
    img = scenes.procedural_image(32, 1)
    for wrap in ("repeat", "clamp", "black"):
        b = scenes.SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
        t = b.image_texture(img, filter="bilinear", wrap=wrap)
        m = b.diffuse(("const", 0.5), reflectance_tex=t)
        b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), m, uv=np.array([[0, 0], [1, 0], [0, 1]], f32))
        sc = b.build()
        levels = b.textures[t]["levels"]
        rng = np.random.default_rng(1)
        q = np.zeros((200, 6), f32)
        q[:, :2] = rng.random((200, 2)) * 3.0 - 1.0                        # outside [0,1]: exercises the wrap modes
        q[:, 2:] = (rng.random((200, 4)) - 0.5) * np.float32(0.2) * (rng.random((200, 1)) < 0.8)
        got = orc.texture_eval(sc, t, q, as_float=True)[:, 0]
        for i in range(200):
            u, v = q[i, 0], q[i, 1]
            st = (f32(u), f32(1.0) - f32(v))                               # texture.rs:396-399: t flipped
            width = f32(2.0) * max(abs(q[i, 2]), abs(q[i, 4]), abs(q[i, 3]), abs(q[i, 5]))     # dst0 = (dsdx, dtdx), dst1 = (dsdy, dtdy)
            nl = len(levels)
            level = f32(nl - 1) + np.log2(max(width, f32(1e-8)))
            if level >= nl - 1:
                exp = levels[-1][0, 0, 0]
            else:
                exp = _np_bilerp(levels[max(0, int(np.floor(level)))], st, 0, wrap)
            assert abs(got[i] - exp) <= 2e-6, (wrap, i, got[i], exp)


def test_pyramid_is_a_box_filter_and_ends_at_1x1():                        # image.rs:699-787
    b = scenes.SceneBuilder()
    t = b.image_texture(scenes.procedural_image(16, 3))
    lv = b.textures[t]["levels"]
    assert [l.shape[:2] for l in lv] == [(16, 16), (8, 8), (4, 4), (2, 2), (1, 1)]
    assert np.allclose(lv[1][3, 2], lv[0][6:8, 4:6].reshape(4, 3).mean(axis=0), atol=1e-6)
    assert np.allclose(lv[-1][0, 0], lv[0].reshape(-1, 3).mean(axis=0), atol=1e-5)
    r = b.image_texture(np.arange(8, dtype=f32).reshape(2, 4))              # non-square: one dimension reaches 1 first
    assert [l.shape[:2] for l in b.textures[r]["levels"]] == [(2, 4), (1, 2), (1, 1)]
    with pytest.raises(ValueError):
        b.image_texture(np.zeros((3, 5), f32))


def test_filters_agree_on_a_constant_image_and_ewa_is_normalised():
    b = scenes.SceneBuilder(); b.set_camera((0, 0, -3), (0, 0, 0), (0, 1, 0), 40.0, (8, 8))
    ids = [b.image_texture(np.full((16, 16), 0.37, f32), filter=f) for f in ("point", "bilinear", "trilinear", "ewa")]
    ramp = b.image_texture(np.tile(np.linspace(0, 1, 32, dtype=f32), (32, 1)), filter="ewa", wrap="clamp")
    m = b.diffuse(("const", 0.5), reflectance_tex=ids[0])
    b.add_mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], f32), np.array([[0, 1, 2]], np.uint32), m)
    sc = b.build()
    rng = np.random.default_rng(2)
    q = np.zeros((64, 6), f32); q[:, :2] = rng.random((64, 2)); q[:, 2:] = (rng.random((64, 4)) - 0.5) * 0.3
    for t in ids:
        assert np.allclose(orc.texture_eval(sc, t, q, as_float=True), 0.37, atol=1e-6)
    # EWA of a horizontal ramp with a symmetric footprint reproduces the ramp value at the lookup point
    q2 = np.zeros((16, 6), f32); q2[:, 0] = np.linspace(0.3, 0.7, 16); q2[:, 1] = 0.5; q2[:, 2] = 0.04; q2[:, 5] = 0.04
    got = orc.texture_eval(sc, ramp, q2, as_float=True)[:, 0]
    assert np.allclose(got, (q2[:, 0] * 32 - 0.5) / 31.0, atol=0.02)


def test_disable_texture_filtering_and_constant_displacement_paths(tex_scene):
    """disable_texture_filtering zeroes every footprint (interaction.rs:287-295) -> a bilinear texture is read at level 0."""
    p = orc.make_params(seed=3, spp=2)
    a, _, _ = orc.render(tex_scene, p)
    p2 = orc.make_params(seed=3, spp=2, flags=0); p2.option_flags = 4        # SG_OPT_DISABLE_TEXTURE_FILTERING
    bfilm, _, _ = orc.render(tex_scene, p2)
    assert np.isfinite(a).all() and np.isfinite(bfilm).all() and not np.array_equal(a, bfilm)


@pytest.mark.parametrize("kind", scenes.TEXTURED_KINDS)
def test_textured_scene_golden_film(kind):
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tiny_films.json")))[kind]
    sc = scenes.tiny_scene(kind, resolution=(16, 16)).build()
    film, st, _ = orc.render(sc, orc.make_params(seed=5, spp=4))
    assert st.closest_hit_rays == gold["closest_hit_rays"] and st.shadow_rays == gold["shadow_rays"]
    assert np.allclose(film.sum(axis=0), gold["film_sum"], rtol=1e-9)
    assert np.isfinite(film).all()
