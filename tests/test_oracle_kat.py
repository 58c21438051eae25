"""Pins the CPU oracle against every known-answer vector the reference's own unit tests hold for the hot
path, plus the published xoshiro256++ / SplitMix64 vectors for the third-party RNG.
All file:line citations are relative to /root/reference/src."""
import ctypes as C

import numpy as np
import pytest

import orc
from orc import fa
from shimmer_b200 import ffi, host, scenes


def L():
    return orc.lib()


# ---- rand 0.8.5 SmallRng (xoshiro256++), sampler.rs:103-132 -------------------------------------
def test_xoshiro256pp_published_vector():
    state = np.array([1, 2, 3, 4], np.uint64)
    out = np.zeros(4, np.uint64)
    L().orc_rng_u64(0, 1, state.ctypes.data, 4, out.ctypes.data, None)
    assert out.tolist() == [41943041, 58720359, 3588806011781223, 3591011842654386]


def test_seed_from_u64_zero_is_splitmix64():
    out = np.zeros(4, np.uint64); st = np.zeros(4, np.uint64)
    L().orc_rng_u64(0, 0, None, 4, out.ctypes.data, st.ctypes.data)
    assert [hex(x) for x in st.tolist()] == ["0xe220a8397b1dcdaf", "0x6e789e6aa1b965f4", "0x6c45d188009454f", "0xf88bb8a8724c81ec"]
    assert [hex(x) for x in out.tolist()] == ["0x53175d61490b23df", "0x61da6f3dc380d507", "0x5c0fdf91ec9a7bfc", "0x2eebf8c3bbe5e1a"]
    f = np.zeros(4, np.float32)
    L().orc_sampler_fill(0, 1, 0, 0, 4, f.ctypes.data)
    # (next_u64 >> 32 >> 8) * 2^-24
    exp = [((x >> 40) * 2.0 ** -24) for x in out.tolist()]
    assert f.tolist() == [np.float32(e) for e in exp]
    assert f.tolist() == pytest.approx([0.3245752453804016, 0.38223928213119507, 0.3596171736717224, 0.0114554762840271], abs=0)


def test_sampler_range_like_triangle_tests():
    # shape/triangle.rs:773-848 draws 100 values from IndependentSampler::new(0, 100) and checks [0,1)
    f = np.zeros(10000, np.float32)
    L().orc_sampler_fill(0, 1, 0, 0, len(f), f.ctypes.data)
    assert f.min() >= 0.0 and f.max() < 1.0 and abs(f.mean() - 0.5) < 0.02


def test_streams_differ_per_pixel_and_sample():
    a = np.zeros(8, np.float32); b = np.zeros(8, np.float32); c = np.zeros(8, np.float32)
    L().orc_sampler_fill(7, 0, 10, 0, 8, a.ctypes.data)
    L().orc_sampler_fill(7, 0, 10, 1, 8, b.ctypes.data)
    L().orc_sampler_fill(7, 0, 11, 0, 8, c.ctypes.data)
    assert not np.array_equal(a, b) and not np.array_equal(a, c) and not np.array_equal(b, c)


# ---- math.rs / float.rs ---------------------------------------------------------------------------
def test_math_rs_vectors():
    assert L().orc_lerp(0.45, 0.0, 10.0) == 4.5                        # math.rs:548-554
    assert L().orc_difference_of_products(10.0, 10.0, 5.0, 5.0) == 75.0   # math.rs:556-566


def test_next_float_up_down_match_nextafter():                          # float.rs:172-211
    rng = np.random.default_rng(0)
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32) * 1e3, fa([0.0, -0.0, 1.0, -1.0, 1e-38, -1e-38, 3e38])])
    for v in vals:
        assert L().orc_next_float_up(float(v)) == np.nextafter(np.float32(v), np.float32(np.inf))
        assert L().orc_next_float_down(float(v)) == np.nextafter(np.float32(v), np.float32(-np.inf))
    assert L().orc_next_float_up(float("inf")) == float("inf")
    assert L().orc_next_float_down(float("-inf")) == float("-inf")


def test_gamma():                                                       # float.rs:88-90
    eps = np.float32(2.0 ** -24)
    for n in (2, 3, 5, 6, 7):
        assert L().orc_gamma(n) == np.float32(np.float32(n) * eps) / np.float32(np.float32(1) - np.float32(n) * eps)


# ---- sampling.rs:801-836 -----------------------------------------------------------------------------
def test_visible_wavelength_pdf_range():
    assert L().orc_visible_wavelengths_pdf(359.9) == 0.0
    assert L().orc_visible_wavelengths_pdf(830.1) == 0.0
    assert L().orc_visible_wavelengths_pdf(538.0) == pytest.approx(0.0039398042)
    # spectrum.rs:861-888: the pdf integrates to ~1 over the sampled range
    u = (np.arange(20000) + 0.5) / 20000
    lam = np.array([L().orc_sample_visible_wavelengths(float(x)) for x in u[::40]])
    assert lam.min() >= 360.0 and lam.max() <= 830.0
    est = np.mean([1.0 / L().orc_visible_wavelengths_pdf(float(l)) for l in lam]) / (830.0 - 360.0)
    assert est == pytest.approx(1.0, rel=0.02)


# ---- bxdf.rs:1839-1903 (values lifted from pbrt; the reference's approx_eq! cannot fail, ours can) ----
def test_mf_distrib_vector():
    wm = fa([-0.430063188, -0.881908476, 0.193088099]); wi = fa([0.568110108, 0.816620350, 0.101893365])
    a = 0.0299999993
    assert L().orc_tr_d(a, a, wm.ctypes.data) == pytest.approx(0.000309075956, rel=2e-5)
    # The reference test also lists g = 0.954060972, but it uses `approx_eq!` (a bool that is discarded), so
    # it cannot fail -- and that number is NOT what the reference's own g() (scattering.rs:130-146) yields for
    # these inputs.  Evaluating the reference formula in f64 gives 0.97391665; that is what we pin.
    def lam(w):
        c2 = w[2] ** 2; s2 = max(0.0, 1 - c2); t2 = s2 / c2; st = s2 ** 0.5
        return (-1 + (1 + ((w[0] / st * a) ** 2 + (w[1] / st * a) ** 2) * t2) ** 0.5) / 2
    g_ref = 1.0 / (1.0 + lam(wm.astype(np.float64)) + lam(wi.astype(np.float64)))
    assert g_ref == pytest.approx(0.97391665, rel=1e-6)
    assert L().orc_tr_g(a, a, wm.ctypes.data, wi.ctypes.data) == pytest.approx(g_ref, rel=2e-6)


def test_dielectric_sample_f_vector():
    wo = fa([-0.419299453, -0.656406343, 0.627151370]); u2 = fa([0.0488742627, 0.941848040])
    out = np.zeros(10, np.float32)
    ok = L().orc_dielectric_sample_f(1.5, 0.0, 0.0, wo.ctypes.data, 0.237656280, u2.ctypes.data, out.ctypes.data)
    assert ok == 1
    assert int(out[8]) == 16 | 2                       # SPECULAR_TRANSMISSION
    assert out[7] == pytest.approx(0.940032840, rel=2e-6)
    assert out[9] == pytest.approx(1.5)
    assert out[:4].tolist() == pytest.approx([0.488867134] * 4, rel=2e-6)
    assert out[4:7].tolist() == pytest.approx([0.279532969, 0.437604219, -0.854613364], rel=2e-6)


# ---- spectra/spectrum.rs:654-888 -----------------------------------------------------------------------
def test_blackbody_vectors():
    for lam, t, ref in ((483.0, 6000.0, 3.1849e13), (600.0, 6000.0, 2.86772e13), (500.0, 3700.0, 1.59845e12), (600.0, 4500.0, 7.46497e12)):
        assert abs(L().orc_blackbody(lam, t) - ref) / ref < 0.001
    for t in (2700.0, 3000.0, 4500.0, 5600.0, 6000.0):
        lm = 2.8977721e-3 / t * 1e9
        r = [L().orc_blackbody(float(np.float32(k * lm)), t) for k in (0.99, 1.0, 1.01)]
        assert r[0] < r[1] > r[2]


def _spectrum_scene():
    b = scenes.tiny_scene("diffuse", resolution=(4, 4))
    ids = dict(const=b.spectrum(("const", 5.0)),
               pl=b.spectrum(("pl", fa([400, 500, 600]), fa([1.0, 3.0, 2.0]))),
               dense=b.spectrum(("dense", np.arange(471, dtype=np.float32))))
    return b.build(), ids


def test_spectrum_get_and_sample():
    sc, ids = _spectrum_scene()
    g = lambda i, l: L().orc_spectrum_get(sc.ptr(), i, l)
    assert g(ids["const"], 999.0) == 5.0                                  # spectrum.rs:654-660
    assert g(ids["pl"], 450.0) == 2.0 and g(ids["pl"], 550.0) == 2.5      # piecewise lerp
    assert g(ids["pl"], 399.0) == 0.0 and g(ids["pl"], 601.0) == 0.0      # outside -> 0 (:409-414)
    assert g(ids["pl"], 400.0) == 1.0 and g(ids["pl"], 600.0) == 2.0
    assert g(ids["dense"], 360.9) == 0.0 and g(ids["dense"], 361.0) == 1.0   # get truncates (:265)
    assert g(ids["dense"], 359.0) == 0.0 and g(ids["dense"], 831.0) == 0.0
    lam = fa([360.4, 360.5, 829.6, 830.6]); out = np.zeros(4, np.float32)
    L().orc_spectrum_sample(sc.ptr(), ids["dense"], lam.ctypes.data, out.ctypes.data)
    assert out.tolist() == [0.0, 1.0, 470.0, 0.0]                         # sample ROUNDS half away (:283)


def test_cie_tables_integrate_to_one():                                   # spectrum.rs:702-722
    for c in "XYZ":
        assert float(np.sum(host.cie(c).astype(np.float64))) / 106.856895 == pytest.approx(1.0, abs=0.005)


def test_d65_normalisation_and_srgb_matrix():
    d65 = host.spectrum_dense(host.named_spectrum("stdillum-D65"))
    y = float(np.sum(d65.astype(np.float64) * host.cie("Y").astype(np.float64)))
    assert y == pytest.approx(106.856895, rel=1e-4)                       # from_interleaved(normalize = true)
    m = host.srgb_output_matrix()
    ref = np.array([[3.2406, -1.5372, -0.4986], [-0.9689, 1.8758, 0.0415], [0.0557, -0.2040, 1.0570]])
    assert np.allclose(m, ref, atol=3e-3)


# ---- geometry: watertight triangle + slab test --------------------------------------------------------
def test_triangle_intersection_basics():
    tri = fa([0, 0, 1, 1, 0, 1, 0, 1, 1]); out = np.zeros(4, np.float32)
    o = fa([0.25, 0.25, 0]); d = fa([0, 0, 1])
    assert L().orc_tri_intersect(o.ctypes.data, d.ctypes.data, float("inf"), tri.ctypes.data, out.ctypes.data) == 1
    assert out.tolist() == [0.5, 0.25, 0.25, 1.0]
    assert L().orc_tri_intersect(o.ctypes.data, d.ctypes.data, 0.5, tri.ctypes.data, out.ctypes.data) == 0      # t_max
    d2 = fa([0, 0, -1])
    assert L().orc_tri_intersect(o.ctypes.data, d2.ctypes.data, float("inf"), tri.ctypes.data, out.ctypes.data) == 0  # behind
    deg = fa([0, 0, 1, 1, 0, 1, 2, 0, 1])
    assert L().orc_tri_intersect(o.ctypes.data, d.ctypes.data, float("inf"), deg.ctypes.data, out.ctypes.data) == 0   # degenerate
    # shared edge: exactly one of two triangles sharing the diagonal reports the hit (watertightness)
    t1 = fa([0, 0, 1, 1, 0, 1, 1, 1, 1]); t2 = fa([0, 0, 1, 1, 1, 1, 0, 1, 1])
    oe = fa([0.5, 0.5, 0])
    h1 = L().orc_tri_intersect(oe.ctypes.data, d.ctypes.data, float("inf"), t1.ctypes.data, out.ctypes.data)
    h2 = L().orc_tri_intersect(oe.ctypes.data, d.ctypes.data, float("inf"), t2.ctypes.data, out.ctypes.data)
    assert h1 + h2 >= 1


def test_slab_test_axis_parallel_and_nan():
    bmin, bmax = fa([0, 0, 0]), fa([1, 1, 1])
    def f(o, d, t=float("inf")):
        o, d = fa(o), fa(d)      # keep the arrays alive across the call
        return L().orc_bounds_intersect(bmin.ctypes.data, bmax.ctypes.data, o.ctypes.data, d.ctypes.data, t)
    assert f([0.5, 0.5, -1], [0, 0, 1]) == 1            # zero components -> +-inf inv_dir
    assert f([1.5, 0.5, -1], [0, 0, 1]) == 0
    assert f([0.5, 0.5, -1], [0, 0, 1], 0.5) == 0       # t_min < ray_t_max
    assert f([0.5, 0.5, 2], [0, 0, 1]) == 0             # t_max > 0
    assert f([0.5, 0.5, 0.5], [1, 0, 0]) == 1           # origin inside
    assert f([0.0, 0.5, -1], [0, 0, 1]) in (0, 1)       # on the slab plane: 0*inf = NaN, must not crash


# ---- film output stage: f16 quantisation + get_image ---------------------------------
def test_f16_round_matches_ieee_rne():
    """`half::f16::from_f32` (third-party, pinned 2.2.1) is IEEE round-to-nearest-even; numpy's float16 cast is an
    independent implementation of the same rounding: all 2^16 half values, their neighbours/midpoints, and 2^20 random
    bit patterns must agree bit for bit."""
    rng = np.random.default_rng(11)
    halves = np.arange(65536, dtype=np.uint16).view(np.float16).astype(np.float32)
    fin = halves[np.isfinite(halves)]
    mids = (fin[:-1].astype(np.float64) + np.sort(fin)[1:].astype(np.float64)) / 2          # not all are midpoints; fine
    x = np.concatenate([halves, np.nextafter(fin, np.float32(np.inf)), np.nextafter(fin, np.float32(-np.inf)),
                        mids.astype(np.float32), rng.integers(0, 2 ** 32, 1 << 20, dtype=np.uint64).astype(np.uint32).view(np.float32),
                        np.array([65504.0, 65519.99, 65520.0, 65536.0, 1e-8, 2.0 ** -25, 2.0 ** -24, 5.96e-8, -0.0, np.inf, -np.inf], np.float32)])
    with np.errstate(over="ignore", invalid="ignore"):
        exp = x.astype(np.float16)
    got, bits = orc.f16_round(x)
    nan = np.isnan(x)
    assert np.array_equal(bits[~nan], exp.view(np.uint16)[~nan])
    assert np.isnan(got[nan]).all()
    assert np.array_equal(got[~nan], exp.astype(np.float32)[~nan])


def test_film_get_image_semantics(cornell64):
    """film.rs:647-707: weight-normalise, output matrix, the fp16 clamp as written (a too-large g clamps r and stays,
    so it becomes +inf in f16), NaN -> 0 (image.rs:649), bottom-up raster order (image.rs:1350)."""
    M = np.array(list(cornell64.desc.film.output_rgb_from_sensor_rgb), np.float32).reshape(3, 3)
    Minv = np.linalg.inv(M.astype(np.float64))
    want = np.array([[0.25, 0.5, 1.0], [1e5, 1.0, 2.0], [1.0, 1e5, 2.0], [1.0, 2.0, 1e5], [np.nan, 1.0, 1.0], [0.1, 0.2, 0.3]])
    film = np.zeros((6, 4)); film[:, :3] = (want @ Minv.T) * 2.0; film[:, 3] = 2.0
    film[5, 3] = 0.0; film[5, :3] = want[5] @ Minv.T                                         # zero weight: no division
    img = orc.film_get_image(cornell64, film, 3, 2, fp16=True).reshape(-1, 3)
    assert np.allclose(img[0], [0.25, 0.5, 1.0], rtol=2e-3)
    assert img[1, 0] == 65504.0 and abs(img[1, 1] - 1.0) < 0.05           # f32 cancellation through the 3x3 matrix
    assert img[2, 0] == 65504.0 and np.isinf(img[2, 1])                                      # the reference's g/r slip
    assert img[3, 2] == 65504.0 and abs(img[3, 0] - 1.0) < 0.05
    assert img[4, 0] == 0.0                                                                   # NaN -> 0
    assert np.allclose(img[5], [0.1, 0.2, 0.3], rtol=2e-3)
    f32img = orc.film_get_image(cornell64, film, 3, 2, fp16=False).reshape(-1, 3)
    assert np.array_equal(f32img[[0, 5]], orc.develop(cornell64, film)[[0, 5]])
    flip = orc.film_get_image(cornell64, film, 3, 2, fp16=True, bottom_up=True)
    assert np.array_equal(flip[0], img.reshape(2, 3, 3)[1], equal_nan=True) and np.array_equal(flip[1], img.reshape(2, 3, 3)[0], equal_nan=True)
