"""Host-side stand-in for shimmer's `rgbtospec/*.spec` coefficient tables.

shimmer loads pre-computed RGB -> sigmoid-polynomial tables through the third-party `rgb2spec` crate
(src/rgb_to_spectra.rs:16-45); the six `.spec` blobs are missing from the reference checkout
(.MISSING_LARGE_BLOBS).  In the real integration the Rust host passes its loaded table across the C ABI
(SgSceneDesc.rgb2spec_*).  For the synthetic benchmark scenes this module regenerates an sRGB table with the
published optimiser (Jakob & Hanika 2019, `rgb2spec_opt`): Gauss-Newton fit of c0 x^2 + c1 x + c2 in CIELAB,
warm-started along the brightness axis.  Host preparation only -- nothing here is on the render path, and both
the CUDA path and the oracle read the same table.
"""
import os

import numpy as np

_CACHE = {}


def _smoothstep(x):
    return x * x * (3.0 - 2.0 * x)


def _tables():
    from . import host
    T = host.tables()
    lam = np.asarray(T["CIE_LAMBDA"], dtype=np.float64)
    n = 95 * 3 + 1
    fine = np.linspace(360.0, 830.0, n)
    xyz = np.stack([np.interp(fine, lam, np.asarray(T["CIE_" + c], dtype=np.float64)) for c in "XYZ"], axis=1)
    illum_raw = np.asarray(T["CIE_ILLUM_D6500"], dtype=np.float64)           # interleaved (lambda, value)
    il, iv = illum_raw[0::2], illum_raw[1::2]
    illum = np.interp(fine, il, iv)
    h = (830.0 - 360.0) / (n - 1)
    w = np.full(n, 3.0 / 8.0 * h)
    idx = np.arange(n)
    inner = (idx > 0) & (idx < n - 1)
    w[inner & ((idx - 1) % 3 == 2)] *= 2.0
    w[inner & ((idx - 1) % 3 != 2)] *= 3.0
    xyz_to_rgb = np.array([[3.240479, -1.537150, -0.498535], [-0.969256, 1.875991, 0.041556], [0.055648, -0.204043, 1.057311]])
    rgb_to_xyz = np.array([[0.412453, 0.357580, 0.180423], [0.212671, 0.715160, 0.072169], [0.019334, 0.119193, 0.950227]])
    illum = illum / np.sum(illum * w * xyz[:, 1])                 # normalised so that the white point has Y = 1
    wi = (illum * w)[:, None] * xyz                               # n x 3
    whitepoint = wi.sum(axis=0)
    rgb_tbl = wi @ xyz_to_rgb.T                                   # n x 3: rgb response per wavelength
    return fine, rgb_tbl, rgb_to_xyz, whitepoint


def _lab(rgb, rgb_to_xyz, wp):
    xyz = rgb @ rgb_to_xyz.T / wp
    d = 6.0 / 29.0
    f = np.where(xyz > d ** 3, np.cbrt(np.maximum(xyz, 1e-300)), xyz / (3 * d * d) + 4.0 / 29.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], axis=-1)


def build_table(res=16):
    """Returns (scale[res] f32, data[3, res, res, res, 3] f32) in the rgb2spec crate's layout."""
    key = ("srgb", res)
    if key in _CACHE:
        return _CACHE[key]
    cache_file = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "rgb2spec_srgb_%d.npz" % res)
    if os.path.exists(cache_file):
        d = np.load(cache_file)
        _CACHE[key] = (d["scale"], d["data"])
        return _CACHE[key]
    fine, rgb_tbl, rgb_to_xyz, wp = _tables()
    lam_n = (fine - 360.0) / (830.0 - 360.0)
    scale = _smoothstep(_smoothstep(np.arange(res) / (res - 1.0)))
    data = np.zeros((3, res, res, res, 3))
    xs = np.arange(res) / (res - 1.0)
    X, Y = np.meshgrid(xs, xs, indexing="xy")                      # [j, i]: x = i/(res-1), y = j/(res-1)

    def spectrum_rgb(c):                                           # c: (..., 3) -> rgb (..., 3)
        x = (c[..., 0:1] * lam_n + c[..., 1:2]) * lam_n + c[..., 2:3]
        s = 0.5 * x / np.sqrt(1.0 + x * x) + 0.5
        return s @ rgb_tbl

    def residual(c, target_lab):
        return target_lab - _lab(spectrum_rgb(c), rgb_to_xyz, wp)

    def fit(c, rgb):
        target = _lab(rgb, rgb_to_xyz, wp)
        for _ in range(15):
            r = residual(c, target)
            J = np.empty(c.shape[:-1] + (3, 3))
            for k in range(3):
                e = np.zeros(3); e[k] = 1e-4
                J[..., :, k] = (residual(c + e, target) - residual(c - e, target)) / 2e-4
            try:
                step = np.linalg.solve(J, r[..., None])[..., 0]
            except np.linalg.LinAlgError:
                step = np.einsum("...ij,...j->...i", np.linalg.pinv(J), r)
            c = c - step
            m = np.abs(c).max(axis=-1, keepdims=True)
            c = np.where(m > 200.0, c * 200.0 / np.maximum(m, 1e-30), c)
        return c

    c0, c1 = 360.0, 1.0 / (830.0 - 360.0)
    for l in range(3):
        start = res // 5
        for ks in (range(start, res), range(start, -1, -1)):
            c = np.zeros((res, res, 3))
            for k in ks:
                b = scale[k]
                rgb = np.empty((res, res, 3))
                rgb[..., l] = b; rgb[..., (l + 1) % 3] = X * b; rgb[..., (l + 2) % 3] = Y * b
                c = fit(c, rgb)
                A, B, Cc = c[..., 0], c[..., 1], c[..., 2]
                data[l, k, :, :, 0] = A * c1 * c1
                data[l, k, :, :, 1] = B * c1 - 2 * A * c0 * c1 * c1
                data[l, k, :, :, 2] = Cc - B * c0 * c1 + A * (c0 * c1) ** 2
    out = (scale.astype(np.float32), np.ascontiguousarray(data, dtype=np.float32))
    _CACHE[key] = out
    return out
