"""Host-side scene assembly: the Python stand-in for what shimmer's Rust host does between
`parse_files` and `integrator.render` (src/render.rs:8-55, src/loading/scene.rs:381-907).

In the real integration these objects already exist in Rust (camera, film, sensor, spectra,
lights, materials, BvhAggregate) and the `render_gpu` shim only FLATTENS them into
`SgSceneDesc`.  This module builds the same flattened arrays for the
synthetic benchmark scenes so tests and bench.py can drive the C ABI without a Rust
toolchain.  Everything here is host preparation; nothing on the render hot path.
"""
import ctypes as C
import math
import os

import numpy as np

from . import ffi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "spectra.npz")
LAMBDA_MIN, LAMBDA_MAX = 360, 830           # spectra/spectrum.rs:24-26
CIE_Y_INTEGRAL = np.float32(106.856895)     # spectra/cie.rs:11
f32 = np.float32


# ----------------------------------------------------------------------------------------------
# 4x4 transforms (src/transform.rs).  Matrices are built in f64 and rounded once to f32: they are
# INPUTS to both the CUDA path and the oracle, so only their values matter, not how they are made.
# ----------------------------------------------------------------------------------------------
class Transform:
    def __init__(self, m, m_inv=None):
        self.m = np.asarray(m, dtype=np.float64).reshape(4, 4)
        self.m_inv = np.linalg.inv(self.m) if m_inv is None else np.asarray(m_inv, dtype=np.float64).reshape(4, 4)

    def __mul__(self, o):            # transform.rs:356-361
        return Transform(self.m @ o.m, o.m_inv @ self.m_inv)

    def inverse(self):
        return Transform(self.m_inv, self.m)

    @staticmethod
    def identity():
        return Transform(np.eye(4), np.eye(4))

    @staticmethod
    def translate(d):                # transform.rs:94-108
        m = np.eye(4); m[:3, 3] = d
        mi = np.eye(4); mi[:3, 3] = -np.asarray(d, dtype=np.float64)
        return Transform(m, mi)

    @staticmethod
    def scale(x, y, z):              # transform.rs:110-124
        return Transform(np.diag([x, y, z, 1.0]), np.diag([1.0 / x, 1.0 / y, 1.0 / z, 1.0]))

    @staticmethod
    def rotate(theta_deg, axis):     # transform.rs:198-226
        a = np.asarray(axis, dtype=np.float64); a = a / np.linalg.norm(a)
        s, c = math.sin(math.radians(theta_deg)), math.cos(math.radians(theta_deg))
        m = np.eye(4)
        m[0, :3] = [a[0] * a[0] + (1 - a[0] * a[0]) * c, a[0] * a[1] * (1 - c) - a[2] * s, a[0] * a[2] * (1 - c) + a[1] * s]
        m[1, :3] = [a[0] * a[1] * (1 - c) + a[2] * s, a[1] * a[1] + (1 - a[1] * a[1]) * c, a[1] * a[2] * (1 - c) - a[0] * s]
        m[2, :3] = [a[0] * a[2] * (1 - c) - a[1] * s, a[1] * a[2] * (1 - c) + a[0] * s, a[2] * a[2] + (1 - a[2] * a[2]) * c]
        return Transform(m, m.T)

    @staticmethod
    def look_at(pos, look, up):      # transform.rs:272-307 -> camera_from_world
        pos, look, up = (np.asarray(v, dtype=np.float64) for v in (pos, look, up))
        d = look - pos; d /= np.linalg.norm(d)
        right = np.cross(up / np.linalg.norm(up), d); right /= np.linalg.norm(right)
        new_up = np.cross(d, right)
        wfc = np.eye(4)
        wfc[:3, 0], wfc[:3, 1], wfc[:3, 2], wfc[:3, 3] = right, new_up, d, pos
        return Transform(np.linalg.inv(wfc), wfc)

    @staticmethod
    def perspective(fov_deg, n, f):  # transform.rs:309-321
        persp = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, f / (f - n), -f * n / (f - n)], [0, 0, 1, 0]], dtype=np.float64)
        inv_tan = 1.0 / math.tan(math.radians(fov_deg) / 2.0)
        return Transform.scale(inv_tan, inv_tan, 1.0) * Transform(persp)

    def swaps_handedness(self):      # transform.rs:331-340
        return np.linalg.det(self.m[:3, :3]) < 0.0

    def m32(self):
        return self.m.astype(np.float32)

    def apply_points_f32(self, p):
        """apply_point_helper (transform.rs:753-767) evaluated in f32, left to right, unfused."""
        m = self.m32(); p = np.asarray(p, dtype=np.float32)
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        out = np.empty_like(p)
        for r in range(3):
            out[:, r] = ((m[r, 0] * x + m[r, 1] * y) + m[r, 2] * z) + m[r, 3]
        wp = ((m[3, 0] * x + m[3, 1] * y) + m[3, 2] * z) + m[3, 3]
        if not np.all(wp == 1.0):
            out = out / wp[:, None]
        return out

    def apply_normals_f32(self, n):
        """apply_normal_helper (transform.rs:779-786): transpose of m_inv."""
        mi = self.m_inv.astype(np.float32); n = np.asarray(n, dtype=np.float32)
        x, y, z = n[:, 0], n[:, 1], n[:, 2]
        out = np.empty_like(n)
        for r in range(3):
            out[:, r] = (mi[0, r] * x + mi[1, r] * y) + mi[2, r] * z
        return out


# ----------------------------------------------------------------------------------------------
# Spectra (src/spectra/*.rs).  All sums are carried in f32 in the reference's order.
# ----------------------------------------------------------------------------------------------
_tables = None


def tables():
    global _tables
    if _tables is None:
        _tables = dict(np.load(_DATA))
    return _tables


def _pl_get(lams, vals, lam):
    """PiecewiseLinearSpectrum::get (spectrum.rs:408-423) for an array of wavelengths, f32."""
    lams = np.asarray(lams, dtype=np.float32); vals = np.asarray(vals, dtype=np.float32)
    lam = np.asarray(lam, dtype=np.float32)
    o = np.clip(np.searchsorted(lams, lam, side="right") - 1, 0, len(lams) - 2)
    t = (lam - lams[o]) / (lams[o + 1] - lams[o])
    out = vals[o] * (f32(1.0) - t) + vals[o + 1] * t
    out[(lam < lams[0]) | (lam > lams[-1])] = 0.0
    return out.astype(np.float32)


def _interleaved(samples, normalize):
    """PiecewiseLinearSpectrum::from_interleaved (spectrum.rs:324-370)."""
    s = np.asarray(samples, dtype=np.float32)
    lam, v = list(s[0::2]), list(s[1::2])
    if lam[0] > LAMBDA_MIN:
        lam.insert(0, f32(LAMBDA_MIN - 1)); v.insert(0, v[0])
    if lam[-1] < LAMBDA_MAX:
        lam.append(f32(LAMBDA_MAX + 1)); v.append(v[-1])
    lam, v = np.asarray(lam, dtype=np.float32), np.asarray(v, dtype=np.float32)
    if normalize:
        ip = f32(0.0)
        y = cie("Y")
        g = _pl_get(lam, v, np.arange(LAMBDA_MIN, LAMBDA_MAX + 1, dtype=np.float32))
        for i in range(LAMBDA_MAX - LAMBDA_MIN + 1):
            ip = f32(ip + f32(g[i] * y[i]))
        v = (v * f32(CIE_Y_INTEGRAL / ip)).astype(np.float32)
    return lam, v


def cie(which):
    """Dense 360..830 CIE matching function (spectra/cie.rs:19-32: PiecewiseLinear over CIE_LAMBDA, densely sampled)."""
    t = tables()
    return _pl_get(t["CIE_LAMBDA"], t["CIE_" + which], np.arange(LAMBDA_MIN, LAMBDA_MAX + 1, dtype=np.float32))


def named_spectrum(name):
    """Spectrum::get_named_spectrum (spectrum.rs:111-133) -> ('pl', lambdas, values)."""
    t = tables()
    key, norm = {
        "stdillum-D65": ("CIE_ILLUM_D6500", True), "illum-acesD60": ("ACES_ILLUM_D60", True),
        "glass-BK7": ("GLASS_BK7_ETA_SAMPLES", False), "glass-BAF10": ("GLASS_BAF10_ETA_SAMPLES", False),
        "glass-F11": ("GLASS_F11_ETA_SAMPLES", False),
        "metal-Cu-eta": ("CU_ETA_SAMPLES", False), "metal-Cu-k": ("CU_K_SAMPLES", False),
        "metal-Au-eta": ("AU_ETA_SAMPLES", False), "metal-Au-k": ("AU_K_SAMPLES", False),
        "metal-Ag-eta": ("AG_ETA_SAMPLES", False), "metal-Ag-k": ("AG_K_SAMPLES", False),
        "metal-Al-eta": ("AL_ETA_SAMPLES", False), "metal-Al-k": ("AL_K_SAMPLES", False),
    }[name]
    lam, v = _interleaved(t[key], norm)
    return ("pl", lam, v)


def spectrum_dense(spec):
    """DenselySampledSpectrum::new (spectrum.rs:179-197): spectrum.get(lambda) for lambda in 360..=830."""
    lam = np.arange(LAMBDA_MIN, LAMBDA_MAX + 1, dtype=np.float32)
    kind = spec[0]
    if kind == "const":
        return np.full(lam.shape, f32(spec[1]), dtype=np.float32)
    if kind == "pl":
        return _pl_get(spec[1], spec[2], lam)
    if kind == "dense":
        return np.asarray(spec[1], dtype=np.float32)
    raise ValueError(kind)


def spectrum_to_photometric(spec):
    """spectrum.rs:617-631, f32 running sum."""
    d, y = spectrum_dense(spec), cie("Y")
    acc = f32(0.0)
    for i in range(len(d)):
        acc = f32(acc + f32(y[i] * d[i]))
    return acc


def srgb_output_matrix():
    """RgbColorSpace::new for sRGB (colorspace.rs:26-62,133-141) -> rgb_from_xyz; the cie1931 sensor
    without white balance has xyz_from_sensor_rgb = identity (film.rs:825-844), so this is
    output_rgb_from_sensor_rgb."""
    ill = spectrum_dense(named_spectrum("stdillum-D65")).astype(np.float64)
    X, Y, Z = (cie(c).astype(np.float64) for c in "XYZ")
    w = np.array([np.sum(X * ill), np.sum(Y * ill), np.sum(Z * ill)]) / float(CIE_Y_INTEGRAL)

    def xyY(xy):
        return np.array([xy[0] / xy[1], 1.0, (1.0 - xy[0] - xy[1]) / xy[1]])
    rgb = np.stack([xyY((0.64, 0.33)), xyY((0.3, 0.6)), xyY((0.15, 0.06))], axis=1)
    c = np.linalg.inv(rgb) @ w
    xyz_from_rgb = rgb @ np.diag(c)
    return np.linalg.inv(xyz_from_rgb).astype(np.float32)


def _equal_area_square_to_sphere(px, py):
    """math.rs:453-484 as written (`vp - up / r + 1.0`), f32; host-side use only (the `illuminance` scale, light.rs:190-207)."""
    u, v = f32(2.0) * f32(px) - f32(1.0), f32(2.0) * f32(py) - f32(1.0)
    up, vp = abs(u), abs(v)
    sd = f32(1.0) - (up + vp)
    r = f32(1.0) - abs(sd)
    phi = (f32(1.0) if r == 0.0 else f32(f32(vp - f32(up / r)) + f32(1.0))) * f32(np.pi) / f32(4.0)
    z = f32(math.copysign(f32(1.0) - r * r, sd))
    c, s_ = f32(math.copysign(np.cos(phi, dtype=np.float32), u)), f32(math.copysign(np.sin(phi, dtype=np.float32), v))
    k = f32(np.sqrt(max(f32(0.0), f32(2.0) - r * r), dtype=np.float32))
    return np.array([c * r * k, s_ * r * k, z], np.float32)


def bilinear_patch_is_rectangle(P4):
    """BilinearPatch::is_rectangle (bilinear_patch.rs:108-143); P4 = p00, p10, p01, p11 (f32)."""
    p00, p10, p01, p11 = (np.asarray(x, np.float32) for x in P4)
    if (p00 == p01).all() or (p01 == p11).all() or (p11 == p10).all() or (p10 == p00).all():
        return False
    n = np.cross(p10 - p00, p01 - p00).astype(np.float32); n /= np.linalg.norm(n)
    d = (p11 - p00); d = d / np.linalg.norm(d)
    if abs(float(np.dot(d, n))) > 1e-5:
        return False
    pc = (p00 + p01 + p10 + p11) * f32(0.25)
    d2 = [float(np.sum((q - pc) ** 2)) for q in (p00, p01, p10, p11)]
    return all(abs(d2[i] - d2[0]) / d2[0] <= 1e-4 for i in range(1, 4))


def bilinear_patch_area(P4):
    """BilinearPatch::new (bilinear_patch.rs:40-76): exact for rectangles, a 3x3 quad approximation otherwise.  Only `phi()` reads it."""
    p00, p10, p01, p11 = (np.asarray(x, np.float32) for x in P4)
    if bilinear_patch_is_rectangle(P4):
        return f32(np.linalg.norm(p00 - p01)) * f32(np.linalg.norm(p00 - p10))
    NA = 3
    lerp = lambda t, a, b: a * f32(1.0 - t) + b * f32(t)
    g = [[lerp(i / NA, lerp(j / NA, p00, p01), lerp(j / NA, p10, p11)) for j in range(NA + 1)] for i in range(NA + 1)]
    area = f32(0.0)
    for i in range(NA):
        for j in range(NA):
            area = f32(area + f32(0.5) * f32(np.linalg.norm(np.cross(g[i + 1][j + 1] - g[i][j], g[i + 1][j] - g[i][j + 1]))))
    return area


def piecewise_constant_1d(f, lo=0.0, hi=1.0):
    """PiecewiseConstant1D::new_bounded (sampling.rs:27-65), f32 running sums in the reference's order -> (func, cdf, func_int)."""
    func = np.abs(np.asarray(f, dtype=np.float32))
    n = len(func)
    step = f32(hi) - f32(lo)
    terms = (func * step / f32(n)).astype(np.float32)          # func[i-1] * (max - min) / n
    cdf = np.zeros(n + 1, np.float32)
    cdf[1:] = np.cumsum(terms, dtype=np.float32)               # sequential f32 adds (numpy's cumsum is a running sum)
    func_int = f32(cdf[n])
    if func_int == 0.0:
        cdf[1:] = (np.arange(1, n + 1, dtype=np.float32) / f32(n)).astype(np.float32)
    else:
        cdf[1:] = (cdf[1:] / func_int).astype(np.float32)
    return func, cdf, func_int


def piecewise_constant_2d(func2d):
    """PiecewiseConstant2D::new over [0,1]^2 (sampling.rs:120-151) -> (func[nv,nu], cdf[nv,nu+1], marg_func[nv], marg_cdf[nv+1], marg_int)."""
    func2d = np.asarray(func2d, dtype=np.float32)
    nv, nu = func2d.shape
    fn = np.empty((nv, nu), np.float32); cdf = np.empty((nv, nu + 1), np.float32); ints = np.empty(nv, np.float32)
    for v in range(nv):
        fn[v], cdf[v], ints[v] = piecewise_constant_1d(func2d[v])
    mf, mcdf, mint = piecewise_constant_1d(ints)
    return fn, cdf, mf, mcdf, mint


# ----------------------------------------------------------------------------------------------
# Scene builder
# ----------------------------------------------------------------------------------------------
class SceneDesc:
    """Owns the numpy arrays behind an SgSceneDesc (keeps them alive for the C call)."""

    def __init__(self):
        self.arrays = {}
        self.desc = ffi.SgSceneDesc()
        self.meta = {}

    def ptr(self):
        return C.byref(self.desc)


def _as_ptr(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


class SceneBuilder:
    def __init__(self, rendering_space="camera-world"):
        self.rendering_space = rendering_space     # main.rs:66-67 default
        self.spectra = []        # list of tuples
        self._spec_cache = {}
        self.materials = []
        self.meshes = []         # dicts
        self.extra_lights = []   # non-area lights in add order (scene.rs add_light)
        self.textures = []       # dicts: levels (list of HxWxC f32 arrays) + SgTexture parameters
        self.n_objects = 0       # object definitions (ObjectBegin/End); meshes carry an `object` id or None
        self.instances = []      # (object id, render_from_instance Transform)
        self.instance_ctms = []  # world_from_instance of each instance, as given
        self.spheres = []        # dicts: Sphere::new fields + material (top-level shapes, after the meshes)
        self.patch_meshes = []   # dicts: BilinearPatchMesh (4 vertex indices per patch); top-level shapes, after the triangle meshes
        self.env_maps = []       # dicts: ImageInfinitelight images + render_from_light (light.rs:805-981)
        self.mappings = []       # SgTextureMapping rows (spherical / cylindrical / planar, texture.rs:938-1035)
        self.fix_instancing = False
        self.camera = None
        self.film = None
        self.world_from_camera = None

    # -- spectra / materials ---------------------------------------------------------------
    def spectrum(self, spec, key=None):
        if key is not None and key in self._spec_cache:
            return self._spec_cache[key]
        self.spectra.append(spec)
        if key is not None:
            self._spec_cache[key] = len(self.spectra) - 1
        return len(self.spectra) - 1

    def image_texture(self, image, filter="bilinear", wrap="repeat", max_anisotropy=8.0, scale=1.0, invert=False,
                      su=1.0, sv=1.0, du=0.0, dv=0.0, spectrum_type="albedo", mapping=None, levels=None):
        """ImageTextureBase::new (texture.rs:283-330) + MIPMap::new -> Image::generate_pyramid (image.rs:699-787).
        `image`: H x W (one channel) or H x W x 3 array of LINEAR values, row 0 = top of the image (what
        Image::get_channel returns after colour decoding).  Power-of-two sizes only: the reference first
        resamples other sizes with a Lanczos filter (image.rs float_resize_up), which this host stand-in omits."""
        if levels is not None:                   # a finished pyramid, e.g. from shimmer_b200.generate_pyramid (sg_image_generate_pyramid)
            levels = [np.ascontiguousarray(l, np.float32).reshape(l.shape[0], l.shape[1], -1) for l in levels]
            img = levels[0]
        else:
            img = np.asarray(image, dtype=np.float32)
            if img.ndim == 2:
                img = img[:, :, None]
        H, W, Cn = img.shape
        if Cn not in (1, 3) or (levels is None and ((W & (W - 1)) or (H & (H - 1)))):
            raise ValueError("image textures must be one- or three-channel; other resolutions than powers of two need levels= "
                             "(shimmer_b200.generate_pyramid resamples them on the device like image.rs float_resize_up)")
        levels = [img] if levels is None else levels
        while levels[-1].shape[0] > 1 or levels[-1].shape[1] > 1:           # 2x2 box filter, image.rs:733-768
            a = levels[-1]
            h, w = a.shape[:2]
            y1 = np.arange(0, h, 2) + (1 if h > 1 else 0); x1 = np.arange(0, w, 2) + (1 if w > 1 else 0)
            y0 = np.arange(0, h, 2); x0 = np.arange(0, w, 2)
            nxt = f32(0.25) * (((a[y0][:, x0] + a[y0][:, x1]) + a[y1][:, x0]) + a[y1][:, x1])
            levels.append(nxt.astype(np.float32))
        self.textures.append(dict(levels=levels, n_channels=Cn,
                                  filter={"point": 0, "bilinear": 1, "trilinear": 2, "ewa": 3, "EWA": 3}[filter],
                                  wrap={"repeat": 0, "black": 1, "clamp": 2}[wrap], max_anisotropy=max_anisotropy, scale=scale,
                                  invert=1 if invert else 0, su=su, sv=sv, du=du, dv=dv,
                                  spectrum_type={"albedo": 0, "unbounded": 1}[spectrum_type],
                                  mapping=-1 if mapping is None else mapping))
        return len(self.textures) - 1

    # -- the non-image members of FloatTexture / SpectrumTexture (texture.rs:88-94,411-417) --------------------
    def _node_texture(self, kind, is_spectrum, **node):
        self.textures.append(dict(kind=kind, n_channels=3 if is_spectrum else 1, node=node))
        return len(self.textures) - 1

    def _is_spectrum_texture(self, tex):
        return self.textures[tex]["n_channels"] != 1

    def constant_texture(self, value=None, spectrum=None):
        """FloatConstantTexture (texture.rs:154-179; `value`) or SpectrumConstantTexture (:485-535; `spectrum` = spectrum id)."""
        if spectrum is not None:
            return self._node_texture(ffi.SG_TEXTURE_CONSTANT, True, spectrum=int(spectrum))
        return self._node_texture(ffi.SG_TEXTURE_CONSTANT, False, value=float(value))

    def scaled_texture(self, tex, scale):
        """Float/SpectrumScaledTexture (texture.rs:180-213,:537-583): `scale` is a float texture id."""
        return self._node_texture(ffi.SG_TEXTURE_SCALED, self._is_spectrum_texture(tex), tex1=tex, tex2=scale)

    def mix_texture(self, tex1, tex2, amount):
        """Float/SpectrumMixTexture (texture.rs:215-262,:585-651): `amount` is a float texture id."""
        return self._node_texture(ffi.SG_TEXTURE_MIX, self._is_spectrum_texture(tex1) or self._is_spectrum_texture(tex2),
                                  tex1=tex1, tex2=tex2, amount=amount)

    def direction_mix_texture(self, tex1, tex2, dir=(0.0, 1.0, 0.0)):
        """Float/SpectrumDirectionMixTexture (texture.rs:264-310,:653-826): amount = dot(n, dir), `dir` in render space as given."""
        return self._node_texture(ffi.SG_TEXTURE_DIRECTION_MIX, self._is_spectrum_texture(tex1) or self._is_spectrum_texture(tex2),
                                  tex1=tex1, tex2=tex2, dir=tuple(float(x) for x in dir))

    def set_material_textures(self, material, **tex):
        """Texture-valued parameters of `material` (SgMaterialTextures): u_roughness, v_roughness, spec_a (conductor eta / coated
        conductor.eta or reflectance), spec_b (conductor k / coated albedo), spec_d (coated conductor.k), thickness, g, u_roughness2,
        v_roughness2 = texture ids.  What material.rs reads through tex_eval.evaluate_float / evaluate_spectrum."""
        bad = set(tex) - set(ffi.SgMaterialTextures.NAMES)
        if bad:
            raise ValueError("unknown material texture parameter(s): %s" % sorted(bad))
        self.materials[material].setdefault("param_textures", {}).update({k: int(v) for k, v in tex.items()})
        return material

    def texture_mapping(self, kind, texture_from_world=None, v1=(1.0, 0.0, 0.0), v2=(0.0, 1.0, 0.0), udelta=0.0, vdelta=0.0):
        """TextureMapping2D::create (texture.rs:853-893) for "spherical" / "cylindrical" / "planar":
        texture_from_render = (render_from_texture)^-1 with render_from_texture = render_from_world * CTM; the planar
        mapping keeps `v1`, `v2`, `udelta`, `vdelta`.  `texture_from_world` is the INVERSE CTM (a Transform) or None."""
        tfw = texture_from_world if texture_from_world is not None else Transform.identity()
        tfr = tfw * self.render_from_world.inverse()
        self.mappings.append(dict(kind={"spherical": ffi.SG_MAPPING_SPHERICAL, "cylindrical": ffi.SG_MAPPING_CYLINDRICAL,
                                        "planar": ffi.SG_MAPPING_PLANAR}[kind], tfr=tfr, vs=v1, vt=v2, ds=udelta, dt=vdelta))
        return len(self.mappings) - 1

    def diffuse(self, reflectance, reflectance_tex=None, displacement_tex=None):
        """DiffuseMaterial::create (material.rs:259-283): displacement is ALWAYS Some(0.0) unless a texture is given.
        `reflectance_tex` / `displacement_tex`: ids from image_texture()."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_DIFFUSE, spec_a=self.spectrum(reflectance), spec_b=-1,
                                   flags=ffi.SG_MAT_HAS_DISPLACEMENT | ffi.SG_MAT_REMAP_ROUGHNESS, ur=0.0, vr=0.0,
                                   tex_reflectance=-1 if reflectance_tex is None else reflectance_tex,
                                   tex_displacement=-1 if displacement_tex is None else displacement_tex))
        return len(self.materials) - 1

    def conductor(self, eta, k, roughness=0.0, remap=True, normal_map=None):
        """ConductorMaterial::create (material.rs:362-431), eta/k form.  `normal_map`: three-channel image_texture() id."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_CONDUCTOR, spec_a=self.spectrum(eta), spec_b=self.spectrum(k),
                                   flags=(ffi.SG_MAT_REMAP_ROUGHNESS if remap else 0), ur=roughness, vr=roughness,
                                   normal_map=-1 if normal_map is None else normal_map))
        return len(self.materials) - 1

    def dielectric(self, eta, roughness=0.0, remap=True, normal_map=None):
        """DielectricMaterial::create (material.rs:538-581)."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_DIELECTRIC, spec_a=self.spectrum(eta), spec_b=-1,
                                   flags=(ffi.SG_MAT_REMAP_ROUGHNESS if remap else 0), ur=roughness, vr=roughness,
                                   normal_map=-1 if normal_map is None else normal_map))
        return len(self.materials) - 1

    def coated_conductor(self, conductor_eta=None, conductor_k=None, reflectance=None, interface_eta=("const", 1.5), interface_roughness=0.0,
                         conductor_roughness=0.0, thickness=0.01, albedo=("const", 0.0), g=0.0, max_depth=10, n_samples=1, remap=True,
                         normal_map=None):
        """CoatedConductorMaterial::create (material.rs:1030-1180): conductor given by eta + k spectra (defaults metal-Cu) or
        by `reflectance`."""
        if reflectance is None and conductor_eta is None:
            conductor_eta, conductor_k = named_spectrum("metal-Cu-eta"), named_spectrum("metal-Cu-k")
        fl = (ffi.SG_MAT_REMAP_ROUGHNESS if remap else 0) | (ffi.SG_MAT_CONDUCTOR_REFLECTANCE if reflectance is not None else 0)
        self.materials.append(dict(kind=ffi.SG_MATERIAL_COATED_CONDUCTOR,
                                   spec_a=self.spectrum(reflectance if reflectance is not None else conductor_eta),
                                   spec_d=-1 if reflectance is not None else self.spectrum(conductor_k),
                                   spec_b=self.spectrum(albedo), spec_c=self.spectrum(interface_eta), flags=fl,
                                   ur=interface_roughness, vr=interface_roughness, ur2=conductor_roughness, vr2=conductor_roughness,
                                   thickness=thickness, g=g, max_depth=max_depth, n_samples=n_samples,
                                   normal_map=-1 if normal_map is None else normal_map))
        return len(self.materials) - 1

    def mix(self, material_a, material_b, amount=0.5, amount_tex=None):
        """MixMaterial::create (material.rs:1296-1307): `amount` float (0.5) or a one-channel image texture."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_MIX, spec_a=-1, spec_b=-1, flags=0, ur=0.0, vr=0.0,
                                   mix_materials=(material_a, material_b), mix_amount=amount,
                                   tex_mix_amount=-1 if amount_tex is None else amount_tex))
        return len(self.materials) - 1

    def thin_dielectric(self, eta):
        """ThinDielectricMaterial::create (material.rs:666-700): `eta` spectrum only."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_THIN_DIELECTRIC, spec_a=self.spectrum(eta), spec_b=-1, flags=0, ur=0.0, vr=0.0))
        return len(self.materials) - 1

    def coated_diffuse(self, reflectance, eta=("const", 1.5), roughness=0.0, thickness=0.01, albedo=("const", 0.0), g=0.0,
                       max_depth=10, n_samples=1, remap=True, reflectance_tex=None, displacement_tex=None):
        """CoatedDiffuseMaterial::create (material.rs:820-903); displacement defaults to None."""
        self.materials.append(dict(kind=ffi.SG_MATERIAL_COATED_DIFFUSE, spec_a=self.spectrum(reflectance), spec_b=self.spectrum(albedo),
                                   spec_c=self.spectrum(eta), flags=(ffi.SG_MAT_REMAP_ROUGHNESS if remap else 0), ur=roughness, vr=roughness,
                                   thickness=thickness, g=g, max_depth=max_depth, n_samples=n_samples,
                                   tex_reflectance=-1 if reflectance_tex is None else reflectance_tex,
                                   tex_displacement=-1 if displacement_tex is None else displacement_tex))
        if displacement_tex is not None:
            self.materials[-1]["flags"] |= ffi.SG_MAT_HAS_DISPLACEMENT
        return len(self.materials) - 1

    # -- camera / film -----------------------------------------------------------------------
    def set_camera(self, pos, look, up, fov, resolution, lens_radius=0.0, focal_distance=1e6, crop=None, kind="perspective",
                   screen_window=None):
        """PerspectiveCamera::create/new (camera.rs:839-963) + CameraTransform::new (:506-523) +
        ProjectiveCameraBase::new (:595-642)."""
        W, H = resolution
        self.camera_args = dict(pos=tuple(pos), look=tuple(look), up=tuple(up), fov=fov, resolution=(W, H), lens_radius=lens_radius,
                                focal_distance=focal_distance, crop=crop, kind=kind, screen_window=screen_window)
        camera_from_world = Transform.look_at(pos, look, up)
        world_from_camera = camera_from_world.inverse()
        if self.rendering_space == "camera-world":
            p_cam = world_from_camera.m[:3, 3]
            world_from_render = Transform.translate(p_cam)
        elif self.rendering_space == "world":
            world_from_render = Transform.identity()
        else:
            world_from_render = world_from_camera
        self.render_from_world = world_from_render.inverse()
        render_from_camera = self.render_from_world * world_from_camera
        frame = W / H
        screen = (-frame, frame, -1.0, 1.0) if frame > 1.0 else (-1.0, 1.0, -1.0 / frame, 1.0 / frame)
        if screen_window is not None:                       # `screenwindow` (camera.rs:691-707)
            screen = (screen_window[0], screen_window[1], screen_window[2], screen_window[3])
        # OrthographicCamera::new: Transform::orthographic(0, 1) = identity (camera.rs:718, transform.rs:125-128)
        screen_from_camera = Transform.identity() if kind == "orthographic" else Transform.perspective(fov, 1e-2, 1000.0)
        ndc_from_screen = Transform.scale(1.0 / (screen[1] - screen[0]), 1.0 / (screen[3] - screen[2]), 1.0) * \
            Transform.translate((-screen[0], -screen[3], 0.0))
        raster_from_ndc = Transform.scale(W, -H, 1.0)
        raster_from_screen = raster_from_ndc * ndc_from_screen
        camera_from_raster = screen_from_camera.inverse() * raster_from_screen.inverse()
        cfr = camera_from_raster

        def app(p):
            v = cfr.m @ np.array([p[0], p[1], p[2], 1.0]); return v[:3] / v[3]
        dx_camera = app((1, 0, 0)) - app((0, 0, 0))
        dy_camera = app((0, 1, 0)) - app((0, 0, 0))
        cam = ffi.SgCamera()
        cam.camera_from_raster[:] = camera_from_raster.m32().ravel().tolist()
        cam.render_from_camera[:] = render_from_camera.m32().ravel().tolist()
        cam.camera_from_render[:] = render_from_camera.m_inv.astype(np.float32).ravel().tolist()
        cam.dx_camera[:] = dx_camera.astype(np.float32).tolist()
        cam.dy_camera[:] = dy_camera.astype(np.float32).tolist()
        cam.lens_radius = lens_radius; cam.focal_distance = focal_distance
        cam.shutter_open = 0.0; cam.shutter_close = 1.0
        if kind == "orthographic":                          # camera.rs:727-735: vector transforms of X / Y, fixed differentials
            cam.kind = ffi.SG_CAMERA_ORTHOGRAPHIC
            m32 = camera_from_raster.m32()
            cam.dx_camera[:] = m32[:3, 0].tolist(); cam.dy_camera[:] = m32[:3, 1].tolist()
            cam.min_pos_differential_x[:] = m32[:3, 0].tolist(); cam.min_pos_differential_y[:] = m32[:3, 1].tolist()
            cam.min_dir_differential_x[:] = [0.0, 0.0, 0.0]; cam.min_dir_differential_y[:] = [0.0, 0.0, 0.0]
        else:
            cam.kind = ffi.SG_CAMERA_PERSPECTIVE
            self._find_minimum_differentials(cam, W, H)
        self.camera = cam
        film = ffi.SgFilm()
        film.full_resolution[:] = [W, H]
        film.pixel_bounds[:] = list(crop) if crop else [0, 0, W, H]
        film.filter_radius[:] = [0.5, 0.5]                  # BoxFilter default radius (filter.rs:64-97)
        film.imaging_ratio = 1.0                            # exposure_time * iso / 100 (film.rs:787)
        film.max_component_value = float("inf")             # `maxcomponentvalue` default (film.rs RgbFilm::create)
        film.output_rgb_from_sensor_rgb[:] = srgb_output_matrix().ravel().tolist()
        self.film = film

    @staticmethod
    def _find_minimum_differentials(cam, W, H):
        """CameraBase::find_minimum_differentials (camera.rs:356-430): 512 film samples along the diagonal with
        p_lens = (0.5, 0.5) (centre of the lens -> the thin-lens and pinhole branches of generate_ray_differential
        coincide).  Host-side camera construction; f32 arithmetic like the reference."""
        cfr = np.array(cam.camera_from_raster[:], np.float32).reshape(4, 4)
        rfc = np.array(cam.render_from_camera[:], np.float32).reshape(4, 4)
        cfrn = np.array(cam.camera_from_render[:], np.float32).reshape(4, 4)
        dxc, dyc = np.array(cam.dx_camera[:], np.float32), np.array(cam.dy_camera[:], np.float32)

        def nrm(v):
            return (v / np.sqrt(np.sum(v * v, dtype=np.float32), dtype=np.float32)).astype(np.float32)

        def length(v):
            return float(np.sqrt(np.sum(v * v, dtype=np.float32)))
        best = [np.full(3, np.inf, np.float32) for _ in range(4)]
        n = 512
        for i in range(n):
            pf = np.array([f32(i) / f32(n - 1) * f32(W), f32(i) / f32(n - 1) * f32(H), 0.0, 1.0], np.float32)
            pc = cfr @ pf
            pc = (pc[:3] / pc[3]).astype(np.float32) if pc[3] != 1.0 else pc[:3]
            d = rfc[:3, :3] @ nrm(pc)
            rxd = rfc[:3, :3] @ nrm(pc + dxc); ryd = rfc[:3, :3] @ nrm(pc + dyc)
            dox = cfrn[:3, :3] @ np.zeros(3, np.float32); doy = dox          # rx_origin == ry_origin == o for a perspective camera
            if length(dox) < length(best[0]): best[0] = dox
            if length(doy) < length(best[1]): best[1] = doy
            d, rxd, ryd = nrm(d), nrm(rxd), nrm(ryd)
            sign = f32(math.copysign(1.0, d[2])); a = f32(-1.0) / (sign + d[2]); b = d[0] * d[1] * a      # Frame::from_z -> coordinate_system
            fx = np.array([f32(1.0) + sign * d[0] * d[0] * a, sign * b, -sign * d[0]], np.float32)
            fy = np.array([b, sign + d[1] * d[1] * a, -d[1]], np.float32)
            F = np.stack([fx, fy, d]).astype(np.float32)
            df = F @ d; dxf = nrm(F @ rxd); dyf = nrm(F @ ryd)
            if length(dxf - df) < length(best[2]): best[2] = (dxf - df).astype(np.float32)
            if length(dyf - df) < length(best[3]): best[3] = (dyf - df).astype(np.float32)
        cam.min_pos_differential_x[:] = best[0].tolist(); cam.min_pos_differential_y[:] = best[1].tolist()
        cam.min_dir_differential_x[:] = best[2].tolist(); cam.min_dir_differential_y[:] = best[3].tolist()

    # -- geometry ------------------------------------------------------------------------------
    def begin_object(self):
        """ObjectBegin: returns an object-definition id for add_mesh(object=...) / add_instance."""
        self.n_objects += 1
        return self.n_objects - 1

    def add_instance(self, obj, world_from_instance=None):
        """ObjectInstance: render_from_instance = render_from_object * world_from_render (scene.rs:2002)."""
        ctm = world_from_instance if world_from_instance is not None else Transform.identity()
        self.instances.append((obj, self.render_from_world * ctm * self.render_from_world.inverse()))
        self.instance_ctms.append(ctm)                      # the CTM at the ObjectInstance directive (pbrt_export.py)

    def add_mesh(self, p, indices, material, n=None, uv=None, area_light=None, reverse_orientation=False,
                 object_from_world=None, object=None):
        """Shape "trianglemesh" (triangle.rs:56-127 + mesh.rs:22-94).  `area_light` =
        dict(L=spectrum tuple, scale=float, two_sided=bool) -> one DiffuseAreaLight per triangle
        (scene.rs:609-622)."""
        ctm = object_from_world if object_from_world is not None else Transform.identity()
        rfo = self.render_from_world * ctm
        src = dict(p=np.asarray(p, dtype=np.float32).reshape(-1, 3).copy(), n=None if n is None else np.asarray(n, dtype=np.float32).reshape(-1, 3).copy(),
                   ctm=object_from_world, reverse_orientation=reverse_orientation)       # what the scene file holds (pbrt_export.py)
        p = rfo.apply_points_f32(np.asarray(p, dtype=np.float32).reshape(-1, 3))
        if n is not None:
            n = rfo.apply_normals_f32(np.asarray(n, dtype=np.float32).reshape(-1, 3))
            if reverse_orientation:
                n = -n
        flags = 0
        if n is not None: flags |= ffi.SG_MESH_HAS_N
        if uv is not None: flags |= ffi.SG_MESH_HAS_UV
        if reverse_orientation: flags |= ffi.SG_MESH_REVERSE_ORIENTATION
        if rfo.swaps_handedness(): flags |= ffi.SG_MESH_SWAPS_HANDEDNESS
        self.meshes.append(dict(p=p, idx=np.asarray(indices, dtype=np.uint32).reshape(-1, 3), n=n,
                                uv=None if uv is None else np.asarray(uv, dtype=np.float32).reshape(-1, 2),
                                flags=flags, material=material, area_light=area_light, object=object, src=src))
        if object is not None and area_light is not None:
            raise ValueError("area lights are not supported inside object definitions")
        return len(self.meshes) - 1

    def add_bilinear_mesh(self, p, indices, material, n=None, uv=None, reverse_orientation=False, object_from_world=None, area_light=None,
                          object=None):
        """Shape "bilinearmesh": BilinearPatchMesh::new (shape/mesh.rs:111-175) + one BilinearPatch per four indices
        (p00, p10, p01, p11; bilinear_patch.rs:77-106).  Top-level patches only.  `area_light` = dict(L=spectrum tuple, scale=float,
        two_sided=bool) -> one DiffuseAreaLight per patch (scene.rs:609-622)."""
        ctm = object_from_world if object_from_world is not None else Transform.identity()
        rfo = self.render_from_world * ctm
        p = rfo.apply_points_f32(np.asarray(p, dtype=np.float32).reshape(-1, 3))
        if n is not None:
            n = rfo.apply_normals_f32(np.asarray(n, dtype=np.float32).reshape(-1, 3))
            if reverse_orientation:
                n = -n
        flags = ffi.SG_MESH_BILINEAR
        if n is not None: flags |= ffi.SG_MESH_HAS_N
        if uv is not None: flags |= ffi.SG_MESH_HAS_UV
        if reverse_orientation: flags |= ffi.SG_MESH_REVERSE_ORIENTATION
        if rfo.swaps_handedness(): flags |= ffi.SG_MESH_SWAPS_HANDEDNESS
        self.patch_meshes.append(dict(p=p, idx=np.asarray(indices, dtype=np.uint32).reshape(-1, 4), n=n,
                                      uv=None if uv is None else np.asarray(uv, dtype=np.float32).reshape(-1, 2), flags=flags, material=material,
                                      area_light=area_light, object=object))
        if object is not None and area_light is not None:
            raise ValueError("area lights are not supported inside object definitions")
        return len(self.patch_meshes) - 1

    def add_sphere(self, radius, material, z_min=None, z_max=None, phi_max=360.0, object_from_world=None, reverse_orientation=False,
                   area_light=None, object=None):
        """Shape "sphere": Sphere::create / Sphere::new (shape/sphere.rs:48-92).  `object_from_world` is the CTM
        (render_from_object = render_from_world * CTM).  Top-level spheres only.  `area_light` = dict(L=spectrum tuple,
        scale=float, two_sided=bool) -> one DiffuseAreaLight over the sphere (scene.rs:609-622)."""
        ctm = object_from_world if object_from_world is not None else Transform.identity()
        rfo = self.render_from_world * ctm
        r = f32(radius)
        zlo = f32(min(-radius if z_min is None else z_min, radius if z_max is None else z_max))
        zhi = f32(max(-radius if z_min is None else z_min, radius if z_max is None else z_max))
        clamp = lambda v, lo, hi: f32(min(max(v, lo), hi))
        flags = (ffi.SG_MESH_REVERSE_ORIENTATION if reverse_orientation else 0) | (ffi.SG_MESH_SWAPS_HANDEDNESS if rfo.swaps_handedness() else 0)
        self.spheres.append(dict(rfo=rfo, radius=r, z_min=clamp(zlo, -r, r), z_max=clamp(zhi, -r, r),
                                 theta_z_min=f32(np.arccos(clamp(zlo / r, f32(-1.0), f32(1.0)))),
                                 theta_z_max=f32(np.arccos(clamp(zhi / r, f32(-1.0), f32(1.0)))),
                                 phi_max=f32(f32(np.pi) / f32(180.0)) * clamp(f32(phi_max), f32(0.0), f32(360.0)),
                                 flags=flags, material=material, area_light=area_light, object=object,
                                 src=dict(radius=radius, z_min=z_min, z_max=z_max, phi_max=phi_max, ctm=object_from_world, reverse_orientation=reverse_orientation)))
        if object is not None and area_light is not None:
            raise ValueError("area lights are not supported inside object definitions")
        return len(self.spheres) - 1

    def add_point_light(self, pos, I, scale=1.0):
        """PointLight::create (light.rs:421-453)."""
        sc = f32(scale) / spectrum_to_photometric(I)
        pr = self.render_from_world.apply_points_f32(np.asarray([pos], dtype=np.float32))[0]
        self.extra_lights.append(dict(kind=ffi.SG_LIGHT_POINT, spectrum=self.spectrum(("dense", spectrum_dense(I))),
                                      scale=float(sc), pos=pr, src=dict(pos=tuple(pos), I=I, scale=scale)))

    def add_uniform_infinite_light(self, L, scale=1.0):
        """Light::create "infinite" with a constant L (light.rs:697-728)."""
        sc = f32(scale) / spectrum_to_photometric(L)
        self.extra_lights.append(dict(kind=ffi.SG_LIGHT_UNIFORM_INFINITE, spectrum=self.spectrum(("dense", spectrum_dense(L))),
                                      scale=float(sc), pos=np.zeros(3, np.float32), src=dict(L=L, scale=scale)))

    def add_image_infinite_light(self, image, scale=1.0, illuminance=None, light_from_world=None):
        """Light::create "infinite" with a `filename` (light.rs:164-232) + ImageInfinitelight::new (:916-964).  `image`: square
        N x N x 3 array of LINEAR sRGB values in the equal-area octahedral parameterisation (what Image::read + select_channels
        hands over).  `light_from_world` is the CTM (render_from_light = render_from_world * CTM)."""
        img = np.ascontiguousarray(image, dtype=np.float32)
        if img.ndim != 3 or img.shape[2] != 3 or img.shape[0] != img.shape[1]:
            raise ValueError("environment maps must be square RGB images (light.rs:934-937)")
        N = img.shape[0]
        ill = named_spectrum("stdillum-D65")                   # the sRGB colour space's illuminant (colorspace.rs:134-141)
        sc = f32(scale) / spectrum_to_photometric(ill)
        if illuminance is not None and illuminance > 0.0:      # light.rs:181-211
            lum = np.array([0.2126, 0.7152, 0.0722], np.float32)   # luminance_vector of sRGB (approximate: host-side constant)
            acc = f32(0.0)
            for y in range(N):
                v = (f32(y) + f32(0.5)) / f32(N)
                for x in range(N):
                    u = (f32(x) + f32(0.5)) / f32(N)
                    w = _equal_area_square_to_sphere(u, v)
                    if w[2] <= 0.0:
                        continue
                    for c in range(3):
                        acc = f32(acc + f32(f32(img[y, x, c] * lum[c]) * f32(w[2])))
            acc = f32(acc * f32(f32(2.0 * np.float32(np.pi)) / f32(N * N)))
            sc = f32(sc * f32(f32(illuminance) / acc))
        ctm = light_from_world if light_from_world is not None else Transform.identity()
        rfl = self.render_from_world * ctm
        # Image::get_default_sampling_distribution (image.rs:1379-1405): channel average per pixel
        d = ((img[:, :, 0] + img[:, :, 1]) + img[:, :, 2]) / f32(3.0)
        avg = f32(0.0)
        for v in d.ravel():                                    # Iterator::sum::<f32>() (light.rs:942)
            avg = f32(avg + v)
        avg = f32(avg / f32(d.size))
        comp = np.maximum(d - avg, f32(0.0)).astype(np.float32)
        if not comp.any():
            comp[:] = 1.0
        self.env_maps.append(dict(image=img, rfl=rfl, dist=d.astype(np.float32), comp=comp))
        self.extra_lights.append(dict(kind=ffi.SG_LIGHT_IMAGE_INFINITE, spectrum=self.spectrum(("dense", spectrum_dense(ill))),
                                      scale=float(sc), pos=np.zeros(3, np.float32), tri=len(self.env_maps) - 1))

    # -- flatten ---------------------------------------------------------------------------------
    def build(self):
        out = SceneDesc()
        A = out.arrays
        host = ffi.load_host_library()
        # spectra pool
        pool, spec_rows = [], np.zeros(len(self.spectra), dtype=np.dtype(ffi.SgSpectrum))
        off = 0
        recs = []
        for sp in self.spectra:
            r = ffi.SgSpectrum()
            if sp[0] == "const":
                r.kind, r.c = ffi.SG_SPECTRUM_CONSTANT, float(sp[1])
            elif sp[0] == "dense":
                v = np.asarray(sp[1], dtype=np.float32)
                r.kind, r.n, r.lambda_min, r.off_a = ffi.SG_SPECTRUM_DENSE, len(v), LAMBDA_MIN, off
                pool.append(v); off += len(v)
            elif sp[0] == "pl":
                lam, v = np.asarray(sp[1], dtype=np.float32), np.asarray(sp[2], dtype=np.float32)
                r.kind, r.n, r.off_a, r.off_b = ffi.SG_SPECTRUM_PIECEWISE_LINEAR, len(lam), off, off + len(lam)
                pool.extend([lam, v]); off += 2 * len(lam)
            else:
                raise ValueError(sp[0])
            recs.append(r)
        # film sensor spectra appended last
        film_ids = []
        for cname in "XYZ":
            v = cie(cname)
            r = ffi.SgSpectrum(); r.kind, r.n, r.lambda_min, r.off_a = ffi.SG_SPECTRUM_DENSE, len(v), LAMBDA_MIN, off
            pool.append(v); off += len(v); recs.append(r); film_ids.append(len(recs) - 1)
        # lights: non-area first, then one per emissive triangle in shape order (scene.rs:532-632)
        lights = []
        for el in self.extra_lights:
            L = ffi.SgLight(); L.kind = el["kind"]; L.spectrum = el["spectrum"]; L.scale = el["scale"]
            L.pos[:] = [float(x) for x in el["pos"]]
            L.tri = el.get("tri", 0)
            lights.append(L)
        mesh_light_base = {}
        for mi, m in enumerate(self.meshes):
            al = m["area_light"]
            if al is None:
                continue
            sc = f32(al.get("scale", 1.0)) / spectrum_to_photometric(al["L"])
            dense = spectrum_dense(al["L"])
            r = ffi.SgSpectrum(); r.kind, r.n, r.lambda_min, r.off_a = ffi.SG_SPECTRUM_DENSE, len(dense), LAMBDA_MIN, off
            pool.append(dense); off += len(dense); recs.append(r); sid = len(recs) - 1
            mesh_light_base[mi] = len(lights)
            P, I = m["p"], m["idx"]
            e1 = P[I[:, 1]] - P[I[:, 0]]; e2 = P[I[:, 2]] - P[I[:, 0]]
            area = 0.5 * np.linalg.norm(np.cross(e1.astype(np.float64), e2.astype(np.float64)), axis=1)
            for t in range(len(I)):
                L = ffi.SgLight(); L.kind = ffi.SG_LIGHT_DIFFUSE_AREA; L.spectrum = sid; L.scale = float(sc)
                L.two_sided = 1 if al.get("two_sided", False) else 0
                L.mesh, L.tri, L.area = mi, t, float(area[t])
                lights.append(L)
        patch_light_base = {}
        for pi_, m in enumerate(self.patch_meshes):                   # bilinear meshes follow the triangle meshes in shape order
            al = m.get("area_light")
            if al is None:
                continue
            sc = f32(al.get("scale", 1.0)) / spectrum_to_photometric(al["L"])
            dense = spectrum_dense(al["L"])
            r = ffi.SgSpectrum(); r.kind, r.n, r.lambda_min, r.off_a = ffi.SG_SPECTRUM_DENSE, len(dense), LAMBDA_MIN, off
            pool.append(dense); off += len(dense); recs.append(r); sid = len(recs) - 1
            patch_light_base[pi_] = len(lights)
            for t in range(len(m["idx"])):
                L = ffi.SgLight(); L.kind = ffi.SG_LIGHT_DIFFUSE_AREA_PATCH; L.spectrum = sid; L.scale = float(sc)
                L.two_sided = 1 if al.get("two_sided", False) else 0
                L.mesh, L.tri, L.area = len(self.meshes) + pi_, t, float(bilinear_patch_area(m["p"][m["idx"][t]]))
                lights.append(L)
        sphere_light = {}
        for si_, sp in enumerate(self.spheres):                       # spheres follow the meshes in shape order
            al = sp.get("area_light")
            if al is None:
                continue
            sc = f32(al.get("scale", 1.0)) / spectrum_to_photometric(al["L"])
            dense = spectrum_dense(al["L"])
            r = ffi.SgSpectrum(); r.kind, r.n, r.lambda_min, r.off_a = ffi.SG_SPECTRUM_DENSE, len(dense), LAMBDA_MIN, off
            pool.append(dense); off += len(dense); recs.append(r)
            L = ffi.SgLight(); L.kind = ffi.SG_LIGHT_DIFFUSE_AREA_SPHERE; L.spectrum = len(recs) - 1; L.scale = float(sc)
            L.two_sided = 1 if al.get("two_sided", False) else 0
            L.mesh, L.tri = 0, si_
            L.area = float(f32(sp["phi_max"]) * f32(sp["radius"]) * (f32(sp["z_max"]) - f32(sp["z_min"])))       # Sphere::area sphere.rs:295-297
            sphere_light[si_] = len(lights); lights.append(L)
        # geometry arrays
        all_meshes = self.meshes + self.patch_meshes
        any_n = any(m["n"] is not None for m in all_meshes)
        any_uv = any(m["uv"] is not None for m in all_meshes)
        nv = sum(len(m["p"]) for m in all_meshes); nt = sum(len(m["idx"]) for m in self.meshes)
        n_patches = sum(len(m["idx"]) for m in self.patch_meshes)
        A["p"] = np.empty((nv, 3), np.float32); A["idx"] = np.empty((nt, 3), np.uint32)
        A["n"] = np.zeros((nv, 3), np.float32) if any_n else None
        A["uv"] = np.zeros((nv, 2), np.float32) if any_uv else None
        mesh_rows = (ffi.SgMesh * len(all_meshes))()
        prim_in = np.empty((nt, 4), np.int64)       # mesh, tri, material, light (input order)
        gidx = np.empty((nt, 3), np.uint32)
        v0 = t0 = 0
        for mi, m in enumerate(self.meshes):
            k, t = len(m["p"]), len(m["idx"])
            A["p"][v0:v0 + k] = m["p"]; A["idx"][t0:t0 + t] = m["idx"]; gidx[t0:t0 + t] = m["idx"] + v0
            if m["n"] is not None: A["n"][v0:v0 + k] = m["n"]
            if m["uv"] is not None: A["uv"][v0:v0 + k] = m["uv"]
            mr = mesh_rows[mi]
            mr.first_index, mr.first_vertex, mr.n_triangles, mr.n_vertices, mr.flags = 3 * t0, v0, t, k, m["flags"]
            prim_in[t0:t0 + t, 0] = mi; prim_in[t0:t0 + t, 1] = np.arange(t); prim_in[t0:t0 + t, 2] = m["material"]
            prim_in[t0:t0 + t, 3] = (mesh_light_base[mi] + np.arange(t)) if mi in mesh_light_base else -1
            v0 += k; t0 += t
        # bilinear patch meshes: vertices after the triangle meshes', four indices per patch after the triangle indices
        A["patch_idx"] = np.empty((n_patches, 4), np.uint32)
        patch_prim_in = np.empty((n_patches, 4), np.int64); patch_bounds = np.empty((n_patches, 6), np.float32)
        q0 = 0
        for pi_, m in enumerate(self.patch_meshes):
            mi = len(self.meshes) + pi_
            k, t = len(m["p"]), len(m["idx"])
            A["p"][v0:v0 + k] = m["p"]; A["patch_idx"][q0:q0 + t] = m["idx"]
            if m["n"] is not None: A["n"][v0:v0 + k] = m["n"]
            if m["uv"] is not None: A["uv"][v0:v0 + k] = m["uv"]
            mr = mesh_rows[mi]
            mr.first_index, mr.first_vertex, mr.n_triangles, mr.n_vertices, mr.flags = 3 * nt + 4 * q0, v0, t, k, m["flags"]
            patch_prim_in[q0:q0 + t, 0] = mi; patch_prim_in[q0:q0 + t, 1] = np.arange(t); patch_prim_in[q0:q0 + t, 2] = m["material"]
            patch_prim_in[q0:q0 + t, 3] = (patch_light_base[pi_] + np.arange(t)) if pi_ in patch_light_base else -1
            P4 = m["p"][m["idx"]]                                           # BilinearPatch::bounds bilinear_patch.rs:430-433
            patch_bounds[q0:q0 + t, :3] = P4.min(axis=1); patch_bounds[q0:q0 + t, 3:] = P4.max(axis=1)
            v0 += k; q0 += t
        A["all_idx"] = np.concatenate([A["idx"].ravel(), A["patch_idx"].ravel()]).astype(np.uint32) if n_patches else A["idx"]
        # Triangle::bounds per triangle (input order)
        bounds = np.empty((nt, 6), np.float32)
        host.sh_triangle_bounds(nt, gidx.ctypes.data, A["p"].ctypes.data, bounds.ctypes.data)
        tri_obj = np.full(nt, -1, np.int64)
        t0 = 0
        for m in self.meshes:
            t = len(m["idx"])
            if m["object"] is not None: tri_obj[t0:t0 + t] = m["object"]
            t0 += t

        def build_bvh(b):            # BvhAggregate::new (aggregate.rs:207-290) -> (nodes, leaf order)
            k = len(b)
            nd = np.zeros(max(2 * k - 1, 1), dtype=np.dtype(ffi.SgBvhNode)); od = np.empty(k, np.uint32)
            bc = np.ascontiguousarray(b, np.float32)
            nn = host.sh_bvh_build(k, bc.ctypes.data, nd.ctypes.data, od.ctypes.data)
            if nn <= 0:
                raise ffi.ShimmerGpuError("sh_bvh_build failed")
            return nd[:nn].copy(), od
        # object definitions: one BvhAggregate each when they hold more than one primitive (scene.rs:818-830)
        obj_rows = (ffi.SgObject * max(self.n_objects, 1))()
        obj_nodes, obj_prims, obj_root_bounds = [], [], []
        patch_obj = np.full(n_patches, -1, np.int64)
        q0 = 0
        for m in self.patch_meshes:
            if m.get("object") is not None: patch_obj[q0:q0 + len(m["idx"])] = m["object"]
            q0 += len(m["idx"])
        sphere_obj = np.array([-1 if sp.get("object") is None else sp["object"] for sp in self.spheres], np.int64).reshape(-1)
        pending_objects = []                                 # filled in once sphere rows / bounds exist (below)
        for o in range(self.n_objects):
            pending_objects.append(o)
        # top level: shapes first, then one TransformedPrimitive per instance (scene.rs:806,849-866)
        top_sel = np.nonzero(tri_obj < 0)[0]
        inst_rows = (ffi.SgInstance * max(len(self.instances), 1))()
        for ii, (o, xf) in enumerate(self.instances):
            r = inst_rows[ii]
            r.render_from_primitive[:] = xf.m32().ravel().tolist(); r.primitive_from_render[:] = xf.m_inv.astype(np.float32).ravel().tolist()
            r.object = o
        # spheres: Sphere::bounds (sphere.rs:273-279) = Transform::apply(Bounds3f) of the object-space box (transform.rs:557-570)
        sph_rows = (ffi.SgSphere * max(len(self.spheres), 1))()
        sph_bounds = np.empty((len(self.spheres), 6), np.float32)
        for si_, sp in enumerate(self.spheres):
            r = sph_rows[si_]
            if not np.array_equal(sp["rfo"].m[3], [0.0, 0.0, 0.0, 1.0]):
                raise ValueError("sphere transforms must be affine")
            r.render_from_object[:] = sp["rfo"].m32().ravel().tolist(); r.object_from_render[:] = sp["rfo"].m_inv.astype(np.float32).ravel().tolist()
            r.radius, r.z_min, r.z_max = float(sp["radius"]), float(sp["z_min"]), float(sp["z_max"])
            r.theta_z_min, r.theta_z_max, r.phi_max, r.flags = float(sp["theta_z_min"]), float(sp["theta_z_max"]), float(sp["phi_max"]), sp["flags"]
            lo = np.array([-sp["radius"], -sp["radius"], sp["z_min"]], np.float32); hi = np.array([sp["radius"], sp["radius"], sp["z_max"]], np.float32)
            corners = np.array([[(hi if (c >> a) & 1 else lo)[a] for a in range(3)] for c in range(8)], np.float32)
            pc = sp["rfo"].apply_points_f32(corners)
            sph_bounds[si_, :3] = pc.min(axis=0); sph_bounds[si_, 3:] = pc.max(axis=0)
        A["spheres"] = sph_rows
        sp_in = np.zeros((len(self.spheres), 4), np.int64)
        if len(self.spheres):
            sp_in[:, 0] = ffi.SG_PRIM_SPHERE; sp_in[:, 1] = np.arange(len(self.spheres)); sp_in[:, 2] = [sp["material"] for sp in self.spheres]
            sp_in[:, 3] = [sphere_light.get(i, -1) for i in range(len(self.spheres))]
        # object definitions: one BvhAggregate each when they hold more than one primitive (scene.rs:818-830); shape order inside a
        # definition: triangle meshes, bilinear meshes, spheres (the builder's order, as at the top level)
        for o in pending_objects:
            sel_t, sel_p, sel_s = np.nonzero(tri_obj == o)[0], np.nonzero(patch_obj == o)[0], np.nonzero(sphere_obj == o)[0]
            o_prims = np.concatenate([prim_in[sel_t], patch_prim_in[sel_p], sp_in[sel_s]])
            o_bounds = np.concatenate([bounds[sel_t], patch_bounds[sel_p], sph_bounds[sel_s]])
            if len(o_prims) == 0:
                raise ValueError("empty object definition")
            if len(o_prims) > 1:
                nd, od = build_bvh(o_bounds)
                obj_nodes.append(nd); obj_prims.append(o_prims[od])
                obj_root_bounds.append(np.concatenate([nd[0]["bmin"], nd[0]["bmax"]]))
            else:
                obj_nodes.append(np.zeros(0, dtype=np.dtype(ffi.SgBvhNode))); obj_prims.append(o_prims)
                obj_root_bounds.append(o_bounds[0])
        inst_bounds = np.empty((len(self.instances), 6), np.float32)
        for ii, (o, xf) in enumerate(self.instances):
            lo, hi = obj_root_bounds[o][:3], obj_root_bounds[o][3:]
            corners = np.array([[(hi if (c >> a) & 1 else lo)[a] for a in range(3)] for c in range(8)], np.float32)   # Transform::apply(Bounds3f) transform.rs:557-570
            pc = xf.apply_points_f32(corners)
            inst_bounds[ii, :3] = pc.min(axis=0); inst_bounds[ii, 3:] = pc.max(axis=0)
        top_bounds = bounds[top_sel]
        top_prim_in = prim_in[top_sel]
        if n_patches:
            top_prim_in = np.concatenate([top_prim_in, patch_prim_in[patch_obj < 0]]); top_bounds = np.concatenate([top_bounds, patch_bounds[patch_obj < 0]])
        if len(self.spheres):
            top_prim_in = np.concatenate([top_prim_in, sp_in[sphere_obj < 0]]); top_bounds = np.concatenate([top_bounds, sph_bounds[sphere_obj < 0]])
        if len(self.instances):
            top_bounds = np.concatenate([top_bounds, inst_bounds])
        if len(self.instances):
            ip = np.zeros((len(self.instances), 4), np.int64)
            ip[:, 0] = ffi.SG_PRIM_INSTANCE; ip[:, 1] = np.arange(len(self.instances)); ip[:, 3] = -1
            top_prim_in = np.concatenate([top_prim_in, ip])
        nodes, order = build_bvh(top_bounds)
        n_top_nodes, n_top_prims = len(nodes), len(top_prim_in)
        all_nodes, all_prims = [nodes], [top_prim_in[order]]
        node_off, prim_off = n_top_nodes, n_top_prims
        for o in range(self.n_objects):
            obj_rows[o].first_node, obj_rows[o].n_nodes = node_off, len(obj_nodes[o])
            obj_rows[o].first_prim, obj_rows[o].n_prims = prim_off, len(obj_prims[o])
            all_nodes.append(obj_nodes[o]); all_prims.append(obj_prims[o])
            node_off += len(obj_nodes[o]); prim_off += len(obj_prims[o])
        A["nodes"] = np.concatenate(all_nodes); A["order"] = order; A["prim_bounds"] = top_bounds
        n_nodes = len(A["nodes"])
        po = np.concatenate(all_prims)
        nprim = len(po)
        prims = np.zeros(nprim, dtype=np.dtype(ffi.SgPrimitive))
        prims["mesh"], prims["tri"], prims["material"], prims["light"] = po[:, 0], po[:, 1], po[:, 2], po[:, 3]
        A["prims"] = prims
        A["objects"] = obj_rows; A["instances"] = inst_rows
        # scene bounds -> infinite-light preprocess (light.rs:797-802, bounding_box.rs:460-468)
        root = A["nodes"][0]
        bmin, bmax = np.array(root["bmin"], np.float32), np.array(root["bmax"], np.float32)
        center = (bmin + bmax) / f32(2.0)
        inside = bool(np.all(center >= bmin) and np.all(center <= bmax))
        radius = float(np.sqrt(np.sum((center - bmax).astype(np.float32) ** 2, dtype=np.float32))) if inside else 0.0
        for L in lights:
            if L.kind in (ffi.SG_LIGHT_UNIFORM_INFINITE, ffi.SG_LIGHT_IMAGE_INFINITE):
                L.scene_center[:] = center.tolist(); L.scene_radius = radius
        env_rows = (ffi.SgEnvMap * max(len(self.env_maps), 1))()
        for ei, em in enumerate(self.env_maps):
            E = env_rows[ei]
            E.render_from_light[:] = em["rfl"].m32().ravel().tolist(); E.light_from_render[:] = em["rfl"].m_inv.astype(np.float32).ravel().tolist()
            E.res = em["image"].shape[0]
            for name, func2d in (("distribution", em["dist"]), ("compensated", em["comp"])):
                fn, cdf, mf, mcdf, mint = piecewise_constant_2d(func2d)
                D2 = getattr(E, name)
                D2.nu, D2.nv = func2d.shape[1], func2d.shape[0]
                D2.func_off = off; pool.append(fn.ravel()); off += fn.size
                D2.cdf_off = off; pool.append(cdf.ravel()); off += cdf.size
                D2.marg_func_off = off; pool.append(mf); off += mf.size
                D2.marg_cdf_off = off; pool.append(mcdf); off += mcdf.size
                D2.marg_integral = float(mint)
        A["env_maps"] = env_rows
        A["pool"] = np.concatenate(pool).astype(np.float32) if pool else np.zeros(1, np.float32)
        A["spectra"] = (ffi.SgSpectrum * len(recs))(*recs)
        mats = (ffi.SgMaterial * max(len(self.materials), 1))()
        for i, m in enumerate(self.materials):
            mats[i].kind, mats[i].spec_a, mats[i].spec_b, mats[i].flags = m["kind"], m["spec_a"], m["spec_b"], m["flags"]
            mats[i].u_roughness, mats[i].v_roughness, mats[i].displacement = m["ur"], m["vr"], 0.0
            mats[i].spec_c = m.get("spec_c", -1); mats[i].thickness = m.get("thickness", 0.0); mats[i].g = m.get("g", 0.0)
            mats[i].max_depth = m.get("max_depth", 0); mats[i].n_samples = m.get("n_samples", 0)
            mats[i].tex_reflectance = m.get("tex_reflectance", -1); mats[i].tex_displacement = m.get("tex_displacement", -1)
            mats[i].spec_d = m.get("spec_d", -1); mats[i].u_roughness2 = m.get("ur2", 0.0); mats[i].v_roughness2 = m.get("vr2", 0.0)
            mats[i].normal_map = m.get("normal_map", -1)
            mm = m.get("mix_materials", (-1, -1)); mats[i].mix_materials[0], mats[i].mix_materials[1] = mm
            mats[i].mix_amount = m.get("mix_amount", 0.0); mats[i].tex_mix_amount = m.get("tex_mix_amount", -1)
        A["materials"] = mats
        A["material_textures"] = None
        if any(m.get("param_textures") for m in self.materials):
            mt_rows = (ffi.SgMaterialTextures * len(self.materials))()
            for i, m in enumerate(self.materials):
                for nm in ffi.SgMaterialTextures.NAMES:
                    setattr(mt_rows[i], nm, m.get("param_textures", {}).get(nm, -1))
            A["material_textures"] = mt_rows
        A["lights"] = (ffi.SgLight * max(len(lights), 1))(*lights)
        A["meshes"] = mesh_rows
        # image textures: every MIP level, linear f32 texels, channels interleaved
        tex_rows = (ffi.SgTexture * max(len(self.textures), 1))()
        level_rows, texel_chunks, toff = [], [], 0
        node_rows = []
        for ti, t in enumerate(self.textures):
            r = tex_rows[ti]
            if t.get("kind", 0) != ffi.SG_TEXTURE_IMAGE:
                nd = ffi.SgTextureNode(); nn = t["node"]
                nd.tex1, nd.tex2, nd.amount, nd.spectrum = nn.get("tex1", -1), nn.get("tex2", -1), nn.get("amount", -1), nn.get("spectrum", -1)
                nd.value = nn.get("value", 0.0); nd.dir[:] = list(nn.get("dir", (0.0, 1.0, 0.0)))
                r.n_channels, r.kind, r.node, r.mapping = t["n_channels"], t["kind"], len(node_rows), -1
                node_rows.append(nd)
                continue
            r.n_channels, r.n_levels, r.first_level = t["n_channels"], len(t["levels"]), len(level_rows)
            r.wrap, r.filter, r.max_anisotropy, r.scale, r.invert = t["wrap"], t["filter"], t["max_anisotropy"], t["scale"], t["invert"]
            r.su, r.sv, r.du, r.dv, r.spectrum_type = t["su"], t["sv"], t["du"], t["dv"], t["spectrum_type"]
            r.mapping = t.get("mapping", -1)
            for lv in t["levels"]:
                L = ffi.SgImageLevel(); L.offset = toff; L.res[:] = [lv.shape[1], lv.shape[0]]
                level_rows.append(L); texel_chunks.append(np.ascontiguousarray(lv, np.float32).ravel()); toff += lv.size
        A["textures"] = tex_rows
        A["texture_nodes"] = (ffi.SgTextureNode * max(len(node_rows), 1))(*node_rows)
        map_rows = (ffi.SgTextureMapping * max(len(self.mappings), 1))()
        for mi_, mp in enumerate(self.mappings):
            r = map_rows[mi_]
            r.kind = mp["kind"]; r.texture_from_render[:] = mp["tfr"].m32().ravel().tolist()
            r.vs[:] = [float(x) for x in mp["vs"]]; r.vt[:] = [float(x) for x in mp["vt"]]; r.ds, r.dt = mp["ds"], mp["dt"]
        A["texture_mappings"] = map_rows
        for em in self.env_maps:                               # environment maps live in the texel pool after the MIP levels
            em["texel_offset"] = toff
            texel_chunks.append(em["image"].ravel()); toff += em["image"].size
        for ei, em in enumerate(self.env_maps):
            A["env_maps"][ei].texel_offset = em["texel_offset"]
        A["image_levels"] = (ffi.SgImageLevel * max(len(level_rows), 1))(*level_rows)
        A["texels"] = np.concatenate(texel_chunks) if texel_chunks else np.zeros(1, np.float32)
        A["mip_lut"] = np.ascontiguousarray(tables()["MIP_FILTER_LUT"], np.float32)
        need_rgb = any(t["n_channels"] == 3 and t.get("kind", 0) == ffi.SG_TEXTURE_IMAGE for t in self.textures) or len(self.env_maps) > 0
        if need_rgb:
            from . import rgb2spec
            sc_, dt_ = rgb2spec.build_table(16)
            A["rgb2spec_scale"] = np.ascontiguousarray(sc_, np.float32); A["rgb2spec_data"] = np.ascontiguousarray(dt_, np.float32).ravel()
        d = out.desc
        d.abi_version = ffi.SG_ABI_VERSION
        d.n_nodes = n_nodes; d.nodes = _as_ptr(A["nodes"], ffi.SgBvhNode)
        d.n_primitives = nprim; d.primitives = _as_ptr(A["prims"], ffi.SgPrimitive)
        d.n_top_nodes, d.n_top_primitives = n_top_nodes, n_top_prims
        d.n_objects = self.n_objects; d.objects = A["objects"]
        d.n_instances = len(self.instances); d.instances = A["instances"]
        d.n_spheres = len(self.spheres); d.spheres = A["spheres"]
        d.scene_flags = ffi.SG_SCENE_FIX_INSTANCING if self.fix_instancing else 0
        d.n_meshes = len(all_meshes); d.meshes = A["meshes"]
        d.n_indices = 3 * nt + 4 * n_patches; d.indices = _as_ptr(A["all_idx"], C.c_uint32)
        d.n_vertices = nv; d.p = _as_ptr(A["p"], C.c_float)
        d.n = _as_ptr(A["n"], C.c_float) if any_n else None
        d.uv = _as_ptr(A["uv"], C.c_float) if any_uv else None
        d.s = None
        d.n_spectra = len(recs); d.spectra = A["spectra"]
        d.n_pool = len(A["pool"]); d.spectrum_pool = _as_ptr(A["pool"], C.c_float)
        d.n_materials = len(self.materials); d.materials = A["materials"]
        d.n_lights = len(lights); d.lights = A["lights"]
        d.n_textures = len(self.textures); d.textures = A["textures"]
        d.n_image_levels = len(level_rows); d.image_levels = A["image_levels"]
        d.n_texels = len(A["texels"]) if texel_chunks else 0; d.texels = _as_ptr(A["texels"], C.c_float)
        d.mip_filter_lut = _as_ptr(A["mip_lut"], C.c_float)
        if need_rgb:
            d.rgb2spec_res = len(A["rgb2spec_scale"]); d.rgb2spec_scale = _as_ptr(A["rgb2spec_scale"], C.c_float)
            d.rgb2spec_data = _as_ptr(A["rgb2spec_data"], C.c_float)
        d.n_texture_mappings = len(self.mappings); d.texture_mappings = A["texture_mappings"]
        d.n_texture_nodes = len(node_rows); d.texture_nodes = A["texture_nodes"] if node_rows else None
        d.material_textures = A["material_textures"]
        d.n_env_maps = len(self.env_maps); d.env_maps = A["env_maps"]
        d.camera = self.camera
        self.film.r_bar, self.film.g_bar, self.film.b_bar = film_ids
        d.film = self.film
        inst_tris = sum(int(obj_rows[o].n_prims) for o, _ in self.instances)
        upload_bytes = sum(int(getattr(v, "nbytes", 0) or (C.sizeof(v) if isinstance(v, C.Array) else 0)) for v in A.values() if v is not None)
        out.meta = dict(upload_bytes=upload_bytes, n_triangles=nt, n_instanced_triangles=int(len(top_sel) + inst_tris), n_nodes=int(n_nodes), n_lights=len(lights), n_spheres=len(self.spheres), n_patches=int(n_patches),
                        resolution=tuple(self.film.full_resolution), window=tuple(self.film.pixel_bounds))
        return out
