"""ctypes loader for the CPU oracle (oracle/_build/liborc.so).  TEST INFRASTRUCTURE: importable only
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORC_DIR = os.path.join(ROOT, "oracle")
ORC_LIB = os.path.join(ORC_DIR, "_build", "liborc.so")

import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from shimmer_b200 import ffi  # noqa: E402  (struct definitions only)

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORC_DIR], check=True)
    return ORC_LIB


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()      # `make` is a no-op when liborc.so is newer than its sources and the ABI header
    L = C.CDLL(ORC_LIB)
    vp = C.c_void_p
    L.orc_header.restype = C.c_char_p
    L.orc_sampler_fill.argtypes = [C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_int64, vp]
    L.orc_rng_u64.argtypes = [C.c_uint64, C.c_int, vp, C.c_int64, vp, vp]
    L.orc_bvh_build.argtypes = [C.c_int64, vp, vp, vp]; L.orc_bvh_build.restype = C.c_int64
    L.orc_trace.argtypes = [vp, C.c_int64, vp, vp, vp, C.c_int, vp, vp, C.c_int]
    L.orc_camera_rays.argtypes = [vp, vp, C.c_int64, vp, vp, vp, vp]
    L.orc_render.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int]; L.orc_render.restype = C.c_double
    L.orc_film_develop.argtypes = [vp, vp, C.c_int64, vp]
    L.orc_film_get_image.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_uint32, vp]
    L.orc_f16_round.argtypes = [C.c_int64, vp, vp, vp]
    for name, nargs in (("orc_difference_of_products", 4), ("orc_lerp", 3), ("orc_next_float_up", 1), ("orc_next_float_down", 1),
                        ("orc_visible_wavelengths_pdf", 1), ("orc_sample_visible_wavelengths", 1), ("orc_fresnel_dielectric", 2),
                        ("orc_fresnel_complex", 3), ("orc_blackbody", 2)):
        f = getattr(L, name); f.argtypes = [C.c_float] * nargs; f.restype = C.c_float
    L.orc_gamma.argtypes = [C.c_int]; L.orc_gamma.restype = C.c_float
    L.orc_tr_d.argtypes = [C.c_float, C.c_float, vp]; L.orc_tr_d.restype = C.c_float
    L.orc_tr_g.argtypes = [C.c_float, C.c_float, vp, vp]; L.orc_tr_g.restype = C.c_float
    L.orc_spectrum_get.argtypes = [vp, C.c_int, C.c_float]; L.orc_spectrum_get.restype = C.c_float
    L.orc_spectrum_sample.argtypes = [vp, C.c_int, vp, vp]
    L.orc_dielectric_sample_f.argtypes = [C.c_float, C.c_float, C.c_float, vp, C.c_float, vp, vp]; L.orc_dielectric_sample_f.restype = C.c_int
    L.orc_bxdf_eval.argtypes = [C.c_int, vp, vp, vp, vp]
    L.orc_bxdf_sample.argtypes = [C.c_int, vp, vp, C.c_float, vp, vp]; L.orc_bxdf_sample.restype = C.c_int
    L.orc_tri_intersect.argtypes = [vp, vp, C.c_float, vp, vp]; L.orc_tri_intersect.restype = C.c_int
    L.orc_bounds_intersect.argtypes = [vp, vp, vp, vp, C.c_float]; L.orc_bounds_intersect.restype = C.c_int
    L.orc_light_sample.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]; L.orc_light_sample.restype = C.c_int
    L.orc_light_pdf.argtypes = [vp, C.c_int, vp, vp, vp, vp]; L.orc_light_pdf.restype = C.c_float
    L.orc_light_sample_complete.argtypes = [vp, C.c_int, vp, vp, vp, vp]; L.orc_light_sample_complete.restype = C.c_int
    L.orc_light_pdf_complete.argtypes = [vp, C.c_int, vp]; L.orc_light_pdf_complete.restype = C.c_float
    L.orc_light_le.argtypes = [vp, C.c_int, vp, vp, vp]
    L.orc_equal_area_square_to_sphere.argtypes = [vp, vp]; L.orc_equal_area_sphere_to_square.argtypes = [vp, vp]
    L.orc_rotate_from_to.argtypes = [vp, vp, vp]
    L.orc_sigmoid_poly_get.argtypes = [vp, C.c_float]; L.orc_sigmoid_poly_get.restype = C.c_float
    L.orc_rgb2spec_fetch.argtypes = [vp, vp, vp]
    L.orc_texture_eval.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp]
    L.orc_texture_eval_p.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp]
    L.orc_texture_eval_ctx.argtypes = [vp, C.c_int, C.c_int, C.c_int64, vp, vp, vp, vp, vp]
    L.orc_image_pyramid_levels.argtypes = [C.c_int32, C.c_int32, vp]; L.orc_image_pyramid_levels.restype = C.c_int32
    L.orc_image_generate_pyramid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]; L.orc_image_generate_pyramid.restype = C.c_int32
    L.orc_path_rays.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int64, vp]; L.orc_path_rays.restype = C.c_int64
    L.orc_set_debug_perturb.argtypes = [C.c_float]
    L.orc_approximate_dp_dxy.argtypes = [vp, vp, vp, C.c_int, C.c_uint32, vp]
    _lib = L
    return L


def fa(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def make_params(seed=0, spp=4, sample_range=None, max_depth=5, regularize=False, flags=0, integrator="path", sample_lights=True,
                sample_bsdf=True):
    p = ffi.SgRenderParams()
    p.integrator = {"path": ffi.SG_INTEGRATOR_PATH, "simplepath": ffi.SG_INTEGRATOR_SIMPLE_PATH, "randomwalk": ffi.SG_INTEGRATOR_RANDOM_WALK}[integrator]
    p.integrator_flags = (ffi.SG_SIMPLEPATH_SAMPLE_LIGHTS if sample_lights else 0) | (ffi.SG_SIMPLEPATH_SAMPLE_BSDF if sample_bsdf else 0)
    p.seed = seed; p.samples_per_pixel = spp
    p.sample_begin, p.sample_end = sample_range if sample_range else (0, spp)
    p.max_depth = max_depth; p.regularize = int(regularize); p.option_flags = flags
    return p


def render(scene, params, n_threads=None, stream_mode=0):
    """Returns (film (H*W,4) f64, stats, seconds)."""
    x0, y0, x1, y1 = scene.desc.film.pixel_bounds
    film = np.zeros(((y1 - y0) * (x1 - x0), 4), np.float64)
    st = ffi.SgStats()
    nt = n_threads or os.cpu_count() or 1
    secs = lib().orc_render(scene.ptr(), C.byref(params), film.ctypes.data, C.byref(st), nt, stream_mode)
    return film, st, secs


def trace(scene, o, d, t_max, any_hit=False, n_threads=None):
    o, d, t_max = fa(o), fa(d), fa(t_max)
    n = len(t_max)
    out = np.zeros(n, dtype=np.dtype(ffi.SgHit))
    st = ffi.SgStats()
    lib().orc_trace(scene.ptr(), n, o.ctypes.data, d.ctypes.data, t_max.ctypes.data, int(any_hit), out.ctypes.data,
                    C.byref(st), n_threads or os.cpu_count() or 1)
    return out, st


def camera_rays(scene, params, pixel_xy, sample_index):
    pixel_xy = np.ascontiguousarray(pixel_xy, np.int32); sample_index = np.ascontiguousarray(sample_index, np.int32)
    n = len(sample_index)
    rays = np.zeros((n, 6), np.float32); lam = np.zeros((n, 8), np.float32)
    lib().orc_camera_rays(scene.ptr(), C.byref(params), n, pixel_xy.ctypes.data, sample_index.ctypes.data,
                          rays.ctypes.data, lam.ctypes.data)
    return rays, lam


def develop(scene, film):
    out = np.zeros((len(film), 3), np.float32)
    lib().orc_film_develop(scene.ptr(), np.ascontiguousarray(film).ctypes.data, len(film), out.ctypes.data)
    return out


def film_get_image(scene, film, width, height, fp16=True, bottom_up=False):
    out = np.zeros((height, width, 3), np.float32)
    lib().orc_film_get_image(scene.ptr(), np.ascontiguousarray(film).ctypes.data, width, height, (1 if fp16 else 0) | (2 if bottom_up else 0),
                             out.ctypes.data)
    return out


def f16_round(x):
    x = fa(x).ravel(); out = np.zeros_like(x); bits = np.zeros(len(x), np.uint16)
    lib().orc_f16_round(len(x), x.ctypes.data, out.ctypes.data, bits.ctypes.data)
    return out, bits


def texture_eval(scene, tex, q, lambda4=None, as_float=False):
    """q: n x 6 (u v dudx dudy dvdx dvdy); returns n x 4."""
    q = fa(q).reshape(-1, 6); n = len(q)
    lam = fa(np.tile([450.0, 520.0, 600.0, 680.0], (n, 1)) if lambda4 is None else lambda4).reshape(-1, 4)
    out = np.zeros((n, 4), np.float32)
    lib().orc_texture_eval(scene.ptr(), tex, 1 if as_float else 0, n, q.ctypes.data, lam.ctypes.data, out.ctypes.data)
    return out


def texture_eval_p(scene, tex, p, q=None, dpdx=None, dpdy=None, lambda4=None, as_float=False):
    """Texture lookups with a full TextureEvalContext (p, dpdx, dpdy in render space) -- the non-UV mappings."""
    p = fa(p).reshape(-1, 3); n = len(p)
    q = fa(np.zeros((n, 6)) if q is None else q).reshape(-1, 6)
    pdp = fa(np.concatenate([p, np.zeros((n, 3)) if dpdx is None else fa(dpdx).reshape(-1, 3),
                             np.zeros((n, 3)) if dpdy is None else fa(dpdy).reshape(-1, 3)], axis=1))
    lam = fa(np.tile([450.0, 520.0, 600.0, 680.0], (n, 1)) if lambda4 is None else lambda4).reshape(-1, 4)
    out = np.zeros((n, 4), np.float32)
    lib().orc_texture_eval_p(scene.ptr(), int(tex), 1 if as_float else 0, n, q.ctypes.data, pdp.ctypes.data, lam.ctypes.data, out.ctypes.data)
    return out


def texture_eval_ctx(scene, tex, q, n, p=None, lambda4=None, as_float=False):
    """Texture lookups with TextureEvalContext::n (and optionally p) -- the direction-mix textures."""
    q = fa(q).reshape(-1, 6); cnt = len(q)
    nrm = fa(n).reshape(-1, 3)
    pdp = None if p is None else fa(np.concatenate([fa(p).reshape(-1, 3), np.zeros((cnt, 6), np.float32)], axis=1))
    lam = fa(np.tile([450.0, 520.0, 600.0, 680.0], (cnt, 1)) if lambda4 is None else lambda4).reshape(-1, 4)
    out = np.zeros((cnt, 4), np.float32)
    lib().orc_texture_eval_ctx(scene.ptr(), int(tex), 1 if as_float else 0, cnt, q.ctypes.data, None if pdp is None else pdp.ctypes.data,
                               nrm.ctypes.data, lam.ctypes.data, out.ctypes.data)
    return out


def generate_pyramid(image, wrap="repeat"):
    """Image::generate_pyramid (image.rs:699-787) incl. the resize of non-power-of-two images -> list of (H, W, C) f32 levels."""
    img = fa(image)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    res = np.zeros(64, np.int32)
    n = lib().orc_image_pyramid_levels(w, h, res.ctypes.data)
    sizes = [(int(res[2 * l + 1]), int(res[2 * l])) for l in range(n)]
    out = np.zeros(sum(a * b for a, b in sizes) * c, np.float32)
    lib().orc_image_generate_pyramid(img.ctypes.data, w, h, c, {"repeat": 0, "black": 1, "clamp": 2}[wrap], out.ctypes.data)
    levels, off = [], 0
    for a, b in sizes:
        levels.append(out[off:off + a * b * c].reshape(a, b, c).copy()); off += a * b * c
    return levels


def path_rays(scene, params, px, py, sample, max_rays=64):
    """Debug: the rays one (pixel, sample) path traces in the oracle -> (n, 10) array: o, d, t_max, any_hit, hit prim, hit t."""
    out = np.zeros((max_rays, 10), np.float32)
    n = lib().orc_path_rays(scene.ptr(), C.byref(params), px, py, sample, max_rays, out.ctypes.data)
    return out[:n]
