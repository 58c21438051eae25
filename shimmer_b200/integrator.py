"""Host-side mirror of the reference's integrator interface for the GPU path.

Reference interface (paths relative to /root/reference/src):
  * `pub trait Integrator { fn render(&mut self, options: &Options); }`        integrator.rs:52-54
  * `create_integrator(name, parameters, camera, sampler, aggregate, lights, color_space)`
    with the string registry "simplepath" | "randomwalk" | "path"                integrator.rs:16-42
  * `Options` (seed, pixel_samples, disable_pixel_jitter, disable_wavelength_jitter,
    disable_texture_filtering, force_diffuse, wavefront, ...)                     options.rs:15-36
  * `render_cpu(scene, options)`                                                  render.rs:8-55

Here the registry gains "wavefront" (selected by the already-parsed `Options.wavefront`,
main.rs:89-91,152-155) and `render_gpu` is the sibling of `render_cpu`.  Unknown names raise
like the reference panics.  There is no CPU fallback: if libshimmer_gpu.so or a GPU is
missing the constructor raises ShimmerGpuError.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import ffi


@dataclass
class Options:
    """options.rs:15-36 (only the fields the hot path reads)."""
    seed: int = 0
    pixel_samples: Optional[int] = None
    disable_pixel_jitter: bool = False
    disable_wavelength_jitter: bool = False
    disable_texture_filtering: bool = False
    force_diffuse: bool = False
    wavefront: bool = True
    gpu_device: int = 0

    def flags(self):
        f = 0
        if self.disable_pixel_jitter: f |= ffi.SG_OPT_DISABLE_PIXEL_JITTER
        if self.disable_wavelength_jitter: f |= ffi.SG_OPT_DISABLE_WAVELENGTH_JITTER
        if self.disable_texture_filtering: f |= ffi.SG_OPT_DISABLE_TEXTURE_FILTERING
        if self.force_diffuse: f |= ffi.SG_OPT_FORCE_DIFFUSE
        return f


class Integrator:
    def render(self, options: Options):
        raise NotImplementedError


_initialised_device = None


def _ensure_init(device):
    """`device`: one CUDA device index (sg_init) or a list of them (sg_init_multi: this process drives all of them, scenes are
    replicated and `render` splits the sample range across the devices with one in-library NCCL reduce)."""
    global _initialised_device
    lib = ffi.load_library()
    key = tuple(int(d) for d in device) if isinstance(device, (list, tuple)) else int(device)
    if _initialised_device != key:
        if isinstance(key, tuple):
            from .distributed import _preload_process_nccl
            _preload_process_nccl()
            arr = (C.c_int * len(key))(*key)
            ffi.check(lib.sg_init_multi(arr, len(key)), "sg_init_multi")
        else:
            ffi.check(lib.sg_init(key), "sg_init")
        _initialised_device = key
    return lib


class WavefrontPathIntegrator(Integrator):
    """GPU replacement of `ImageTileIntegrator` + `PathIntegrator` (integrator.rs:119-321,730-963).

    parameters: the integrator's ParameterDictionary -- `maxdepth` (5), `regularize` (false),
    `lightsampler` ("uniform"; anything else raises, light_sampler.rs:30-33); `integrator` picks the
    RayPathLiEvaluator (integrator.rs:398-403): "path" (default), "simplepath" (+ `samplelights`, `samplebsdf`, both
    true by default, integrator.rs:131-141) or "randomwalk".
    sampler: dict with `pixelsamples` (4) and `seed` (sampler.rs:95-99)."""

    def __init__(self, scene, parameters=None, sampler=None, device=0, max_paths_in_flight=0):
        parameters = dict(parameters or {})
        sampler = dict(sampler or {})
        self.max_depth = int(parameters.get("maxdepth", 5))
        self.regularize = bool(parameters.get("regularize", False))
        ls = parameters.get("lightsampler", "uniform")
        if ls != "uniform":
            raise ffi.ShimmerGpuError(f"Unknown light sampler: {ls}")
        li = parameters.get("integrator", "path")
        kinds = {"path": ffi.SG_INTEGRATOR_PATH, "simplepath": ffi.SG_INTEGRATOR_SIMPLE_PATH, "randomwalk": ffi.SG_INTEGRATOR_RANDOM_WALK}
        if li not in kinds:
            raise ffi.ShimmerGpuError(f"Unknown integrator {li}")
        self.integrator = kinds[li]
        self.integrator_flags = (ffi.SG_SIMPLEPATH_SAMPLE_LIGHTS if parameters.get("samplelights", True) else 0) | \
                                (ffi.SG_SIMPLEPATH_SAMPLE_BSDF if parameters.get("samplebsdf", True) else 0)
        self.samples_per_pixel = int(sampler.get("pixelsamples", 4))
        self.sampler_seed = sampler.get("seed", None)
        self.scene = scene
        self.max_paths_in_flight = int(max_paths_in_flight)
        self.device = device
        self._lib = _ensure_init(device)
        h = C.c_void_p()
        ffi.check(self._lib.sg_scene_create(scene.ptr(), C.byref(h)), "sg_scene_create")
        self._handle = h
        x0, y0, x1, y1 = scene.desc.film.pixel_bounds
        self.width, self.height = x1 - x0, y1 - y0
        self.film = np.zeros((self.height * self.width, 4), dtype=np.float64)
        self.stats = ffi.SgStats()

    # -- helpers -----------------------------------------------------------------------------
    def _params(self, options, sample_range=None, flags=0):
        spp = options.pixel_samples if options.pixel_samples is not None else self.samples_per_pixel
        seed = self.sampler_seed if self.sampler_seed is not None else options.seed
        p = ffi.SgRenderParams()
        p.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        p.samples_per_pixel = spp
        p.sample_begin, p.sample_end = sample_range if sample_range else (0, spp)
        p.max_depth = self.max_depth; p.regularize = int(self.regularize)
        p.option_flags = options.flags(); p.max_paths_in_flight = self.max_paths_in_flight; p.flags = flags
        p.integrator = self.integrator; p.integrator_flags = self.integrator_flags
        return p

    # -- Integrator::render ------------------------------------------------------------------
    def render(self, options: Options, sample_range=None, flags=0):
        """Renders into self.film (host, f64 rgb_sum[3] + weight_sum per pixel, row-major)."""
        p = self._params(options, sample_range, flags)
        ffi.check(self._lib.sg_render(self._handle, C.byref(p), self.film.ctypes.data, C.byref(self.stats)), "sg_render")
        return self.film

    def render_device(self, options: Options, d_film_ptr, sample_range=None, stream=None, flags=0):
        """Accumulates into a device film buffer (e.g. a torch.float64 CUDA tensor's data_ptr()).  `stream`: the cudaStream_t handle
        the caller's work on that buffer is ordered on; None / 0 = the legacy default stream (torch's default stream)."""
        p = self._params(options, sample_range, flags)
        ffi.check(self._lib.sg_render_device(self._handle, C.byref(p), C.c_void_p(d_film_ptr), C.byref(self.stats),
                                             C.c_void_p(stream) if stream else None), "sg_render_device")

    def trace(self, o, d, t_max, any_hit=False, want_stats=False):
        """PrimitiveI::intersect / intersect_predicate for a batch of rays (primitive.rs:15-28)."""
        o = np.ascontiguousarray(o, np.float32); d = np.ascontiguousarray(d, np.float32)
        t_max = np.ascontiguousarray(t_max, np.float32)
        n = len(t_max)
        out = np.zeros(n, dtype=np.dtype(ffi.SgHit))
        st = ffi.SgStats()
        ffi.check(self._lib.sg_trace(self._handle, n, o.ctypes.data, d.ctypes.data, t_max.ctypes.data, int(any_hit),
                                     out.ctypes.data, C.byref(st) if want_stats else None), "sg_trace")
        return (out, st) if want_stats else out

    def camera_rays(self, options, pixel_xy, sample_index):
        pixel_xy = np.ascontiguousarray(pixel_xy, np.int32); sample_index = np.ascontiguousarray(sample_index, np.int32)
        n = len(sample_index)
        rays = np.zeros((n, 6), np.float32); lam = np.zeros((n, 8), np.float32)
        p = self._params(options)
        ffi.check(self._lib.sg_camera_rays(self._handle, C.byref(p), n, pixel_xy.ctypes.data, sample_index.ctypes.data,
                                           rays.ctypes.data, lam.ctypes.data), "sg_camera_rays")
        return rays, lam

    def texture_eval(self, tex, q, lambda4=None, as_float=False):
        """SpectrumImageTexture / FloatImageTexture::evaluate (texture.rs:393-404,777-808) for n lookups; q = n x 6
        (u v dudx dudy dvdx dvdy) -> n x 4."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 6); n = len(q)
        lam = np.ascontiguousarray(np.tile([450.0, 520.0, 600.0, 680.0], (n, 1)) if lambda4 is None else lambda4, np.float32).reshape(-1, 4)
        out = np.zeros((n, 4), np.float32)
        ffi.check(self._lib.sg_texture_eval(self._handle, int(tex), 1 if as_float else 0, n, q.ctypes.data, lam.ctypes.data,
                                            out.ctypes.data), "sg_texture_eval")
        return out

    def texture_eval_ctx(self, tex, q, n, p=None, lambda4=None, as_float=False):
        """Texture lookups with TextureEvalContext::n (and optionally p): the direction-mix textures (texture.rs:295-310,:810-826)."""
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 6); cnt = len(q)
        nrm = np.ascontiguousarray(n, np.float32).reshape(-1, 3)
        pdp = None
        if p is not None:
            pdp = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 3), np.zeros((cnt, 6), np.float32)], axis=1), np.float32)
        lam = np.ascontiguousarray(np.tile([450.0, 520.0, 600.0, 680.0], (cnt, 1)) if lambda4 is None else lambda4, np.float32).reshape(-1, 4)
        out = np.zeros((cnt, 4), np.float32)
        ffi.check(self._lib.sg_texture_eval_ctx(self._handle, int(tex), 1 if as_float else 0, cnt, q.ctypes.data,
                                                None if pdp is None else pdp.ctypes.data, nrm.ctypes.data, lam.ctypes.data, out.ctypes.data),
                  "sg_texture_eval_ctx")
        return out

    def texture_eval_p(self, tex, p, q=None, dpdx=None, dpdy=None, lambda4=None, as_float=False):
        """Texture lookups with the full TextureEvalContext (p, dpdx, dpdy in render space): the non-UV mappings
        (texture.rs:938-1035)."""
        p = np.ascontiguousarray(p, np.float32).reshape(-1, 3); n = len(p)
        q = np.ascontiguousarray(np.zeros((n, 6)) if q is None else q, np.float32).reshape(-1, 6)
        z = np.zeros((n, 3), np.float32)
        pdp = np.ascontiguousarray(np.concatenate([p, z if dpdx is None else np.asarray(dpdx, np.float32).reshape(-1, 3),
                                                   z if dpdy is None else np.asarray(dpdy, np.float32).reshape(-1, 3)], axis=1), np.float32)
        lam = np.ascontiguousarray(np.tile([450.0, 520.0, 600.0, 680.0], (n, 1)) if lambda4 is None else lambda4, np.float32).reshape(-1, 4)
        out = np.zeros((n, 4), np.float32)
        ffi.check(self._lib.sg_texture_eval_p(self._handle, int(tex), 1 if as_float else 0, n, q.ctypes.data, pdp.ctypes.data,
                                              lam.ctypes.data, out.ctypes.data), "sg_texture_eval_p")
        return out

    def develop(self, film=None):
        """RgbFilm::get_pixel_rgb (film.rs:720-738) -> (H, W, 3) f32 output RGB."""
        film = self.film if film is None else film
        out = np.zeros((self.height * self.width, 3), np.float32)
        ffi.check(self._lib.sg_film_develop(self._handle, np.ascontiguousarray(film).ctypes.data, len(out), out.ctypes.data),
                  "sg_film_develop")
        return out.reshape(self.height, self.width, 3)

    def get_image(self, film=None, write_fp16=True, bottom_up=False):
        """RgbFilm::get_image (film.rs:647-707): developed RGB after the fp16 clamp + f16 quantisation of the default
        `savefp16` film (film.rs:491) -> (H, W, 3) f32; bottom_up=True gives PFM raster order (image.rs:1350)."""
        film = self.film if film is None else film
        out = np.zeros((self.height, self.width, 3), np.float32)
        ffi.check(self._lib.sg_film_get_image(self._handle, np.ascontiguousarray(film).ctypes.data, self.width, self.height,
                                              (1 if write_fp16 else 0) | (2 if bottom_up else 0), out.ctypes.data), "sg_film_get_image")
        return out

    def write_image(self, path, film=None, write_fp16=True):
        """RgbFilm::write_image (film.rs:709-713) for the one output format the reference can write (PFM, image.rs:1314-1327)."""
        if not str(path).lower().endswith(".pfm"):
            raise ffi.ShimmerGpuError("Invalid file extension!")                   # image.rs:1322-1326
        raster = self.get_image(film, write_fp16=write_fp16, bottom_up=True)
        with open(path, "wb") as f:
            f.write(b"PF\n%d %d\n-1\n" % (self.width, self.height))            # Rust `{}` of -1.0f64 prints "-1"
            f.write(np.ascontiguousarray(raster, dtype="<f4").tobytes())

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.sg_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sampler_fill(seed, n, pixel_index=0, sample_index=0, raw=False, device=0):
    """`IndependentSampler::get_1d` stream on the device (sampler.rs:123-125)."""
    lib = _ensure_init(device)
    out = np.zeros(n, np.float32)
    ffi.check(lib.sg_sampler_fill(int(seed) & 0xFFFFFFFFFFFFFFFF, int(raw), pixel_index, sample_index, n, out.ctypes.data),
              "sg_sampler_fill")
    return out


def generate_pyramid(image, wrap="repeat", device=0):
    """`Image::generate_pyramid` on the device (image.rs:699-787, incl. the resize of non-power-of-two images): `image` is an
    (H, W) or (H, W, C) array of LINEAR values; returns the list of (h, w, C) f32 levels, level 0 first -- what
    SceneBuilder.image_texture(levels=...) takes."""
    lib = _ensure_init(device)
    img = np.ascontiguousarray(image, np.float32)
    if img.ndim == 2:
        img = img[:, :, None]
    h, w, c = img.shape
    n_levels = C.c_int32(); n_texels = C.c_uint64(); rows = (ffi.SgImageLevel * 32)()
    ffi.check(lib.sg_image_pyramid_layout(w, h, c, C.byref(n_levels), rows, C.byref(n_texels)), "sg_image_pyramid_layout")
    out = np.zeros(n_texels.value, np.float32)
    ffi.check(lib.sg_image_generate_pyramid(img.ctypes.data, w, h, c, {"repeat": ffi.SG_WRAP_REPEAT, "clamp": ffi.SG_WRAP_CLAMP, "black": ffi.SG_WRAP_BLACK}[wrap],
                                            out.ctypes.data), "sg_image_generate_pyramid")
    return [out[r.offset:r.offset + r.res[0] * r.res[1] * c].reshape(r.res[1], r.res[0], c).copy() for r in rows[:n_levels.value]]


def create_integrator(name, parameters, scene, sampler=None, device=0, **kw):
    """integrator.rs:16-42 with the GPU entry added.  "path" stays the CPU integrator inside
    shimmer; here only the GPU backend exists, so any other name raises (the reference panics
    with `Unknown integrator {name}`)."""
    if name == "wavefront":
        return WavefrontPathIntegrator(scene, parameters, sampler, device=device, **kw)
    raise ffi.ShimmerGpuError(f"Unknown integrator {name}")


def render_gpu(scene, options: Options, parameters=None, sampler=None):
    """Sibling of `render_cpu` (render.rs:8-55): create the integrator, render, return it."""
    integ = create_integrator("wavefront", parameters, scene, sampler, device=options.gpu_device)
    integ.render(options)
    return integ


def write_pfm(path, rgb):
    """Image::write_pfm (image.rs:1333-1377): 'PF', w h, scale -1 (little endian), rows bottom-to-top, f32."""
    h, w, _ = rgb.shape
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1\n" % (w, h))
        f.write(np.ascontiguousarray(rgb[::-1], dtype="<f4").tobytes())
