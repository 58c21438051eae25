"""ctypes binding of the C ABI in include/shimmer_gpu.h.

This is the Python twin of the Rust `extern "C"` block shown in INTEGRATION.md: plain
pointers and sizes, no torch types.  The product path fails loudly when the CUDA library is
missing -- there is no CPU fallback (the CPU oracle under oracle/ is test infrastructure
and is never imported from this package).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SHIMMER_GPU_LIB") or os.path.join(_HERE, "libshimmer_gpu.so")   # override: A/B builds only
HOST_LIB_PATH = os.path.join(_HERE, "libshimmer_host.so")

SG_ABI_VERSION = 9

# enums (mirror include/shimmer_gpu.h)
SG_MESH_HAS_N, SG_MESH_HAS_UV, SG_MESH_HAS_S = 1, 2, 4
SG_MESH_BILINEAR = 32
SG_CAMERA_PERSPECTIVE, SG_CAMERA_ORTHOGRAPHIC = 0, 1
SG_MESH_REVERSE_ORIENTATION, SG_MESH_SWAPS_HANDEDNESS = 8, 16
SG_SPECTRUM_CONSTANT, SG_SPECTRUM_DENSE, SG_SPECTRUM_PIECEWISE_LINEAR, SG_SPECTRUM_BLACKBODY = 0, 1, 2, 3
SG_MATERIAL_DIFFUSE, SG_MATERIAL_CONDUCTOR, SG_MATERIAL_DIELECTRIC, SG_MATERIAL_COATED_DIFFUSE, SG_MATERIAL_THIN_DIELECTRIC = 0, 1, 2, 3, 4
SG_MATERIAL_COATED_CONDUCTOR, SG_MATERIAL_MIX = 5, 6
SG_MAT_REMAP_ROUGHNESS, SG_MAT_HAS_DISPLACEMENT, SG_MAT_CONDUCTOR_REFLECTANCE = 1, 2, 4
SG_LIGHT_DIFFUSE_AREA, SG_LIGHT_POINT, SG_LIGHT_UNIFORM_INFINITE, SG_LIGHT_DIFFUSE_AREA_SPHERE, SG_LIGHT_IMAGE_INFINITE = 0, 1, 2, 3, 4
SG_LIGHT_DIFFUSE_AREA_PATCH = 5
SG_MAPPING_SPHERICAL, SG_MAPPING_CYLINDRICAL, SG_MAPPING_PLANAR = 1, 2, 3
SG_INTEGRATOR_PATH, SG_INTEGRATOR_SIMPLE_PATH, SG_INTEGRATOR_RANDOM_WALK = 0, 1, 2
SG_SIMPLEPATH_SAMPLE_LIGHTS, SG_SIMPLEPATH_SAMPLE_BSDF = 1, 2
SG_OPT_DISABLE_PIXEL_JITTER, SG_OPT_DISABLE_WAVELENGTH_JITTER = 1, 2
SG_OPT_DISABLE_TEXTURE_FILTERING, SG_OPT_FORCE_DIFFUSE = 4, 8
SG_RENDER_COUNT_VISITS, SG_RENDER_TIME_KERNELS, SG_RENDER_OVERWRITE_FILM = 1, 2, 4
SG_RENDER_SPLIT_SAMPLES, SG_RENDER_REDUCE_FILM = 8, 16
SG_COMM_ID_BYTES = 128


class SgBvhNode(C.Structure):
    _fields_ = [("bmin", C.c_float * 3), ("bmax", C.c_float * 3), ("offset", C.c_uint32),
                ("n_prims", C.c_uint16), ("axis", C.c_uint8), ("pad", C.c_uint8)]


class SgPrimitive(C.Structure):
    _fields_ = [("mesh", C.c_uint32), ("tri", C.c_uint32), ("material", C.c_uint32), ("light", C.c_int32)]


SG_PRIM_INSTANCE = 0xffffffff
SG_PRIM_SPHERE = 0xfffffffe
SG_SCENE_FIX_INSTANCING = 1


class SgObject(C.Structure):
    _fields_ = [("first_node", C.c_uint32), ("n_nodes", C.c_uint32), ("first_prim", C.c_uint32), ("n_prims", C.c_uint32)]


class SgInstance(C.Structure):
    _fields_ = [("render_from_primitive", C.c_float * 16), ("primitive_from_render", C.c_float * 16), ("object", C.c_uint32),
                ("pad", C.c_uint32 * 3)]


class SgMesh(C.Structure):
    _fields_ = [("first_index", C.c_uint32), ("first_vertex", C.c_uint32), ("n_triangles", C.c_uint32),
                ("n_vertices", C.c_uint32), ("flags", C.c_uint32), ("pad", C.c_uint32 * 3)]


class SgSpectrum(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n", C.c_int32), ("lambda_min", C.c_int32), ("c", C.c_float),
                ("scale", C.c_float), ("off_a", C.c_uint32), ("off_b", C.c_uint32), ("pad", C.c_uint32)]


class SgMaterial(C.Structure):
    _fields_ = [("kind", C.c_int32), ("spec_a", C.c_int32), ("spec_b", C.c_int32), ("flags", C.c_int32),
                ("u_roughness", C.c_float), ("v_roughness", C.c_float), ("displacement", C.c_float), ("spec_c", C.c_int32),
                ("thickness", C.c_float), ("g", C.c_float), ("max_depth", C.c_int32), ("n_samples", C.c_int32),
                ("tex_reflectance", C.c_int32), ("tex_displacement", C.c_int32), ("pad2", C.c_int32 * 2),
                ("spec_d", C.c_int32), ("u_roughness2", C.c_float), ("v_roughness2", C.c_float), ("normal_map", C.c_int32),
                ("mix_materials", C.c_int32 * 2), ("mix_amount", C.c_float), ("tex_mix_amount", C.c_int32)]


class SgMaterialTextures(C.Structure):
    _fields_ = [("u_roughness", C.c_int32), ("v_roughness", C.c_int32), ("spec_a", C.c_int32), ("spec_b", C.c_int32), ("spec_d", C.c_int32),
                ("thickness", C.c_int32), ("g", C.c_int32), ("u_roughness2", C.c_int32), ("v_roughness2", C.c_int32), ("pad", C.c_int32 * 3)]
    NAMES = ("u_roughness", "v_roughness", "spec_a", "spec_b", "spec_d", "thickness", "g", "u_roughness2", "v_roughness2")


class SgImageLevel(C.Structure):
    _fields_ = [("offset", C.c_uint32), ("res", C.c_int32 * 2), ("pad", C.c_uint32)]


class SgTexture(C.Structure):
    _fields_ = [("n_channels", C.c_int32), ("n_levels", C.c_int32), ("first_level", C.c_uint32), ("wrap", C.c_int32),
                ("filter", C.c_int32), ("max_anisotropy", C.c_float), ("scale", C.c_float), ("invert", C.c_int32),
                ("su", C.c_float), ("sv", C.c_float), ("du", C.c_float), ("dv", C.c_float),
                ("spectrum_type", C.c_int32), ("mapping", C.c_int32), ("kind", C.c_int32), ("node", C.c_int32)]


SG_TEXTURE_IMAGE, SG_TEXTURE_CONSTANT, SG_TEXTURE_SCALED, SG_TEXTURE_MIX, SG_TEXTURE_DIRECTION_MIX = range(5)
SG_MAX_TEXTURE_DEPTH = 3


class SgTextureNode(C.Structure):
    _fields_ = [("tex1", C.c_int32), ("tex2", C.c_int32), ("amount", C.c_int32), ("spectrum", C.c_int32),
                ("value", C.c_float), ("dir", C.c_float * 3)]


class SgTextureMapping(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture_from_render", C.c_float * 16), ("vs", C.c_float * 3), ("vt", C.c_float * 3),
                ("ds", C.c_float), ("dt", C.c_float), ("pad", C.c_int32 * 3)]


class SgDistribution2D(C.Structure):
    _fields_ = [("nu", C.c_int32), ("nv", C.c_int32), ("func_off", C.c_uint32), ("cdf_off", C.c_uint32),
                ("marg_func_off", C.c_uint32), ("marg_cdf_off", C.c_uint32), ("marg_integral", C.c_float), ("pad", C.c_uint32)]


class SgEnvMap(C.Structure):
    _fields_ = [("render_from_light", C.c_float * 16), ("light_from_render", C.c_float * 16), ("texel_offset", C.c_uint64),
                ("res", C.c_int32), ("pad", C.c_int32), ("distribution", SgDistribution2D), ("compensated", SgDistribution2D)]


SG_WRAP_REPEAT, SG_WRAP_BLACK, SG_WRAP_CLAMP = 0, 1, 2
SG_FILTER_POINT, SG_FILTER_BILINEAR, SG_FILTER_TRILINEAR, SG_FILTER_EWA = 0, 1, 2, 3
SG_SPECTRUM_TYPE_ALBEDO, SG_SPECTRUM_TYPE_UNBOUNDED = 0, 1


class SgLight(C.Structure):
    _fields_ = [("kind", C.c_int32), ("spectrum", C.c_int32), ("scale", C.c_float), ("two_sided", C.c_int32),
                ("mesh", C.c_uint32), ("tri", C.c_uint32), ("area", C.c_float), ("pos", C.c_float * 3),
                ("scene_center", C.c_float * 3), ("scene_radius", C.c_float), ("pad", C.c_float * 2)]


class SgCamera(C.Structure):
    _fields_ = [("camera_from_raster", C.c_float * 16), ("render_from_camera", C.c_float * 16),
                ("camera_from_render", C.c_float * 16), ("dx_camera", C.c_float * 3), ("dy_camera", C.c_float * 3),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float), ("shutter_open", C.c_float),
                ("shutter_close", C.c_float),
                ("min_pos_differential_x", C.c_float * 3), ("min_pos_differential_y", C.c_float * 3),
                ("min_dir_differential_x", C.c_float * 3), ("min_dir_differential_y", C.c_float * 3),
                ("kind", C.c_int32), ("pad", C.c_int32)]


class SgFilm(C.Structure):
    _fields_ = [("full_resolution", C.c_int32 * 2), ("pixel_bounds", C.c_int32 * 4), ("filter_radius", C.c_float * 2),
                ("r_bar", C.c_int32), ("g_bar", C.c_int32), ("b_bar", C.c_int32), ("imaging_ratio", C.c_float),
                ("max_component_value", C.c_float), ("output_rgb_from_sensor_rgb", C.c_float * 9)]


class SgSphere(C.Structure):
    _fields_ = [("render_from_object", C.c_float * 16), ("object_from_render", C.c_float * 16),
                ("radius", C.c_float), ("z_min", C.c_float), ("z_max", C.c_float), ("theta_z_min", C.c_float),
                ("theta_z_max", C.c_float), ("phi_max", C.c_float), ("flags", C.c_uint32), ("pad", C.c_uint32)]


class SgSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32),
                ("n_nodes", C.c_uint32), ("nodes", C.POINTER(SgBvhNode)),
                ("n_primitives", C.c_uint32), ("primitives", C.POINTER(SgPrimitive)),
                ("n_top_nodes", C.c_uint32), ("n_top_primitives", C.c_uint32),
                ("n_objects", C.c_uint32), ("objects", C.POINTER(SgObject)),
                ("n_instances", C.c_uint32), ("instances", C.POINTER(SgInstance)),
                ("n_spheres", C.c_uint32), ("spheres", C.POINTER(SgSphere)),
                ("scene_flags", C.c_uint32),
                ("n_meshes", C.c_uint32), ("meshes", C.POINTER(SgMesh)),
                ("n_indices", C.c_uint32), ("indices", C.POINTER(C.c_uint32)),
                ("n_vertices", C.c_uint32), ("p", C.POINTER(C.c_float)),
                ("n", C.POINTER(C.c_float)), ("uv", C.POINTER(C.c_float)), ("s", C.POINTER(C.c_float)),
                ("n_spectra", C.c_uint32), ("spectra", C.POINTER(SgSpectrum)),
                ("n_pool", C.c_uint32), ("spectrum_pool", C.POINTER(C.c_float)),
                ("n_materials", C.c_uint32), ("materials", C.POINTER(SgMaterial)),
                ("n_lights", C.c_uint32), ("lights", C.POINTER(SgLight)),
                ("n_textures", C.c_uint32), ("textures", C.POINTER(SgTexture)),
                ("n_image_levels", C.c_uint32), ("image_levels", C.POINTER(SgImageLevel)),
                ("n_texels", C.c_uint64), ("texels", C.POINTER(C.c_float)),
                ("mip_filter_lut", C.POINTER(C.c_float)),
                ("rgb2spec_res", C.c_uint32), ("rgb2spec_scale", C.POINTER(C.c_float)), ("rgb2spec_data", C.POINTER(C.c_float)),
                ("n_texture_mappings", C.c_uint32), ("texture_mappings", C.POINTER(SgTextureMapping)),
                ("n_env_maps", C.c_uint32), ("env_maps", C.POINTER(SgEnvMap)),
                ("n_texture_nodes", C.c_uint32), ("texture_nodes", C.POINTER(SgTextureNode)),
                ("material_textures", C.POINTER(SgMaterialTextures)),
                ("camera", SgCamera), ("film", SgFilm)]


class SgRenderParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("samples_per_pixel", C.c_int32), ("sample_begin", C.c_int32),
                ("sample_end", C.c_int32), ("max_depth", C.c_int32), ("regularize", C.c_int32),
                ("option_flags", C.c_uint32), ("max_paths_in_flight", C.c_int32), ("flags", C.c_int32),
                ("integrator", C.c_int32), ("integrator_flags", C.c_int32)]


class SgFilmPixel(C.Structure):
    _fields_ = [("rgb_sum", C.c_double * 3), ("weight_sum", C.c_double)]


class SgStats(C.Structure):
    _fields_ = [("camera_paths", C.c_uint64), ("closest_hit_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("render_ms", C.c_double), ("trace_ms", C.c_double),
                ("closest_nodes", C.c_uint64), ("closest_tris", C.c_uint64), ("closest_launches", C.c_uint64),
                ("shadow_launches", C.c_uint64), ("closest_ms", C.c_double), ("shadow_ms", C.c_double),
                ("reduce_ms", C.c_double), ("d2h_ms", C.c_double), ("n_devices", C.c_uint32), ("rank", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SgHit(C.Structure):
    _fields_ = [("prim", C.c_int32), ("t", C.c_float), ("b0", C.c_float), ("b1", C.c_float), ("b2", C.c_float),
                ("ng", C.c_float * 3)]


# every symbol include/shimmer_gpu.h declares; tests/test_abi.py checks the library exports all
ABI_SYMBOLS = ["sg_init", "sg_shutdown", "sg_last_error", "sg_abi_version", "sg_scene_create", "sg_scene_destroy",
               "sg_render", "sg_render_device", "sg_trace", "sg_trace_device", "sg_sampler_fill", "sg_camera_rays",
               "sg_film_develop", "sg_texture_eval", "sg_texture_eval_p", "sg_texture_eval_ctx", "sg_film_get_image", "sg_image_pyramid_layout",
               "sg_image_generate_pyramid", "sg_init_multi", "sg_device_count", "sg_comm_get_unique_id", "sg_comm_init_rank",
               "sg_comm_destroy", "sg_comm_rank", "sg_sample_range_for_rank", "sg_film_reduce_device"]


class ShimmerGpuError(RuntimeError):
    pass


_lib = None
_host = None


def load_library():
    """dlopen libshimmer_gpu.so and declare the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShimmerGpuError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The GPU path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64
    fp = C.POINTER(C.c_float)
    lib.sg_init.argtypes = [C.c_int]; lib.sg_init.restype = C.c_int
    lib.sg_shutdown.argtypes = []; lib.sg_shutdown.restype = C.c_int
    lib.sg_init_multi.argtypes = [C.POINTER(C.c_int), C.c_int]; lib.sg_init_multi.restype = C.c_int
    lib.sg_device_count.argtypes = []; lib.sg_device_count.restype = C.c_int
    lib.sg_comm_get_unique_id.argtypes = [vp]; lib.sg_comm_get_unique_id.restype = C.c_int
    lib.sg_comm_init_rank.argtypes = [vp, C.c_int, C.c_int]; lib.sg_comm_init_rank.restype = C.c_int
    lib.sg_comm_destroy.argtypes = []; lib.sg_comm_destroy.restype = C.c_int
    lib.sg_comm_rank.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]; lib.sg_comm_rank.restype = C.c_int
    lib.sg_sample_range_for_rank.argtypes = [i32, i32, C.c_int, C.c_int, C.POINTER(i32), C.POINTER(i32)]; lib.sg_sample_range_for_rank.restype = C.c_int
    lib.sg_film_reduce_device.argtypes = [vp, i64, vp]; lib.sg_film_reduce_device.restype = C.c_int
    lib.sg_last_error.argtypes = []; lib.sg_last_error.restype = C.c_char_p
    lib.sg_abi_version.argtypes = []; lib.sg_abi_version.restype = C.c_int
    lib.sg_scene_create.argtypes = [C.POINTER(SgSceneDesc), C.POINTER(vp)]; lib.sg_scene_create.restype = C.c_int
    lib.sg_scene_destroy.argtypes = [vp]; lib.sg_scene_destroy.restype = C.c_int
    lib.sg_render.argtypes = [vp, C.POINTER(SgRenderParams), vp, C.POINTER(SgStats)]; lib.sg_render.restype = C.c_int
    lib.sg_render_device.argtypes = [vp, C.POINTER(SgRenderParams), vp, C.POINTER(SgStats), vp]
    lib.sg_render_device.restype = C.c_int
    lib.sg_trace.argtypes = [vp, i64, vp, vp, vp, C.c_int, vp, C.POINTER(SgStats)]; lib.sg_trace.restype = C.c_int
    lib.sg_trace_device.argtypes = [vp, i64, vp, vp, vp, C.c_int, vp, C.POINTER(SgStats), vp]
    lib.sg_trace_device.restype = C.c_int
    lib.sg_sampler_fill.argtypes = [u64, C.c_int, u32, u32, i64, vp]; lib.sg_sampler_fill.restype = C.c_int
    lib.sg_camera_rays.argtypes = [vp, C.POINTER(SgRenderParams), i64, vp, vp, vp, vp]
    lib.sg_camera_rays.restype = C.c_int
    lib.sg_film_develop.argtypes = [vp, vp, i64, vp]; lib.sg_film_develop.restype = C.c_int
    lib.sg_film_get_image.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_uint32, vp]; lib.sg_film_get_image.restype = C.c_int
    lib.sg_texture_eval.argtypes = [vp, C.c_int, C.c_int, i64, vp, vp, vp]; lib.sg_texture_eval.restype = C.c_int
    lib.sg_texture_eval_p.argtypes = [vp, C.c_int, C.c_int, i64, vp, vp, vp, vp]; lib.sg_texture_eval_p.restype = C.c_int
    lib.sg_texture_eval_ctx.argtypes = [vp, C.c_int, C.c_int, i64, vp, vp, vp, vp, vp]; lib.sg_texture_eval_ctx.restype = C.c_int
    lib.sg_image_pyramid_layout.argtypes = [C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]; lib.sg_image_pyramid_layout.restype = C.c_int
    lib.sg_image_generate_pyramid.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp]; lib.sg_image_generate_pyramid.restype = C.c_int
    _lib = lib
    return lib


def load_host_library():
    """Host-side helpers (BVH build standing in for shimmer's Rust BvhAggregate::new)."""
    global _host
    if _host is not None:
        return _host
    if not os.path.exists(HOST_LIB_PATH):
        raise ShimmerGpuError(f"{HOST_LIB_PATH} not found: run __graft_entry__.build()")
    h = C.CDLL(HOST_LIB_PATH)
    h.sh_bvh_build.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]; h.sh_bvh_build.restype = C.c_int64
    h.sh_triangle_bounds.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]; h.sh_triangle_bounds.restype = None
    h.sh_spectrum_lut.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]; h.sh_spectrum_lut.restype = C.c_int
    _host = h
    return h


def check(rc, what="shimmer_gpu call"):
    """Turn a non-zero status into an exception -- the reference panics (integrator.rs:36)."""
    if rc != 0:
        msg = load_library().sg_last_error()
        raise ShimmerGpuError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
