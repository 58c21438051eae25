"""The C-ABI library loads and exports every symbol include/shimmer_gpu.h declares; without a GPU the
entry points fail with an error code instead of crashing or silently falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from shimmer_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "shimmer_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_match_binding_list():
    assert _declared() == sorted(ffi.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ffi.load_library()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.sg_abi_version() == ffi.SG_ABI_VERSION


def test_struct_sizes_match_the_header_layout():
    assert C.sizeof(ffi.SgBvhNode) == 32
    assert C.sizeof(ffi.SgPrimitive) == 16
    assert C.sizeof(ffi.SgMesh) == 32
    assert C.sizeof(ffi.SgSpectrum) == 32
    assert C.sizeof(ffi.SgMaterial) == 96
    assert C.sizeof(ffi.SgTexture) == 64 and C.sizeof(ffi.SgImageLevel) == 16
    assert C.sizeof(ffi.SgLight) == 64
    assert C.sizeof(ffi.SgObject) == 16 and C.sizeof(ffi.SgInstance) == 144 and C.sizeof(ffi.SgSceneDesc) == 736 and C.sizeof(ffi.SgSphere) == 160
    assert C.sizeof(ffi.SgFilmPixel) == 32
    assert C.sizeof(ffi.SgHit) == 32
    assert C.sizeof(ffi.SgRenderParams) == 48
    assert C.sizeof(ffi.SgStats) == 136          # ABI v9: + reduce_ms, d2h_ms, n_devices, rank
    assert C.sizeof(ffi.SgTextureMapping) == 112 and C.sizeof(ffi.SgTextureNode) == 32 and C.sizeof(ffi.SgMaterialTextures) == 48 and C.sizeof(ffi.SgDistribution2D) == 32 and C.sizeof(ffi.SgEnvMap) == 208
    assert np.dtype(ffi.SgBvhNode).itemsize == 32


def test_host_library_loads():
    h = ffi.load_host_library()
    assert hasattr(h, "sh_bvh_build") and hasattr(h, "sh_triangle_bounds")


def test_no_cpu_fallback_without_init(cornell64):
    """Before sg_init (or on a box with no GPU) every entry point reports an error."""
    import torch
    lib = ffi.load_library()
    h = C.c_void_p()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    rc = lib.sg_init(0)
    assert rc < 0
    assert b"CUDA" in lib.sg_last_error() or b"device" in lib.sg_last_error()
    rc = lib.sg_scene_create(cornell64.ptr(), C.byref(h))
    assert rc == -5 and h.value is None       # SG_ERR_NOT_INITIALIZED
    out = np.zeros(4, np.float32)
    assert lib.sg_sampler_fill(0, 1, 0, 0, 4, out.ctypes.data) == -5
    assert not out.any()


def test_integrator_registry_rejects_unknown_names(cornell64):
    from shimmer_b200 import create_integrator, ShimmerGpuError
    with pytest.raises(ShimmerGpuError, match="Unknown integrator"):
        create_integrator("path", {}, cornell64)       # the CPU integrator lives in shimmer, not here
    with pytest.raises(ShimmerGpuError):
        create_integrator("wavefront", {"lightsampler": "bvh"}, cornell64)   # light_sampler.rs:30-33 panics


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shimmer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liborc" not in txt and "import orc" not in txt and "oracle/" not in txt.replace("oracle/ is test", ""), f


def test_host_only_entry_points_work_without_a_gpu():
    """The multi-GPU bookkeeping that needs no device: the sample-range split the ranks use (must agree with the Python
    mirror rank by rank), the communicator query, and the not-initialised errors."""
    import ctypes as C
    from shimmer_b200.distributed import sample_range_for_rank
    lib = ffi.load_library()
    for begin, end in ((0, 1), (0, 7), (3, 64), (0, 1024), (5, 5)):
        for world in (1, 2, 3, 4, 8):
            for rank in range(world):
                b, e = C.c_int32(), C.c_int32()
                assert lib.sg_sample_range_for_rank(begin, end, rank, world, C.byref(b), C.byref(e)) == 0
                pb, pe = sample_range_for_rank(end - begin, rank, world)
                assert (b.value, e.value) == (begin + pb, begin + pe)
    b, e = C.c_int32(), C.c_int32()
    assert lib.sg_sample_range_for_rank(0, 4, 2, 2, C.byref(b), C.byref(e)) == -1          # rank out of range
    r, n = C.c_int(-1), C.c_int(-1)
    assert lib.sg_comm_rank(C.byref(r), C.byref(n)) == 0 and (r.value, n.value) == (0, 1)
    import torch
    if not torch.cuda.is_available():
        assert lib.sg_device_count() == 0
        assert lib.sg_comm_init_rank(C.create_string_buffer(ffi.SG_COMM_ID_BYTES), 0, 1) == -5   # SG_ERR_NOT_INITIALIZED
        assert lib.sg_film_reduce_device(C.c_void_p(8), 1, None) == -5
