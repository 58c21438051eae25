"""shimmer_b200 -- B200-native wavefront path tracing backend for shimmer's `path` integrator.

Layout (only what the hot path needs):
  csrc/          CUDA kernels (sm_100a) + the C ABI of include/shimmer_gpu.h + host BVH helper
  ffi.py         ctypes binding of the C ABI
  host.py        scene flattening that stands in for shimmer's Rust host (camera, film, spectra, BVH)
  scenes.py      deterministic generators for the BASELINE.json configurations
  integrator.py  mirror of the reference's Integrator API (`create_integrator`, `.render(options)`)
"""
from .ffi import ShimmerGpuError  # noqa: F401
from .integrator import Options, WavefrontPathIntegrator, create_integrator, generate_pyramid, render_gpu  # noqa: F401
