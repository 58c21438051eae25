//! GPU backend of the `path` integrator: `--wavefront` (options.rs, main.rs:89-91,152-155) routes
//! `render::render_gpu` here instead of `render_cpu`'s `integrator.render(options)` (render.rs:51-54).
//!
//! * `ffi`        -- `extern "C"` bindings of include/shimmer_gpu.h (ABI v9), field for field
//! * `flatten`    -- walks the objects `render_cpu` builds (the `Arc<Primitive>` tree, lights, materials, textures,
//!                   camera, film, sampler) and lays them out as the POD arrays of `SgSceneDesc`
//! * `render_gpu` -- upload, render (1 or n GPUs), copy the film sums back into `RgbFilm`, `write_image`
//!
//! Nothing here falls back to the CPU silently: content the device path does not cover makes `flatten` return
//! `Err(Unsupported)` and `render::render_gpu` hands the scene to `render_cpu` with a warning.
pub mod ffi;
pub mod flatten;
pub mod render_gpu;

#[derive(Debug)]
pub enum GpuError {
    /// scene content that is not on the GPU path (the caller falls back to render_cpu)
    Unsupported(String),
    /// a non-zero status from libshimmer_gpu.so with sg_last_error()
    Library(i32, String),
}

pub(crate) fn check(rc: i32) -> Result<(), GpuError> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { std::ffi::CStr::from_ptr(ffi::sg_last_error()) }
        .to_string_lossy()
        .into_owned();
    if rc == ffi::SG_ERR_UNSUPPORTED {
        Err(GpuError::Unsupported(msg))
    } else {
        Err(GpuError::Library(rc, msg))
    }
}
