//! `render::render_gpu`: the sibling of `render_cpu` (render.rs:8-55).  Identical up to the point where `render_cpu` calls
//! `scene.create_integrator(..)` + `integrator.render(options)`; from there the hot path runs on the GPU(s):
//!
//!   flatten (host, once)  ->  sg_scene_create (upload, once per GPU)  ->  sg_render  ->  film sums into RgbFilm.pixels
//!   ->  the unchanged `Film::write_image` (film.rs:709-713)
//!
//! Multi-GPU: `SHIMMER_GPUS=0,1,2,3` (or `--gpus` if the CLI grows one) makes this process drive several devices; the library
//! replicates the scene, splits the sample indices across the devices -- the analogue of the tile fan-out over rayon's pool,
//! integrator.rs:235-245 -- and sums the films with one NCCL reduce.  Nothing in this file changes for that.
use std::collections::HashMap;
use std::sync::{Arc, Mutex};

use log::{info, warn};
use string_interner::StringInterner;

use super::ffi::*;
use super::flatten::{FlatScene, FlattenOptions};
use super::{check, GpuError};
use crate::color::ColorEncodingCache;
use crate::film::{Film, FilmI};
use crate::image::ImageMetadata;
use crate::loading::scene::BasicScene;
use crate::mipmap::MIPMap;
use crate::options::Options;
use crate::spectra::Spectrum;
use crate::texture::TexInfo;
use crate::Float;

/// Integrator parameters the GPU path needs, read the way `create_integrator` reads them (integrator.rs:16-42,188-192).
struct IntegratorParams {
    kind: i32,
    flags: i32,
    max_depth: i32,
    regularize: bool,
}

fn integrator_params(scene: &mut BasicScene, string_interner: &StringInterner) -> Result<IntegratorParams, GpuError> {
    let ent = scene.integrator.as_mut().expect("integrator entity");          // scene.rs:888-907 unwraps the same way
    let name = string_interner.resolve(ent.name).unwrap().to_owned();
    let p = &mut ent.parameters;
    let max_depth = p.get_one_int("maxdepth", 5);
    let light_sampler = p.get_one_string("lightsampler", "uniform");
    if light_sampler != "uniform" {
        // light_sampler.rs:30-33 knows nothing else either ("bvh" / "power" panic there)
        return Err(GpuError::Unsupported(format!("light sampler {}", light_sampler)));
    }
    let (kind, flags, regularize) = match name.as_str() {
        "path" => (SG_INTEGRATOR_PATH, 0, p.get_one_bool("regularize", false)),
        "simplepath" => {
            let sl = p.get_one_bool("samplelights", true);
            let sb = p.get_one_bool("samplebsdf", true);
            (SG_INTEGRATOR_SIMPLE_PATH, (if sl { SG_SIMPLEPATH_SAMPLE_LIGHTS } else { 0 }) | (if sb { SG_SIMPLEPATH_SAMPLE_BSDF } else { 0 }), false)
        }
        "randomwalk" => (SG_INTEGRATOR_RANDOM_WALK, 0, false),
        other => return Err(GpuError::Unsupported(format!("integrator {}", other))),
    };
    Ok(IntegratorParams { kind, flags, max_depth, regularize })
}

fn option_flags(o: &Options) -> u32 {
    (if o.disable_pixel_jitter { SG_OPT_DISABLE_PIXEL_JITTER } else { 0 })
        | (if o.disable_wavelength_jitter { SG_OPT_DISABLE_WAVELENGTH_JITTER } else { 0 })
        | (if o.disable_texture_filtering { SG_OPT_DISABLE_TEXTURE_FILTERING } else { 0 })
        | (if o.force_diffuse { SG_OPT_FORCE_DIFFUSE } else { 0 })
}

fn gpu_list() -> Vec<i32> {
    match std::env::var("SHIMMER_GPUS") {
        Ok(s) => s.split(',').filter_map(|t| t.trim().parse().ok()).collect(),
        Err(_) => vec![0],
    }
}

/// RAII for the opaque scene handle: sg_scene_destroy on every exit path.
struct SceneHandle(*mut SgScene);
impl Drop for SceneHandle {
    fn drop(&mut self) {
        if !self.0.is_null() {
            unsafe { sg_scene_destroy(self.0) };
        }
    }
}

pub fn render_gpu(
    mut scene: Box<BasicScene>,
    options: &Options,
    string_interner: &mut StringInterner,
    cached_spectra: &mut HashMap<String, Arc<Spectrum>>,
    texture_cache: &&Arc<Mutex<HashMap<TexInfo, Arc<MIPMap>>>>,
    gamma_encoding_cache: &mut ColorEncodingCache,
) {
    // ---- identical to render_cpu (render.rs:16-49): textures, lights, materials, aggregate, camera, sampler
    let media = HashMap::new();
    let textures = scene.create_textures(cached_spectra, string_interner, options, texture_cache, gamma_encoding_cache);
    let (lights, shape_index_to_area_lights) = scene.create_lights(&textures, &string_interner, options);
    let (named_materials, materials) = scene.create_materials(&textures, &string_interner, cached_spectra, options);
    let accelerator = scene.create_aggregate(&textures, &shape_index_to_area_lights, &media, &named_materials, &materials, &string_interner, options);
    let camera = scene.get_camera().unwrap();
    let sampler = scene.get_sampler().unwrap();

    let outcome = (|| -> Result<SgStats, GpuError> {
        let ip = integrator_params(&mut scene, string_interner)?;
        info!("Flattening the scene for the GPU...");
        let flat = FlatScene::new(&accelerator, &lights, &camera, &sampler, FlattenOptions::default())?;
        let desc = flat.desc();
        let gpus = gpu_list();
        check(unsafe { sg_init_multi(gpus.as_ptr(), gpus.len() as i32) })?;
        let mut h = SceneHandle(std::ptr::null_mut());
        check(unsafe { sg_scene_create(&desc, &mut h.0) })?;          // everything is in HBM now; `flat` may go
        drop(flat);

        let spp = options.pixel_samples.unwrap_or(sampler.samples_per_pixel());          // --spp overrides the scene file (main.rs)
        let params = SgRenderParams {
            seed: options.seed as u64,
            samples_per_pixel: spp,
            sample_begin: 0,
            sample_end: spp,
            max_depth: ip.max_depth,
            regularize: ip.regularize as i32,
            option_flags: option_flags(options),
            max_paths_in_flight: 0,
            flags: SG_RENDER_OVERWRITE_FILM,
            integrator: ip.kind,
            integrator_flags: ip.flags,
        };
        let film_arc: &Arc<Film> = camera.get_film();
        let bounds = film_arc.pixel_bounds();
        let n_pixels = bounds.area() as usize;
        let mut sums = vec![SgFilmPixel { rgb_sum: [0.0; 3], weight_sum: 0.0 }; n_pixels];
        let mut stats = SgStats::default();
        info!("Rendering on {} GPU(s)...", gpus.len());
        check(unsafe { sg_render(h.0, &params, sums.as_mut_ptr(), &mut stats) })?;

        // ---- film copy-back.  RgbFilm.pixels is a Vec2d<RgbFilmPixel> (film.rs:465,470-479) whose `data` is row-major over the
        // pixel bounds, index (y - y0) * width + (x - x0) (vec2d.rs:24-38) -- exactly the order sg_render fills.  rgb_splat
        // stays 0 (no splats on this path).  The film sits behind the camera's Arc<Film>; the CPU integrator writes through
        // `Arc::get_mut_unchecked` (integrator.rs:287-295, nightly feature `get_mut_unchecked`, lib.rs:3) and so does this.
        let mut film_clone = film_arc.clone();
        unsafe {
            let Film::RgbFilm(rgb) = Arc::get_mut_unchecked(&mut film_clone);
            assert_eq!(rgb.pixels.data.len(), n_pixels);
            for (dst, src) in rgb.pixels.data.iter_mut().zip(sums.iter()) {
                dst.rgb_sum = src.rgb_sum;
                dst.weight_sum = src.weight_sum;
            }
        }
        // ---- the unchanged output stage (integrator.rs:311-319): splat_scale = 1 / spp
        let mut metadata = ImageMetadata::default();
        film_arc.write_image(&mut metadata, 1.0 / spp as Float).unwrap();
        Ok(stats)
    })();

    match outcome {
        Ok(st) => info!(
            "GPU render: {:.1} ms on {} device(s) ({:.1} Mpaths/s, {:.1} Mrays/s), film reduce {:.2} ms, film D2H {:.2} ms",
            st.render_ms, st.n_devices, st.camera_paths as f64 / st.render_ms / 1e3,
            (st.closest_hit_rays + st.shadow_rays) as f64 / st.render_ms / 1e3, st.reduce_ms, st.d2h_ms
        ),
        Err(GpuError::Unsupported(why)) => {
            // The LIBRARY never falls back; the host decides.  The scene objects are already built, so the CPU path continues
            // from where render_cpu would be (render.rs:51-54).
            warn!("--wavefront: {} is not on the GPU path; rendering on the CPU instead", why);
            let mut integrator = scene.create_integrator(camera, sampler, accelerator, lights, &string_interner);
            integrator.render(options);
        }
        Err(GpuError::Library(rc, msg)) => panic!("shimmer_gpu failed ({}): {}", rc, msg),   // the reference's convention: panic (integrator.rs:36)
    }
}
