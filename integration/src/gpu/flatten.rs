//! Scene flattener: the objects `render_cpu` builds (render.rs:19-49) -> the POD arrays of `SgSceneDesc`.
//!
//! This is the only place shimmer and libshimmer_gpu.so touch.  It walks
//!   * the top-level `Arc<Primitive>` returned by `BasicScene::create_aggregate` (scene.rs:868-885): a
//!     `Primitive::BvhAggregate` whose ordered `primitives` are `Simple` / `Geometric` (one shape each,
//!     primitive.rs:67-134) or `Transformed` (an object instance, primitive.rs:136-176; built in scene.rs:814-866);
//!   * each instanced object: either one bare primitive or its own `BvhAggregate` (scene.rs:821-833);
//!   * shapes (`Shape::{Triangle, BilinearPatch, Sphere}`, shape.rs:60-64) -- triangle / patch meshes are shared
//!     `Arc<TriangleMesh>` / `Arc<BilinearPatchMesh>` and are uploaded ONCE (dedup by `Arc` pointer);
//!   * materials (`Material::{Single, Mix}`, material.rs:48-108), their textures (`FloatTexture` / `SpectrumTexture`
//!     trees, texture.rs:88-94,411-417) and spectra (spectrum.rs:39-48), all deduplicated by `Arc` pointer;
//!   * the light list in light-sampler order (`Arc<Vec<Arc<Light>>>`, light_sampler.rs:78-112) -- a primitive's
//!     `area_light` is found in that list by pointer;
//!   * camera (camera.rs:54-57), film + sensor (film.rs:410-417,456-466,754-765), sampler (sampler.rs:82-87).
//!
//! Field access: the reference keeps most of these fields private.  `patches/visibility.md` lists the
//! `pub(crate)` changes (one keyword per line, no behaviour change) this module needs; everything below is written
//! against those.  Instancing semantics: `scene_flags = 0` reproduces `TransformedPrimitive` literally (shadow rays
//! through the FORWARD transform, primitive.rs:172-175; interaction vectors through the inverse, transform.rs:573-609);
//! `FlattenOptions::fix_instancing` selects pbrt's semantics instead.
use std::collections::HashMap;
use std::sync::Arc;

use super::ffi::*;
use super::GpuError;
use crate::aggregate::BvhAggregate;
use crate::camera::Camera;
use crate::film::{Film, FilmI};
use crate::filter::FilterI;
use crate::image::Image;
use crate::light::Light;
use crate::material::{Material, SingleMaterial};
use crate::loading::paramdict::SpectrumType;
use crate::mipmap::{FilterFunction, MIPMap, MIP_FILTER_LUT};
use crate::rgb_to_spectra::Gamut;
use crate::primitive::Primitive;
use crate::sampler::Sampler;
use crate::shape::{BilinearPatchMesh, Shape, TriangleMesh};
use crate::spectra::spectrum::{DenselySampledSpectrum, Spectrum};
use crate::texture::{FloatTexture, ImageTextureBase, SpectrumTexture, TextureMapping2D};
use crate::transform::Transform;

#[derive(Default, Clone, Copy)]
pub struct FlattenOptions {
    /// pbrt's instancing semantics instead of the reference's literal ones (see module docs)
    pub fix_instancing: bool,
}

/// Owns every array `desc` points into; keep it alive until `sg_scene_create` has returned (the library copies
/// everything to HBM and borrows nothing afterwards).
#[derive(Default)]
pub struct FlatScene {
    nodes: Vec<SgBvhNode>,
    primitives: Vec<SgPrimitive>,
    objects: Vec<SgObject>,
    instances: Vec<SgInstance>,
    spheres: Vec<SgSphere>,
    meshes: Vec<SgMesh>,
    indices: Vec<u32>,
    p: Vec<f32>,
    n: Vec<f32>,
    uv: Vec<f32>,
    s: Vec<f32>,
    spectra: Vec<SgSpectrum>,
    pool: Vec<f32>,
    materials: Vec<SgMaterial>,
    material_textures: Vec<SgMaterialTextures>,
    any_material_texture: bool,
    lights: Vec<SgLight>,
    textures: Vec<SgTexture>,
    image_levels: Vec<SgImageLevel>,
    texels: Vec<f32>,
    texture_mappings: Vec<SgTextureMapping>,
    texture_nodes: Vec<SgTextureNode>,
    env_maps: Vec<SgEnvMap>,
    mip_lut: Vec<f32>,
    rgb2spec_scale: Vec<f32>,
    rgb2spec_data: Vec<f32>,
    rgb2spec_res: u32,
    n_top_nodes: u32,
    n_top_primitives: u32,
    scene_flags: u32,
    camera: Option<SgCamera>,
    film: Option<SgFilm>,
    pub samples_per_pixel: i32,
    pub sampler_seed: u64,
    // dedup maps, keyed by Arc pointer
    mesh_ids: HashMap<usize, u32>,
    spectrum_ids: HashMap<usize, i32>,
    material_ids: HashMap<usize, u32>,
    float_tex_ids: HashMap<usize, i32>,
    spectrum_tex_ids: HashMap<usize, i32>,
    mip_ids: HashMap<usize, (u32, i32, i32)>, // Arc<MIPMap> -> (first_level, n_levels, n_channels)
    object_ids: HashMap<usize, u32>,
    light_ids: HashMap<usize, i32>,
    // area lights whose primitive has not been met yet while walking the aggregate (light indices)
    area_lights_pending: std::collections::HashSet<usize>,
}

fn unsupported<T>(what: impl Into<String>) -> Result<T, GpuError> {
    Err(GpuError::Unsupported(what.into()))
}
fn m16(t: &crate::square_matrix::SquareMatrix<4>) -> [f32; 16] {
    let mut o = [0.0f32; 16];
    for r in 0..4 {
        for c in 0..4 {
            o[4 * r + c] = t.m[r][c];
        }
    }
    o
}
fn u32_checked(v: usize, what: &str) -> Result<u32, GpuError> {
    u32::try_from(v).map_err(|_| GpuError::Unsupported(format!("{} does not fit the 32-bit indices of the GPU scene", what)))
}
fn ptr_key<T: ?Sized>(a: &Arc<T>) -> usize {
    Arc::as_ptr(a) as *const () as usize
}

impl FlatScene {
    /// Flattens the scene exactly as `render_cpu` has built it.  `lights` is the vector handed to the light sampler
    /// (its order IS the sampler's order: `UniformLightSampler::sample_light` indexes it, light_sampler.rs:91-103).
    pub fn new(
        aggregate: &Arc<Primitive>,
        lights: &Arc<Vec<Arc<Light>>>,
        camera: &Camera,
        sampler: &Sampler,
        opts: FlattenOptions,
    ) -> Result<FlatScene, GpuError> {
        let mut f = FlatScene::default();
        f.scene_flags = if opts.fix_instancing { SG_SCENE_FIX_INSTANCING } else { 0 };
        f.mip_lut = MIP_FILTER_LUT.to_vec();                    // mipmap.rs:388-518 (EWA weights)
        f.flatten_lights(lights)?;
        let top = match aggregate.as_ref() {
            Primitive::BvhAggregate(a) => a,
            // create_accelerator (aggregate.rs:20-32) only ever returns a BvhAggregate
            _ => return unsupported("top-level primitive is not a BvhAggregate"),
        };
        // the top-level BVH first, then every object definition (their node / primitive ranges follow the top level)
        let mut pending_objects: Vec<Arc<Primitive>> = Vec::new();
        f.flatten_bvh(top, true, &mut pending_objects)?;
        f.n_top_nodes = f.nodes.len() as u32;
        f.n_top_primitives = f.primitives.len() as u32;
        let mut next = 0;
        while next < pending_objects.len() {
            let obj = pending_objects[next].clone();
            next += 1;
            f.flatten_object(&obj)?;
        }
        if !f.area_lights_pending.is_empty() {
            // every DiffuseAreaLight belongs to exactly one GeometricPrimitive (scene.rs:609-622 / :771-781 create them pairwise)
            return unsupported("an area light without a primitive in the aggregate");
        }
        f.flatten_camera(camera)?;
        f.flatten_film(camera.get_film())?;
        match sampler {
            Sampler::Independent(s) => {
                f.samples_per_pixel = s.samples_per_pixel;
                f.sampler_seed = s.seed;
            }
        }
        Ok(f)
    }

    /// The descriptor over the owned arrays.  Valid while `self` is alive and unmodified.
    pub fn desc(&self) -> SgSceneDesc {
        fn p<T>(v: &Vec<T>) -> *const T {
            if v.is_empty() { std::ptr::null() } else { v.as_ptr() }
        }
        SgSceneDesc {
            abi_version: SG_ABI_VERSION,
            n_nodes: self.nodes.len() as u32, nodes: p(&self.nodes),
            n_primitives: self.primitives.len() as u32, primitives: p(&self.primitives),
            n_top_nodes: self.n_top_nodes, n_top_primitives: self.n_top_primitives,
            n_objects: self.objects.len() as u32, objects: p(&self.objects),
            n_instances: self.instances.len() as u32, instances: p(&self.instances),
            n_spheres: self.spheres.len() as u32, spheres: p(&self.spheres),
            scene_flags: self.scene_flags,
            n_meshes: self.meshes.len() as u32, meshes: p(&self.meshes),
            n_indices: self.indices.len() as u32, indices: p(&self.indices),
            n_vertices: (self.p.len() / 3) as u32, p: p(&self.p), n: p(&self.n), uv: p(&self.uv), s: p(&self.s),
            n_spectra: self.spectra.len() as u32, spectra: p(&self.spectra),
            n_pool: self.pool.len() as u32, spectrum_pool: p(&self.pool),
            n_materials: self.materials.len() as u32, materials: p(&self.materials),
            n_lights: self.lights.len() as u32, lights: p(&self.lights),
            n_textures: self.textures.len() as u32, textures: p(&self.textures),
            n_image_levels: self.image_levels.len() as u32, image_levels: p(&self.image_levels),
            n_texels: self.texels.len() as u64, texels: p(&self.texels),
            mip_filter_lut: p(&self.mip_lut),
            rgb2spec_res: self.rgb2spec_res, rgb2spec_scale: p(&self.rgb2spec_scale), rgb2spec_data: p(&self.rgb2spec_data),
            n_texture_mappings: self.texture_mappings.len() as u32, texture_mappings: p(&self.texture_mappings),
            n_env_maps: self.env_maps.len() as u32, env_maps: p(&self.env_maps),
            n_texture_nodes: self.texture_nodes.len() as u32, texture_nodes: p(&self.texture_nodes),
            material_textures: if self.any_material_texture { p(&self.material_textures) } else { std::ptr::null() },
            camera: self.camera.expect("flatten_camera ran"),
            film: self.film.expect("flatten_film ran"),
        }
    }

    // ---------------------------------------------------------------------------------------- acceleration structure
    /// One `BvhAggregate` (aggregate.rs:39-44): its `LinearBvhNode`s verbatim (offsets stay relative to this BVH's own
    /// node / primitive ranges, which is what the ABI wants for object definitions and what the top level has anyway)
    /// and its ordered primitives.
    fn flatten_bvh(&mut self, bvh: &BvhAggregate, top_level: bool, pending: &mut Vec<Arc<Primitive>>) -> Result<(), GpuError> {
        for nd in bvh.nodes.iter() {
            let leaf = nd.n_primitives > 0;
            self.nodes.push(SgBvhNode {
                bmin: [nd.bounds.min.x, nd.bounds.min.y, nd.bounds.min.z],
                bmax: [nd.bounds.max.x, nd.bounds.max.y, nd.bounds.max.z],
                offset: u32_checked(if leaf { nd.primitive_offset } else { nd.second_child_offset }, "BVH offset")?,
                n_prims: nd.n_primitives,
                axis: nd.axis,
                pad: 0,
            });
        }
        for prim in bvh.primitives.iter() {
            self.flatten_primitive(prim, top_level, pending)?;
        }
        Ok(())
    }

    fn flatten_primitive(&mut self, prim: &Arc<Primitive>, top_level: bool, pending: &mut Vec<Arc<Primitive>>) -> Result<(), GpuError> {
        match prim.as_ref() {
            Primitive::Simple(sp) => self.push_shape(&sp.shape, &sp.material, None, top_level),
            Primitive::Geometric(gp) => self.push_shape(&gp.shape, &gp.material, gp.area_light.as_ref(), top_level),
            Primitive::Transformed(tp) => {
                if !top_level {
                    return unsupported("nested object instances (pbrt-v4 scene format forbids them)");
                }
                // the instanced object: dedup by pointer, flattened after the top level
                let key = ptr_key(&tp.primitive);
                let object = match self.object_ids.get(&key) {
                    Some(&id) => id,
                    None => {
                        let id = (self.object_ids.len()) as u32;
                        self.object_ids.insert(key, id);
                        pending.push(tp.primitive.clone());
                        id
                    }
                };
                let inst = self.instances.len() as u32;
                self.instances.push(SgInstance {
                    render_from_primitive: m16(&tp.render_from_primitive.m),
                    primitive_from_render: m16(&tp.render_from_primitive.m_inv),
                    object,
                    pad: [0; 3],
                });
                self.primitives.push(SgPrimitive { mesh: SG_PRIM_INSTANCE, tri: inst, material: 0, light: -1 });
                Ok(())
            }
            Primitive::BvhAggregate(_) => unsupported("an aggregate nested directly inside an aggregate"),
        }
    }

    /// One object definition (scene.rs:816-841): a `BvhAggregate` over its shapes, or the single primitive itself.
    /// Objects were numbered when their first instance was met; they are flattened in that order.
    fn flatten_object(&mut self, obj: &Arc<Primitive>) -> Result<(), GpuError> {
        let first_node = self.nodes.len() as u32;
        let first_prim = self.primitives.len() as u32;
        let mut none = Vec::new();
        match obj.as_ref() {
            Primitive::BvhAggregate(b) => self.flatten_bvh(b, false, &mut none)?,
            _ => self.flatten_primitive(obj, false, &mut none)?,
        }
        self.objects.push(SgObject {
            first_node,
            n_nodes: self.nodes.len() as u32 - first_node,
            first_prim,
            n_prims: self.primitives.len() as u32 - first_prim,
        });
        Ok(())
    }

    // ------------------------------------------------------------------------------------------------------ shapes
    fn push_shape(&mut self, shape: &Arc<Shape>, material: &Arc<Material>, area_light: Option<&Arc<Light>>, top_level: bool) -> Result<(), GpuError> {
        let material = self.material_id(material)?;
        let light: i32 = match area_light {
            None => -1,
            Some(l) => {
                if !top_level {
                    return unsupported("area lights inside object definitions");
                }
                *self.light_ids.get(&ptr_key(l)).ok_or_else(|| GpuError::Unsupported("a primitive's area light is not in the light list".into()))?
            }
        };
        let (mesh, tri) = match shape.as_ref() {
            Shape::Triangle(t) => (self.triangle_mesh_id(&t.mesh)?, t.tri_index as u32),
            Shape::BilinearPatch(b) => (self.patch_mesh_id(&b.mesh)?, u32_checked(b.blp_index, "patch index")?),
            Shape::Sphere(s) => {
                let id = self.spheres.len() as u32;
                self.spheres.push(SgSphere {
                    render_from_object: m16(&s.render_from_object.m),
                    object_from_render: m16(&s.object_from_render.m),
                    radius: s.radius, z_min: s.z_min, z_max: s.z_max,
                    theta_z_min: s.theta_z_min, theta_z_max: s.theta_z_max, phi_max: s.phi_max,
                    flags: (if s.reverse_orientation { SG_MESH_REVERSE_ORIENTATION } else { 0 })
                        | (if s.transform_swaps_handedness { SG_MESH_SWAPS_HANDEDNESS } else { 0 }),
                    pad: 0,
                });
                (SG_PRIM_SPHERE, id)
            }
        };
        if light >= 0 {
            // Complete the light row now that its shape has a (mesh, tri) address.  The light owns ITS OWN Arc<Shape>:
            // create_lights and create_aggregate each call Shape::create for the same scene entity (scene.rs:590-599, :745-751),
            // so the two shapes are distinct objects over identical geometry -- the light row addresses the primitive's copy.
            let li = light as usize;
            if !self.area_lights_pending.remove(&li) {
                return unsupported("one area light attached to two primitives");
            }
            let row = &mut self.lights[li];
            row.mesh = mesh;
            row.tri = tri;
            row.kind = match shape.as_ref() {
                Shape::Triangle(_) => SG_LIGHT_DIFFUSE_AREA,
                Shape::BilinearPatch(_) => SG_LIGHT_DIFFUSE_AREA_PATCH,
                Shape::Sphere(_) => SG_LIGHT_DIFFUSE_AREA_SPHERE,
            };
        }
        self.primitives.push(SgPrimitive { mesh, tri, material, light });
        Ok(())
    }

    fn triangle_mesh_id(&mut self, mesh: &Arc<TriangleMesh>) -> Result<u32, GpuError> {
        let key = ptr_key(mesh);
        if let Some(&id) = self.mesh_ids.get(&key) {
            return Ok(id);
        }
        let first_index = u32_checked(self.indices.len(), "index count")?;
        let first_vertex = u32_checked(self.p.len() / 3, "vertex count")?;
        for &i in mesh.vertex_indices.iter() {
            self.indices.push(u32_checked(i, "vertex index")?); // mesh-local numbers; `usize` in the reference
        }
        let mut flags = 0u32;
        self.push_vertices(first_vertex as usize, mesh.n_vertices,
                           mesh.p.iter().map(|q| [q.x, q.y, q.z]),
                           if mesh.n.is_empty() { None } else { flags |= SG_MESH_HAS_N; Some(mesh.n.iter().map(|q| [q.x, q.y, q.z]).collect()) },
                           if mesh.uv.is_empty() { None } else { flags |= SG_MESH_HAS_UV; Some(mesh.uv.iter().map(|q| [q.x, q.y]).collect()) },
                           if mesh.s.is_empty() { None } else { flags |= SG_MESH_HAS_S; Some(mesh.s.iter().map(|q| [q.x, q.y, q.z]).collect()) });
        if mesh.reverse_orientation { flags |= SG_MESH_REVERSE_ORIENTATION; }
        if mesh.transform_swaps_handedness { flags |= SG_MESH_SWAPS_HANDEDNESS; }
        let id = self.meshes.len() as u32;
        self.meshes.push(SgMesh { first_index, first_vertex, n_triangles: u32_checked(mesh.n_triangles, "triangle count")?,
                                  n_vertices: u32_checked(mesh.n_vertices, "vertex count")?, flags, pad: [0; 3] });
        self.mesh_ids.insert(key, id);
        Ok(id)
    }

    fn patch_mesh_id(&mut self, mesh: &Arc<BilinearPatchMesh>) -> Result<u32, GpuError> {
        let key = ptr_key(mesh);
        if let Some(&id) = self.mesh_ids.get(&key) {
            return Ok(id);
        }
        let first_index = u32_checked(self.indices.len(), "index count")?;
        let first_vertex = u32_checked(self.p.len() / 3, "vertex count")?;
        for &i in mesh.vertex_indices.iter() {
            self.indices.push(u32_checked(i, "vertex index")?); // FOUR per patch: p00, p10, p01, p11 (bilinear_patch.rs:87-106)
        }
        let mut flags = SG_MESH_BILINEAR;
        self.push_vertices(first_vertex as usize, mesh.n_vertices,
                           mesh.p.iter().map(|q| [q.x, q.y, q.z]),
                           if mesh.n.is_empty() { None } else { flags |= SG_MESH_HAS_N; Some(mesh.n.iter().map(|q| [q.x, q.y, q.z]).collect()) },
                           if mesh.uv.is_empty() { None } else { flags |= SG_MESH_HAS_UV; Some(mesh.uv.iter().map(|q| [q.x, q.y]).collect()) },
                           None);
        if mesh.reverse_orientation { flags |= SG_MESH_REVERSE_ORIENTATION; }
        if mesh.transform_swaps_handedness { flags |= SG_MESH_SWAPS_HANDEDNESS; }
        let id = self.meshes.len() as u32;
        self.meshes.push(SgMesh { first_index, first_vertex, n_triangles: u32_checked(mesh.n_patches, "patch count")?,
                                  n_vertices: u32_checked(mesh.n_vertices, "vertex count")?, flags, pad: [0; 3] });
        self.mesh_ids.insert(key, id);
        Ok(id)
    }

    /// Attribute arrays are scene-global and parallel to `p`: a mesh without normals / uvs / tangents still occupies its
    /// vertex range in them (zero-filled) as soon as ANY mesh has that attribute.
    fn push_vertices(&mut self, first_vertex: usize, n_vertices: usize, p: impl Iterator<Item = [f32; 3]>,
                     n: Option<Vec<[f32; 3]>>, uv: Option<Vec<[f32; 2]>>, s: Option<Vec<[f32; 3]>>) {
        for q in p {
            self.p.extend_from_slice(&q);
        }
        fn grow<const K: usize>(dst: &mut Vec<f32>, first_vertex: usize, n_vertices: usize, src: Option<Vec<[f32; K]>>) {
            match src {
                Some(v) => {
                    dst.resize(first_vertex * K, 0.0); // earlier meshes without this attribute
                    for q in v {
                        dst.extend_from_slice(&q);
                    }
                }
                None => {
                    if !dst.is_empty() {
                        dst.resize((first_vertex + n_vertices) * K, 0.0);
                    }
                }
            }
        }
        grow(&mut self.n, first_vertex, n_vertices, n);
        grow(&mut self.uv, first_vertex, n_vertices, uv);
        grow(&mut self.s, first_vertex, n_vertices, s);
    }

    // ----------------------------------------------------------------------------------------------------- spectra
    fn spectrum_id(&mut self, sp: &Arc<Spectrum>) -> Result<i32, GpuError> {
        let key = ptr_key(sp);
        if let Some(&id) = self.spectrum_ids.get(&key) {
            return Ok(id);
        }
        let id = self.push_spectrum(sp.as_ref())?;
        self.spectrum_ids.insert(key, id);
        Ok(id)
    }
    fn push_spectrum(&mut self, sp: &Spectrum) -> Result<i32, GpuError> {
        let row = match sp {
            Spectrum::Constant(c) => SgSpectrum { kind: SG_SPECTRUM_CONSTANT, n: 0, lambda_min: 0, c: c.c, scale: 1.0, off_a: 0, off_b: 0, pad: 0 },
            Spectrum::DenselySampled(d) => self.dense_row(d),
            Spectrum::PiecewiseLinear(pl) => {
                let off_a = self.pool.len() as u32;
                self.pool.extend_from_slice(&pl.lambdas);
                let off_b = self.pool.len() as u32;
                self.pool.extend_from_slice(&pl.values);
                SgSpectrum { kind: SG_SPECTRUM_PIECEWISE_LINEAR, n: pl.lambdas.len() as i32, lambda_min: 0, c: 0.0, scale: 1.0, off_a, off_b, pad: 0 }
            }
            Spectrum::Blackbody(b) => SgSpectrum { kind: SG_SPECTRUM_BLACKBODY, n: 0, lambda_min: 0, c: b.t, scale: b.normalization_factor, off_a: 0, off_b: 0, pad: 0 },
            // RGB spectra only arise from RGB *parameters* (paramdict.rs) -- sample them densely once, like
            // DenselySampledSpectrum::new does for every light (light.rs:417,556); exact at integer wavelengths, which is
            // where the reference's own dense copies are exact too
            other => {
                let d = DenselySampledSpectrum::new(other);
                self.dense_row(&d)
            }
        };
        self.spectra.push(row);
        Ok(self.spectra.len() as i32 - 1)
    }
    fn dense_row(&mut self, d: &DenselySampledSpectrum) -> SgSpectrum {
        let off_a = self.pool.len() as u32;
        self.pool.extend_from_slice(&d.values);
        SgSpectrum { kind: SG_SPECTRUM_DENSE, n: d.values.len() as i32, lambda_min: d.lambda_min, c: 0.0, scale: 1.0, off_a, off_b: 0, pad: 0 }
    }
    fn dense_id(&mut self, d: &DenselySampledSpectrum) -> i32 {
        let row = self.dense_row(d);
        self.spectra.push(row);
        self.spectra.len() as i32 - 1
    }
    fn const_spectrum(&mut self, c: f32) -> i32 {
        self.spectra.push(SgSpectrum { kind: SG_SPECTRUM_CONSTANT, n: 0, lambda_min: 0, c, scale: 1.0, off_a: 0, off_b: 0, pad: 0 });
        self.spectra.len() as i32 - 1
    }

    // ---------------------------------------------------------------------------------------------------- textures
    /// A float parameter: `Ok(Err(value))` for a `FloatConstantTexture` (stays in SgMaterial), `Ok(Ok(id))` for anything else.
    fn float_param(&mut self, t: &Arc<FloatTexture>) -> Result<Result<i32, f32>, GpuError> {
        if let FloatTexture::Constant(c) = t.as_ref() {
            return Ok(Err(c.value));
        }
        Ok(Ok(self.float_texture_id(t, 0)?))
    }
    /// A spectrum parameter: `Err(spectrum id)` for a `SpectrumConstantTexture`, `Ok(texture id)` otherwise.
    fn spectrum_param(&mut self, t: &Arc<SpectrumTexture>) -> Result<Result<i32, i32>, GpuError> {
        if let SpectrumTexture::Constant(c) = t.as_ref() {
            return Ok(Err(self.spectrum_id(&c.value)?));
        }
        Ok(Ok(self.spectrum_texture_id(t, 0)?))
    }

    fn float_texture_id(&mut self, t: &Arc<FloatTexture>, depth: i32) -> Result<i32, GpuError> {
        let key = ptr_key(t);
        if let Some(&id) = self.float_tex_ids.get(&key) {
            return Ok(id);
        }
        if depth > SG_MAX_TEXTURE_DEPTH {
            return unsupported("texture operands nested deeper than SG_MAX_TEXTURE_DEPTH");
        }
        let id = match t.as_ref() {
            FloatTexture::Image(img) => self.image_texture(&img.base, 1, SG_SPECTRUM_TYPE_ALBEDO)?,
            FloatTexture::Constant(c) => self.node_texture(1, SG_TEXTURE_CONSTANT, SgTextureNode { tex1: -1, tex2: -1, amount: -1, spectrum: -1, value: c.value, dir: [0.0; 3] }),
            FloatTexture::Scaled(s) => {
                let (a, b) = (self.float_texture_id(&s.tex, depth + 1)?, self.float_texture_id(&s.scale, depth + 1)?);
                self.node_texture(1, SG_TEXTURE_SCALED, SgTextureNode { tex1: a, tex2: b, amount: -1, spectrum: -1, value: 0.0, dir: [0.0; 3] })
            }
            FloatTexture::Mix(m) => {
                let (a, b, c) = (self.float_texture_id(&m.tex1, depth + 1)?, self.float_texture_id(&m.tex2, depth + 1)?, self.float_texture_id(&m.amount, depth + 1)?);
                self.node_texture(1, SG_TEXTURE_MIX, SgTextureNode { tex1: a, tex2: b, amount: c, spectrum: -1, value: 0.0, dir: [0.0; 3] })
            }
            FloatTexture::DirectionMix(m) => {
                let (a, b) = (self.float_texture_id(&m.tex1, depth + 1)?, self.float_texture_id(&m.tex2, depth + 1)?);
                self.node_texture(1, SG_TEXTURE_DIRECTION_MIX, SgTextureNode { tex1: a, tex2: b, amount: -1, spectrum: -1, value: 0.0, dir: [m.dir.x, m.dir.y, m.dir.z] })
            }
        };
        self.float_tex_ids.insert(key, id);
        Ok(id)
    }

    fn spectrum_texture_id(&mut self, t: &Arc<SpectrumTexture>, depth: i32) -> Result<i32, GpuError> {
        let key = ptr_key(t);
        if let Some(&id) = self.spectrum_tex_ids.get(&key) {
            return Ok(id);
        }
        if depth > SG_MAX_TEXTURE_DEPTH {
            return unsupported("texture operands nested deeper than SG_MAX_TEXTURE_DEPTH");
        }
        let id = match t.as_ref() {
            SpectrumTexture::Image(img) => {
                let st = match img.spectrum_type {
                    SpectrumType::Albedo => SG_SPECTRUM_TYPE_ALBEDO,
                    SpectrumType::Unbounded => SG_SPECTRUM_TYPE_UNBOUNDED,
                    SpectrumType::Illuminant => return unsupported("illuminant image textures (image area lights are todo!() in the reference too, light.rs:585-590)"),
                };
                let nc = img.base.mipmap.pyramid[0].n_channels() as i32;
                self.image_texture(&img.base, if nc == 1 { 1 } else { 3 }, st)?
            }
            SpectrumTexture::Constant(c) => {
                let sp = self.spectrum_id(&c.value)?;
                self.node_texture(3, SG_TEXTURE_CONSTANT, SgTextureNode { tex1: -1, tex2: -1, amount: -1, spectrum: sp, value: 0.0, dir: [0.0; 3] })
            }
            SpectrumTexture::Scaled(s) => {
                let (a, b) = (self.spectrum_texture_id(&s.tex, depth + 1)?, self.float_texture_id(&s.scale, depth + 1)?);
                self.node_texture(3, SG_TEXTURE_SCALED, SgTextureNode { tex1: a, tex2: b, amount: -1, spectrum: -1, value: 0.0, dir: [0.0; 3] })
            }
            SpectrumTexture::Mix(m) => {
                let (a, b, c) = (self.spectrum_texture_id(&m.tex1, depth + 1)?, self.spectrum_texture_id(&m.tex2, depth + 1)?, self.float_texture_id(&m.amount, depth + 1)?);
                self.node_texture(3, SG_TEXTURE_MIX, SgTextureNode { tex1: a, tex2: b, amount: c, spectrum: -1, value: 0.0, dir: [0.0; 3] })
            }
            SpectrumTexture::DirectionMix(m) => {
                let (a, b) = (self.spectrum_texture_id(&m.tex1, depth + 1)?, self.spectrum_texture_id(&m.tex2, depth + 1)?);
                self.node_texture(3, SG_TEXTURE_DIRECTION_MIX, SgTextureNode { tex1: a, tex2: b, amount: -1, spectrum: -1, value: 0.0, dir: [m.dir.x, m.dir.y, m.dir.z] })
            }
        };
        self.spectrum_tex_ids.insert(key, id);
        Ok(id)
    }

    fn node_texture(&mut self, n_channels: i32, kind: i32, node: SgTextureNode) -> i32 {
        self.texture_nodes.push(node);
        self.textures.push(SgTexture { n_channels, n_levels: 0, first_level: 0, wrap: 0, filter: 0, max_anisotropy: 0.0, scale: 1.0, invert: 0,
                                       su: 1.0, sv: 1.0, du: 0.0, dv: 0.0, spectrum_type: 0, mapping: -1, kind, node: self.texture_nodes.len() as i32 - 1 });
        self.textures.len() as i32 - 1
    }

    /// `ImageTextureBase` (texture.rs:19-26): mapping, scale, invert and the shared `Arc<MIPMap>` -- the pyramid built by
    /// `Image::generate_pyramid` (image.rs:699-787) goes up once per MIPMap as linear f32 texels (`Image::get_channel`
    /// decodes 8-bit / half storage through the colour encoding, image.rs:134-177).
    fn image_texture(&mut self, base: &ImageTextureBase, n_channels: i32, spectrum_type: i32) -> Result<i32, GpuError> {
        let mip: &Arc<MIPMap> = &base.mipmap;
        let key = ptr_key(mip);
        let (first_level, n_levels, nc) = match self.mip_ids.get(&key) {
            Some(&v) => v,
            None => {
                let first_level = self.image_levels.len() as u32;
                for level in mip.pyramid.iter() {
                    let res = level.resolution();
                    let offset = u32_checked(self.texels.len(), "texel pool")?;
                    for y in 0..res.y {
                        for x in 0..res.x {
                            for c in 0..n_channels {
                                self.texels.push(level.get_channel(crate::vecmath::Point2i::new(x, y), c as usize));
                            }
                        }
                    }
                    self.image_levels.push(SgImageLevel { offset, res: [res.x, res.y], pad: 0 });
                }
                let v = (first_level, mip.pyramid.len() as i32, n_channels);
                self.mip_ids.insert(key, v);
                if n_channels == 3 {
                    self.need_rgb2spec(mip.color_space.as_ref())?;
                }
                v
            }
        };
        if nc != n_channels {
            return unsupported("one MIPMap used with two channel counts");
        }
        let (mapping, su, sv, du, dv) = match &base.mapping {
            TextureMapping2D::UV(m) => (-1, m.su, m.sv, m.du, m.dv),
            TextureMapping2D::Spherical(m) => (self.push_mapping(SG_MAPPING_SPHERICAL, &m.texture_from_render, [0.0; 3], [0.0; 3], 0.0, 0.0), 1.0, 1.0, 0.0, 0.0),
            TextureMapping2D::Cylindrical(m) => (self.push_mapping(SG_MAPPING_CYLINDRICAL, &m.texture_from_render, [0.0; 3], [0.0; 3], 0.0, 0.0), 1.0, 1.0, 0.0, 0.0),
            TextureMapping2D::Planar(m) => (self.push_mapping(SG_MAPPING_PLANAR, &m.texture_from_render, [m.vs.x, m.vs.y, m.vs.z], [m.vt.x, m.vt.y, m.vt.z], m.ds, m.dt), 1.0, 1.0, 0.0, 0.0),
        };
        let filter = match mip.options.filter {
            FilterFunction::Point => SG_FILTER_POINT,
            FilterFunction::Bilinear => SG_FILTER_BILINEAR,
            FilterFunction::Trilinear => SG_FILTER_TRILINEAR,
            FilterFunction::EWA => SG_FILTER_EWA,
        };
        let wrap = match mip.wrap_mode {
            crate::image::WrapMode::Repeat => SG_WRAP_REPEAT,
            crate::image::WrapMode::Black => SG_WRAP_BLACK,
            crate::image::WrapMode::Clamp => SG_WRAP_CLAMP,
            crate::image::WrapMode::OctahedralSphere => return unsupported("octahedral-sphere wrap on a surface texture"),
        };
        self.textures.push(SgTexture { n_channels, n_levels, first_level, wrap, filter, max_anisotropy: mip.options.max_anisotropy.into_inner(),
                                       scale: base.scale, invert: base.invert as i32, su, sv, du, dv, spectrum_type, mapping, kind: SG_TEXTURE_IMAGE, node: -1 });
        Ok(self.textures.len() as i32 - 1)
    }

    fn push_mapping(&mut self, kind: i32, texture_from_render: &Transform, vs: [f32; 3], vt: [f32; 3], ds: f32, dt: f32) -> i32 {
        self.texture_mappings.push(SgTextureMapping { kind, texture_from_render: m16(&texture_from_render.m), vs, vt, ds, dt, pad: [0; 3] });
        self.texture_mappings.len() as i32 - 1
    }

    /// The colour space's RGB -> sigmoid-polynomial table (`rgb2spec` 0.1.1, rgb_to_spectra.rs:16-45): one table per scene --
    /// every three-channel texture must share the colour space.  The crate keeps `RGB2Spec`'s fields private, so the same
    /// `rgbtospec/<gamut>.spec` file the reference loads lazily is read here directly: "SPEC", u32 res, f32 scale[res],
    /// f32 data[3 * res^3 * 3] (little endian; the layout `rgb2spec_load` of the original C code reads).
    fn need_rgb2spec(&mut self, cs: Option<&Arc<crate::colorspace::RgbColorSpace>>) -> Result<(), GpuError> {
        if self.rgb2spec_res != 0 {
            return Ok(());
        }
        let cs = cs.ok_or_else(|| GpuError::Unsupported("RGB texture without a colour space".into()))?;
        let file = match cs.gamut {
            Gamut::SRGB => "rgbtospec/srgb.spec",
            Gamut::XYZ => "rgbtospec/xyz.spec",
            Gamut::ERGB => "rgbtospec/ergb.spec",
            Gamut::Aces2065_1 => "rgbtospec/aces2065_1.spec",
            Gamut::ProPhotoRGB => "rgbtospec/prophotorgb.spec",
            Gamut::Rec2020 => "rgbtospec/rec2020.spec",
        };
        let raw = std::fs::read(file).map_err(|e| GpuError::Unsupported(format!("{}: {}", file, e)))?;
        if raw.len() < 8 || &raw[0..4] != b"SPEC" {
            return unsupported(format!("{} is not an rgb2spec table", file));
        }
        let res = u32::from_le_bytes([raw[4], raw[5], raw[6], raw[7]]) as usize;
        let n_data = 3 * res * res * res * 3;
        if raw.len() != 8 + 4 * (res + n_data) {
            return unsupported(format!("{}: unexpected size for resolution {}", file, res));
        }
        let f = |i: usize| f32::from_le_bytes([raw[8 + 4 * i], raw[9 + 4 * i], raw[10 + 4 * i], raw[11 + 4 * i]]);
        self.rgb2spec_scale = (0..res).map(f).collect();
        self.rgb2spec_data = (res..res + n_data).map(f).collect();
        self.rgb2spec_res = res as u32;
        Ok(())
    }

    // --------------------------------------------------------------------------------------------------- materials
    fn material_id(&mut self, m: &Arc<Material>) -> Result<u32, GpuError> {
        let key = ptr_key(m);
        if let Some(&id) = self.material_ids.get(&key) {
            return Ok(id);
        }
        // reserve the row first: a Mix refers to its children by id and may (in a malformed scene) refer to itself
        let id = self.materials.len() as u32;
        self.material_ids.insert(key, id);
        self.materials.push(blank_material());
        self.material_textures.push(blank_material_textures());
        let (row, tex) = self.material_row(m.as_ref())?;
        self.materials[id as usize] = row;
        self.any_material_texture |= tex_row_used(&tex);
        self.material_textures[id as usize] = tex;
        Ok(id)
    }

    fn material_row(&mut self, m: &Material) -> Result<(SgMaterial, SgMaterialTextures), GpuError> {
        let mut r = blank_material();
        let mut t = blank_material_textures();
        // roughness-like float parameter -> SgMaterial field or SgMaterialTextures id
        macro_rules! float_into { ($tex:expr, $field:expr, $tid:expr) => {
            match self.float_param($tex)? { Err(v) => $field = v, Ok(id) => $tid = id }
        } }
        macro_rules! spectrum_into { ($tex:expr, $field:expr, $tid:expr) => {
            match self.spectrum_param($tex)? { Err(sp) => $field = sp, Ok(id) => { $field = self.const_spectrum(0.0); $tid = id } }
        } }
        match m {
            Material::Mix(mix) => {
                r.kind = SG_MATERIAL_MIX;
                r.mix_materials = [self.material_id(&mix.materials[0])? as i32, self.material_id(&mix.materials[1])? as i32];
                match self.float_param(&mix.amount)? { Err(v) => r.mix_amount = v, Ok(id) => r.tex_mix_amount = id }
                return Ok((r, t));
            }
            Material::Single(single) => {
                let (displacement, normal_map): (&Option<Arc<FloatTexture>>, &Option<Arc<Image>>) = match single {
                    SingleMaterial::Diffuse(d) => {
                        r.kind = SG_MATERIAL_DIFFUSE;
                        match self.spectrum_param(&d.reflectance)? { Err(sp) => r.spec_a = sp, Ok(id) => { r.spec_a = self.const_spectrum(0.0); r.tex_reflectance = id } }
                        (&d.displacement, &d.normal_map)
                    }
                    SingleMaterial::Conductor(c) => {
                        r.kind = SG_MATERIAL_CONDUCTOR;
                        if c.remap_roughness { r.flags |= SG_MAT_REMAP_ROUGHNESS; }
                        float_into!(&c.u_roughness, r.u_roughness, t.u_roughness);
                        float_into!(&c.v_roughness, r.v_roughness, t.v_roughness);
                        match (&c.eta, &c.k) {
                            (Some(eta), Some(k)) => { spectrum_into!(eta, r.spec_a, t.spec_a); spectrum_into!(k, r.spec_b, t.spec_b); }
                            // `reflectance` form AS WRITTEN (material.rs:478-495): r is clamped to [0.0, 0.0000], i.e. to 0, so
                            // eta = 1, k = 2 sqrt(0) / sqrt(1 - 0) = 0 whatever the texture says
                            _ => { r.spec_a = self.const_spectrum(1.0); r.spec_b = self.const_spectrum(0.0); }
                        }
                        (&c.displacement, &c.normal_map)
                    }
                    SingleMaterial::Dielectric(d) => {
                        r.kind = SG_MATERIAL_DIELECTRIC;
                        if d.remap_roughness { r.flags |= SG_MAT_REMAP_ROUGHNESS; }
                        float_into!(&d.u_roughness, r.u_roughness, t.u_roughness);
                        float_into!(&d.v_roughness, r.v_roughness, t.v_roughness);
                        r.spec_a = self.spectrum_id(&d.eta)?;     // a `Spectrum`, not a texture: dispersion is decided by its kind (material.rs:609-620)
                        (&d.displacement, &d.normal_map)
                    }
                    SingleMaterial::ThinDielectric(d) => {
                        r.kind = SG_MATERIAL_THIN_DIELECTRIC;
                        r.spec_a = self.spectrum_id(&d.eta)?;
                        (&d.displacement, &d.normal_map)
                    }
                    SingleMaterial::CoatedDiffuse(c) => {
                        r.kind = SG_MATERIAL_COATED_DIFFUSE;
                        if c.remap_roughness { r.flags |= SG_MAT_REMAP_ROUGHNESS; }
                        match self.spectrum_param(&c.reflectance)? { Err(sp) => r.spec_a = sp, Ok(id) => { r.spec_a = self.const_spectrum(0.0); r.tex_reflectance = id } }
                        spectrum_into!(&c.albedo, r.spec_b, t.spec_b);
                        float_into!(&c.u_roughness, r.u_roughness, t.u_roughness);
                        float_into!(&c.v_roughness, r.v_roughness, t.v_roughness);
                        float_into!(&c.thickness, r.thickness, t.thickness);
                        float_into!(&c.g, r.g, t.g);
                        r.spec_c = self.spectrum_id(&c.eta)?;
                        r.max_depth = c.max_depth; r.n_samples = c.n_samples;
                        (&c.displacement, &c.normal_map)
                    }
                    SingleMaterial::CoatedConductor(c) => {
                        r.kind = SG_MATERIAL_COATED_CONDUCTOR;
                        if c.remap_roughness { r.flags |= SG_MAT_REMAP_ROUGHNESS; }
                        float_into!(&c.interface_u_roughness, r.u_roughness, t.u_roughness);
                        float_into!(&c.interface_v_roughness, r.v_roughness, t.v_roughness);
                        float_into!(&c.thickness, r.thickness, t.thickness);
                        float_into!(&c.g, r.g, t.g);
                        r.spec_c = self.spectrum_id(&c.interface_eta)?;
                        spectrum_into!(&c.albedo, r.spec_b, t.spec_b);
                        float_into!(&c.conductor_u_roughness, r.u_roughness2, t.u_roughness2);
                        float_into!(&c.conductor_v_roughness, r.v_roughness2, t.v_roughness2);
                        match (&c.conductor_eta, &c.k, &c.reflectance) {
                            (Some(eta), Some(k), _) => { spectrum_into!(eta, r.spec_a, t.spec_a); spectrum_into!(k, r.spec_d, t.spec_d); }
                            (_, _, Some(refl)) => { r.flags |= SG_MAT_CONDUCTOR_REFLECTANCE; spectrum_into!(refl, r.spec_a, t.spec_a); }   // material.rs:1225-1233
                            _ => return unsupported("coated conductor without eta/k or reflectance"),
                        }
                        r.max_depth = c.max_depth; r.n_samples = c.n_samples;
                        (&c.displacement, &c.normal_map)
                    }
                };
                if let Some(d) = displacement {
                    r.flags |= SG_MAT_HAS_DISPLACEMENT;          // DiffuseMaterial always stores Some (material.rs:280)
                    match self.float_param(d)? { Err(v) => r.displacement = v, Ok(id) => r.tex_displacement = id }
                }
                if let Some(img) = normal_map {
                    r.normal_map = self.normal_map_texture(img)?;
                }
            }
        }
        Ok((r, t))
    }

    /// A normal map is read with `Image::bilerp_channel_wrapped(.., WrapMode::Repeat)` on level 0 only
    /// (material.rs:1453-1474): one three-channel, one-level... the ABI wants a full pyramid ending in 1x1, so the levels are
    /// generated on the device (`sg_image_generate_pyramid`) from the decoded image.
    fn normal_map_texture(&mut self, img: &Arc<Image>) -> Result<i32, GpuError> {
        let res = img.resolution();
        let mut lin = Vec::with_capacity((res.x * res.y * 3) as usize);
        for y in 0..res.y {
            for x in 0..res.x {
                for c in 0..3usize {
                    lin.push(img.get_channel(crate::vecmath::Point2i::new(x, y), c));
                }
            }
        }
        let mut n_levels = 0i32;
        let mut n_texels = 0u64;
        let mut levels = [SgImageLevel { offset: 0, res: [0, 0], pad: 0 }; 32];
        super::check(unsafe { sg_image_pyramid_layout(res.x, res.y, 3, &mut n_levels, levels.as_mut_ptr(), &mut n_texels) })?;
        let base = u32_checked(self.texels.len(), "texel pool")?;
        self.texels.resize(self.texels.len() + n_texels as usize, 0.0);
        super::check(unsafe { sg_image_generate_pyramid(lin.as_ptr(), res.x, res.y, 3, SG_WRAP_REPEAT, self.texels[base as usize..].as_mut_ptr()) })?;
        let first_level = self.image_levels.len() as u32;
        for l in levels[..n_levels as usize].iter() {
            self.image_levels.push(SgImageLevel { offset: base + l.offset, res: l.res, pad: 0 });
        }
        self.textures.push(SgTexture { n_channels: 3, n_levels, first_level, wrap: SG_WRAP_REPEAT, filter: SG_FILTER_BILINEAR, max_anisotropy: 8.0, scale: 1.0,
                                       invert: 0, su: 1.0, sv: 1.0, du: 0.0, dv: 0.0, spectrum_type: SG_SPECTRUM_TYPE_ALBEDO, mapping: -1, kind: SG_TEXTURE_IMAGE, node: -1 });
        Ok(self.textures.len() as i32 - 1)
    }

    // ------------------------------------------------------------------------------------------------------ lights
    /// Light rows in light-sampler order.  Area lights get their (mesh, tri) when `push_shape` meets their primitive.
    fn flatten_lights(&mut self, lights: &Arc<Vec<Arc<Light>>>) -> Result<(), GpuError> {
        for (i, l) in lights.iter().enumerate() {
            self.light_ids.insert(ptr_key(l), i as i32);
            let mut row = SgLight { kind: 0, spectrum: 0, scale: 0.0, two_sided: 0, mesh: 0, tri: 0, area: 0.0, pos: [0.0; 3],
                                    scene_center: [0.0; 3], scene_radius: 0.0, pad: [0.0; 2] };
            match l.as_ref() {
                Light::DiffuseAreaLight(a) => {
                    row.kind = SG_LIGHT_DIFFUSE_AREA;                 // refined by the shape kind in push_shape
                    row.spectrum = self.dense_id(&a.l_emit);
                    row.scale = a.scale;                              // already divided by spectrum_to_photometric (light.rs:583)
                    row.two_sided = a.two_sided as i32;
                    row.area = a.area;                                // Shape::area() cached at construction (light.rs:546)
                    self.area_lights_pending.insert(i);
                }
                Light::Point(p) => {
                    row.kind = SG_LIGHT_POINT;
                    row.spectrum = self.dense_id(&p.i);
                    row.scale = p.scale;
                    let o = p.base.render_from_light.m.m;             // render_from_light applied to the origin (light.rs:470-476)
                    row.pos = [o[0][3], o[1][3], o[2][3]];
                }
                Light::UniformInfinite(u) => {
                    row.kind = SG_LIGHT_UNIFORM_INFINITE;
                    row.spectrum = self.dense_id(&u.l_emit);
                    row.scale = u.scale;
                    row.scene_center = [u.scene_center.x, u.scene_center.y, u.scene_center.z];   // preprocess() ran in create_lights
                    row.scene_radius = u.scene_radius;
                }
                Light::ImageInfinite(im) => {
                    row.kind = SG_LIGHT_IMAGE_INFINITE;
                    row.scale = im.scale;
                    row.scene_center = [im.scene_center.x, im.scene_center.y, im.scene_center.z];
                    row.scene_radius = im.scene_radius;
                    row.spectrum = self.spectrum_id(&im.image_color_space.illuminant)?;
                    row.tri = self.push_env_map(im)?;
                }
            }
            self.lights.push(row);
        }
        Ok(())
    }

    /// `ImageInfinitelight` (light.rs:805-981): the equal-area square image as linear RGB texels plus its two
    /// `PiecewiseConstant2D`s exactly as `PiecewiseConstant2D::new` built them (sampling.rs:101-179).
    fn push_env_map(&mut self, im: &crate::light::ImageInfinitelight) -> Result<u32, GpuError> {
        let res = im.image.resolution();
        if res.x != res.y {
            return unsupported("environment map is not square (the reference asserts the same, light.rs:826-833)");
        }
        self.need_rgb2spec(Some(&im.image_color_space))?;
        let texel_offset = self.texels.len() as u64;
        for y in 0..res.y {
            for x in 0..res.x {
                for c in 0..3usize {
                    self.texels.push(im.image.get_channel(crate::vecmath::Point2i::new(x, y), c));
                }
            }
        }
        let distribution = self.push_distribution(&im.distribution);
        let compensated = self.push_distribution(&im.compensated_distribution);
        self.env_maps.push(SgEnvMap { render_from_light: m16(&im.base.render_from_light.m), light_from_render: m16(&im.base.render_from_light.m_inv),
                                      texel_offset, res: res.x, pad: 0, distribution, compensated });
        Ok(self.env_maps.len() as u32 - 1)
    }
    fn push_distribution(&mut self, d: &crate::sampling::PiecewiseConstant2D) -> SgDistribution2D {
        let nv = d.conditional_v.len();
        let nu = d.conditional_v[0].func.len();
        let func_off = self.pool.len() as u32;
        for row in d.conditional_v.iter() { self.pool.extend_from_slice(&row.func); }
        let cdf_off = self.pool.len() as u32;
        for row in d.conditional_v.iter() { self.pool.extend_from_slice(&row.cdf); }
        let marg_func_off = self.pool.len() as u32;
        self.pool.extend_from_slice(&d.marginal.func);
        let marg_cdf_off = self.pool.len() as u32;
        self.pool.extend_from_slice(&d.marginal.cdf);
        SgDistribution2D { nu: nu as i32, nv: nv as i32, func_off, cdf_off, marg_func_off, marg_cdf_off, marg_integral: d.marginal.func_int, pad: 0 }
    }

    // ------------------------------------------------------------------------------------------- camera, film, sensor
    fn flatten_camera(&mut self, camera: &Camera) -> Result<(), GpuError> {
        let (pb, dx, dy, kind) = match camera {
            Camera::Perspective(c) => (&c.projective_base, c.dx_camera, c.dy_camera, SG_CAMERA_PERSPECTIVE),
            Camera::Orthographic(c) => (&c.projective_base, c.dx_camera, c.dy_camera, SG_CAMERA_ORTHOGRAPHIC),
        };
        let cb = &pb.camera_base;
        let v3 = |v: crate::vecmath::Vector3f| [v.x, v.y, v.z];
        self.camera = Some(SgCamera {
            camera_from_raster: m16(&pb.camera_from_raster.m),
            render_from_camera: m16(&cb.camera_transform.render_from_camera.m),
            camera_from_render: m16(&cb.camera_transform.render_from_camera.m_inv),
            dx_camera: v3(dx), dy_camera: v3(dy),
            lens_radius: pb.lens_radius, focal_distance: pb.focal_distance,
            shutter_open: cb.shutter_open, shutter_close: cb.shutter_close,
            min_pos_differential_x: v3(cb.min_pos_differential_x), min_pos_differential_y: v3(cb.min_pos_differential_y),
            min_dir_differential_x: v3(cb.min_dir_differential_x), min_dir_differential_y: v3(cb.min_dir_differential_y),
            kind, pad: 0,
        });
        Ok(())
    }

    fn flatten_film(&mut self, film: &Arc<Film>) -> Result<(), GpuError> {
        let Film::RgbFilm(rgb) = film.as_ref();
        let full = film.full_resolution();
        let pbnd = film.pixel_bounds();
        let radius = film.get_filter().radius();       // BoxFilter is the only filter (filter.rs:20-24); its weight is 1 (filter.rs:99-105)
        let sensor = film.get_pixel_sensor();
        let (r_bar, g_bar, b_bar) = (self.dense_id(&sensor.r_bar), self.dense_id(&sensor.g_bar), self.dense_id(&sensor.b_bar));
        let mut out = [0.0f32; 9];
        for r in 0..3 { for c in 0..3 { out[3 * r + c] = rgb.output_rgb_from_sensor_rgb.m[r][c]; } }
        self.film = Some(SgFilm {
            full_resolution: [full.x, full.y],
            pixel_bounds: [pbnd.min.x, pbnd.min.y, pbnd.max.x, pbnd.max.y],
            filter_radius: [radius.x, radius.y],
            r_bar, g_bar, b_bar,
            imaging_ratio: sensor.imaging_ratio,
            max_component_value: rgb.max_component_value,
            output_rgb_from_sensor_rgb: out,
        });
        Ok(())
    }
}

fn blank_material() -> SgMaterial {
    SgMaterial { kind: 0, spec_a: 0, spec_b: 0, flags: 0, u_roughness: 0.0, v_roughness: 0.0, displacement: 0.0, spec_c: 0, thickness: 0.01, g: 0.0,
                 max_depth: 10, n_samples: 1, tex_reflectance: -1, tex_displacement: -1, pad2: [0; 2], spec_d: 0, u_roughness2: 0.0, v_roughness2: 0.0,
                 normal_map: -1, mix_materials: [-1, -1], mix_amount: 0.5, tex_mix_amount: -1 }
}
fn blank_material_textures() -> SgMaterialTextures {
    SgMaterialTextures { u_roughness: -1, v_roughness: -1, spec_a: -1, spec_b: -1, spec_d: -1, thickness: -1, g: -1, u_roughness2: -1, v_roughness2: -1, pad: [0; 3] }
}
fn tex_row_used(t: &SgMaterialTextures) -> bool {
    [t.u_roughness, t.v_roughness, t.spec_a, t.spec_b, t.spec_d, t.thickness, t.g, t.u_roughness2, t.v_roughness2].iter().any(|&x| x >= 0)
}
