/*
 * shimmer_gpu.h -- C ABI of the B200 wavefront path-tracing backend for shimmer.
 *
 * The reference (jalberse/shimmer) has NO FFI today: no `extern`, no build.rs,
 * no -sys crate (Cargo.toml:14-39).  This header therefore *defines* the
 * boundary that a `render::render_gpu` sibling of `render_cpu`
 * (src/render.rs:8-55) binds.  Every entry point below names the reference
 * interface it replaces (file:line relative to /root/reference).
 *
 * Conventions
 *   - all functions return 0 on success, a negative SgStatus otherwise; they
 *     never unwind, abort or call exit() (the reference panics instead:
 *     integrator.rs:36, aggregate.rs:27 -- the Rust side turns !=0 into panic!)
 *   - sg_last_error() returns a thread-local, NUL-terminated description
 *   - the host owns every input array; the library only borrows it for the
 *     duration of sg_scene_create (it is copied to HBM once)
 *   - all floats are IEEE f32 (`Float = f32`, src/float.rs:1-4); the film is f64
 *     (`RgbFilmPixel`, src/film.rs:470-479)
 *   - multi-GPU lives behind this boundary in two forms:
 *       single process, n GPUs : sg_init_multi(devices, n) -- every scene is replicated on
 *         every device, sg_render splits the sample range across them and sums the films onto
 *         devices[0] with ONE in-library ncclReduce (communicator from ncclCommInitAll);
 *       one process per GPU    : sg_init(device) + sg_comm_init_rank(id, rank, n) -- the same
 *         split / reduce over an ncclCommInitRank communicator (torchrun, MPI-style launches).
 *     NCCL is loaded lazily (dlopen libnccl.so.2) by the first sg_init_multi / sg_comm_* call,
 *     so single-GPU hosts do not need it.
 */
#ifndef SHIMMER_GPU_H
#define SHIMMER_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SG_ABI_VERSION 9

typedef enum SgStatus {
    SG_OK = 0,
    SG_ERR_INVALID_ARGUMENT = -1,
    SG_ERR_CUDA = -2,
    SG_ERR_OUT_OF_MEMORY = -3,
    SG_ERR_UNSUPPORTED = -4,
    SG_ERR_NOT_INITIALIZED = -5,
    SG_ERR_NCCL = -6
} SgStatus;

/* ---- acceleration structure ------------------------------------------------
 * Flattened `LinearBvhNode` (src/aggregate.rs:471-481): depth-first order, the
 * first child of an interior node is `index+1`, the second is `offset`.
 * `usize` fields are narrowed to u32 (range-checked by the host shim), which
 * packs the reference's 64-byte node into 32 bytes. */
typedef struct SgBvhNode {
    float    bmin[3];
    float    bmax[3];
    uint32_t offset;   /* leaf: primitive_offset; interior: second_child_offset */
    uint16_t n_prims;  /* >0 => leaf */
    uint8_t  axis;     /* interior: split axis 0/1/2 */
    uint8_t  pad;
} SgBvhNode;

/* One entry per `Primitive::{Simple,Geometric}` in BVH leaf order
 * (`ordered_primitives`, src/aggregate.rs:236-266; src/primitive.rs:30-35). */
#define SG_PRIM_INSTANCE 0xffffffffu   /* SgPrimitive.mesh of a `Primitive::Transformed`: tri = index into instances */
#define SG_PRIM_SPHERE   0xfffffffeu   /* SgPrimitive.mesh of a `Shape::Sphere` (shape/sphere.rs): tri = index into spheres */
typedef struct SgPrimitive {
    uint32_t mesh;      /* index into SgSceneDesc.meshes, or SG_PRIM_INSTANCE   */
    uint32_t tri;       /* Triangle::tri_index within that mesh (triangle.rs:49), or the instance index */
    uint32_t material;  /* index into materials                                  */
    int32_t  light;     /* area light index into lights, or -1 (SimplePrimitive) */
} SgPrimitive;

/* ---- object instancing (src/primitive.rs:136-176 TransformedPrimitive, src/loading/scene.rs:814-866) ----
 * An object definition is a group of primitives with its own BvhAggregate (scene.rs:821-830).  Its nodes live in
 * SgSceneDesc.nodes at [first_node, first_node + n_nodes) with second-child offsets RELATIVE to first_node and leaf
 * primitive offsets RELATIVE to first_prim; its primitives are SgSceneDesc.primitives[first_prim, first_prim + n_prims)
 * in that BVH's leaf order.  n_nodes == 0: the definition is one bare primitive (no aggregate).  The top-level
 * BvhAggregate is nodes[0, n_top_nodes) over primitives[0, n_top_primitives); only it may hold instance primitives.
 * A definition may hold triangles, bilinear patches and spheres (their transforms / vertices are relative to the
 * definition's space); emissive shapes inside definitions are rejected (pbrt-v4 scene format). */
typedef struct SgObject {
    uint32_t first_node, n_nodes, first_prim, n_prims;
} SgObject;
typedef struct SgInstance {
    float    render_from_primitive[16];   /* Transform::m, row-major     */
    float    primitive_from_render[16];   /* Transform::m_inv            */
    uint32_t object;                      /* index into objects          */
    uint32_t pad[3];
} SgInstance;
/* The reference transforms shadow rays with the FORWARD instance transform (primitive.rs:172-175; closest-hit uses
 * the inverse, :159-163) and maps interaction vectors/normals through the INVERSE transform (transform.rs:573-609).
 * Default: reproduce exactly that.  SG_SCENE_FIX_INSTANCING selects the pbrt semantics instead. */
enum { SG_SCENE_FIX_INSTANCING = 1 };

/* `Sphere` (src/shape/sphere.rs:26-45) as `Sphere::new` stores it (:69-92): z_min/z_max clamped to [-radius, radius],
 * theta_z_min/max = acos(clamp(z/radius)), phi_max in radians; flags: SG_MESH_REVERSE_ORIENTATION | SG_MESH_SWAPS_HANDEDNESS.
 * The shape works in object space (interval arithmetic on the transformed ray, sphere.rs:95-186) and maps the hit back
 * with Transform::apply(SurfaceInteraction) -- the same routine instancing uses, so SG_SCENE_FIX_INSTANCING also selects
 * the pbrt semantics (vectors through M, normals through M^-T) for spheres.  Transforms must be affine.
 * An emissive sphere's SgPrimitive.light points at an SG_LIGHT_DIFFUSE_AREA_SPHERE light whose `tri` is the sphere index. */
typedef struct SgSphere {
    float    render_from_object[16];
    float    object_from_render[16];
    float    radius, z_min, z_max, theta_z_min, theta_z_max, phi_max;
    uint32_t flags;
    uint32_t pad;
} SgSphere;

/* `TriangleMesh` (src/shape/mesh.rs:9-20); vertices already in render space
 * (mesh.rs:43-46).  Attribute arrays are scene-global; a mesh addresses
 * [first_vertex, first_vertex+n_vertices) and indices
 * [first_index, first_index+3*n_triangles) hold MESH-LOCAL vertex numbers. */
enum {
    SG_MESH_HAS_N = 1, SG_MESH_HAS_UV = 2, SG_MESH_HAS_S = 4,
    SG_MESH_REVERSE_ORIENTATION = 8, SG_MESH_SWAPS_HANDEDNESS = 16,
    SG_MESH_BILINEAR = 32   /* `BilinearPatchMesh` (mesh.rs:98-175): FOUR indices per patch (p00, p10, p01, p11; bilinear_patch.rs:87-106),
                               n_triangles = number of patches, SgPrimitive.tri = patch index.  Patches may sit inside object definitions; an emissive (top-level) patch's
                               SgPrimitive.light points at an SG_LIGHT_DIFFUSE_AREA_PATCH light over that patch. */
};
typedef struct SgMesh {
    uint32_t first_index;
    uint32_t first_vertex;
    uint32_t n_triangles;
    uint32_t n_vertices;
    uint32_t flags;
    uint32_t pad[3];
} SgMesh;

/* ---- spectra ---------------------------------------------------------------
 * `Spectrum` enum (src/spectra/spectrum.rs:39-48).  Data lives in one float
 * pool. */
typedef enum SgSpectrumKind {
    SG_SPECTRUM_CONSTANT = 0,        /* ConstantSpectrum, spectrum.rs:146-168: c        */
    SG_SPECTRUM_DENSE = 1,           /* DenselySampledSpectrum, :171-292: pool[off_a..+n], lambda_min */
    SG_SPECTRUM_PIECEWISE_LINEAR = 2,/* PiecewiseLinearSpectrum, :295-440: lambdas at off_a, values at off_b, n; both runs must lie
                                      * inside spectrum_pool (sg_scene_create rejects the scene otherwise); `pad` is ignored */
    SG_SPECTRUM_BLACKBODY = 3        /* BlackbodySpectrum, :443-495: c = T, scale = normalization_factor */
} SgSpectrumKind;
typedef struct SgSpectrum {
    int32_t  kind;
    int32_t  n;
    int32_t  lambda_min;
    float    c;
    float    scale;
    uint32_t off_a;
    uint32_t off_b;
    uint32_t pad;
} SgSpectrum;

/* ---- materials (src/material.rs) -------------------------------------------
 * Parameters held directly in SgMaterial are the reference's `*ConstantTexture`s (texture.rs), i.e. a spectrum id or a
 * float; any other texture goes through the texture table (tex_* ids below, SgMaterialTextures for the remaining parameters). */
typedef enum SgMaterialKind {
    SG_MATERIAL_DIFFUSE = 0,    /* DiffuseMaterial    material.rs:298-338  spec_a = reflectance          */
    SG_MATERIAL_CONDUCTOR = 1,  /* ConductorMaterial  material.rs:453-526  spec_a = eta, spec_b = k      */
    SG_MATERIAL_DIELECTRIC = 2, /* DielectricMaterial material.rs:600-662  spec_a = eta (`Spectrum`)     */
    SG_MATERIAL_COATED_DIFFUSE = 3, /* CoatedDiffuseMaterial material.rs:913-992: spec_a = reflectance, spec_b = albedo,
                                      spec_c = eta, thickness, g, max_depth, n_samples (LayeredBxDF bxdf.rs:883-1620) */
    SG_MATERIAL_THIN_DIELECTRIC = 4, /* ThinDielectricMaterial material.rs:666-760 / ThinDielectricBxDF bxdf.rs:797-880: spec_a = eta */
    SG_MATERIAL_COATED_CONDUCTOR = 5, /* CoatedConductorMaterial material.rs:995-1283 / CoatedConductorBxDF bxdf.rs:460-516 (LayeredBxDF<Dielectric,
                                        Conductor>): interface = u/v_roughness, spec_c (eta), thickness, g, spec_b (albedo), max_depth, n_samples;
                                        conductor = spec_a (eta) + spec_d (k), or SG_MAT_CONDUCTOR_REFLECTANCE: spec_a = reflectance;
                                        u/v_roughness2 = conductor roughness (IGNORED when remaproughness is set: material.rs:1237-1241
                                        derives the conductor alpha from the INTERFACE roughness -- reproduced) */
    SG_MATERIAL_MIX = 6             /* MixMaterial material.rs:1286-1330: mix_materials[2], amount = mix_amount or tex_mix_amount.  Resolved per
                                       hit before shading (interaction.rs:206-221).  The reference draws its random number from a per-thread
                                       SmallRng::from_entropy() (integrator.rs:255); here it comes from a generator seeded by the path stream's
                                       state (same rule as the LayeredBxDF's), so renders stay reproducible. */
} SgMaterialKind;
enum {
    SG_MAT_REMAP_ROUGHNESS = 1,  /* `remaproughness`, default true                      */
    SG_MAT_HAS_DISPLACEMENT = 2, /* material stores Some(displacement) -> bump_map runs
                                    (always true for Diffuse: material.rs:280)           */
    SG_MAT_CONDUCTOR_REFLECTANCE = 4 /* coated conductor given by `reflectance` (material.rs:1225-1233): spec_a = reflectance, no k */
};
typedef struct SgMaterial {
    int32_t kind;
    int32_t spec_a;
    int32_t spec_b;
    int32_t flags;
    float   u_roughness;
    float   v_roughness;
    float   displacement;  /* constant displacement texture value (0 by default) */
    int32_t spec_c;        /* coated diffuse: eta spectrum                        */
    float   thickness;     /* coated diffuse: `thickness` (0.01)                  */
    float   g;             /* coated diffuse: HG asymmetry (0)                    */
    int32_t max_depth;     /* coated diffuse: `maxdepth` (10)                     */
    int32_t n_samples;     /* coated diffuse: `nsamples` (1)                      */
    int32_t tex_reflectance;  /* SpectrumTexture id for `reflectance` (diffuse / coated diffuse) or -1 -> spec_a */
    int32_t tex_displacement; /* FloatTexture id for `displacement` (bump_map, material.rs:1477-1509) or -1     */
    int32_t pad2[2];
    int32_t spec_d;           /* coated conductor: conductor k                                                              */
    float   u_roughness2;     /* coated conductor: `conductor.uroughness` / `conductor.vroughness`                          */
    float   v_roughness2;
    int32_t normal_map;       /* three-channel image texture id (level 0 is read with Image::bilerp_channel_wrapped, repeat) or -1;
                                 only consulted when the material has NO displacement (interaction.rs:229-244, material.rs:1453-1474) */
    int32_t mix_materials[2]; /* SG_MATERIAL_MIX: the two materials (may be mixes themselves; cycles are rejected)          */
    float   mix_amount;       /* `amount` (0.5) when tex_mix_amount < 0                                                     */
    int32_t tex_mix_amount;   /* FloatTexture id for `amount` or -1                                                            */
} SgMaterial;

/* Texture-valued material parameters.  Every parameter the reference's materials read through `tex_eval.evaluate_float` /
 * `evaluate_spectrum` (material.rs:456-499, 603-635, 917-963, 1188-1260) may be any texture (image, constant, scaled, mix,
 * direction mix); the common case -- a `*ConstantTexture` -- stays in SgMaterial.  Optional table, one row per material
 * (SgSceneDesc.material_textures, NULL = none): a texture id replaces the SgMaterial field of the same name, -1 keeps it.
 * Float parameters need float textures (n_channels == 1).  `reflectance` of Diffuse / CoatedDiffuse, `displacement` and the mix
 * `amount` keep their ids in SgMaterial (tex_reflectance, tex_displacement, tex_mix_amount). */
typedef struct SgMaterialTextures {
    int32_t u_roughness, v_roughness;   /* `uroughness` / `vroughness` (`roughness`): conductor, dielectric, coated interface        */
    int32_t spec_a;                     /* conductor `eta`; coated conductor `conductor.eta` (or `reflectance` with SG_MAT_CONDUCTOR_REFLECTANCE) */
    int32_t spec_b;                     /* conductor `k`; coated diffuse / coated conductor `albedo`                                  */
    int32_t spec_d;                     /* coated conductor `conductor.k`                                                             */
    int32_t thickness, g;               /* coated diffuse / coated conductor `thickness`, `g`                                         */
    int32_t u_roughness2, v_roughness2; /* coated conductor `conductor.uroughness` / `conductor.vroughness`                           */
    int32_t pad[3];
} SgMaterialTextures;

/* ---- image textures (src/texture.rs:393-404,700-808,896-936, src/mipmap.rs:121-331, src/image.rs:134-177,619-646) ------
 * The host owns image decoding and pyramid generation (image.rs:699-846); it passes every MIP
 * level as linear f32 texels (what `Image::get_channel` returns after colour-decoding), channels
 * interleaved.  Level 0 is the full-resolution image, the last level is 1x1.  One-channel images
 * evaluate to a constant spectrum (texture.rs:803-807); three-channel images go through
 * RgbAlbedoSpectrum / RgbUnboundedSpectrum (spectrum.rs:498-588), i.e. the rgb2spec coefficient
 * table of the texture's colour space (SgSceneDesc.rgb2spec_*) and RgbSigmoidPolynomial::get
 * (color.rs:352-383).  Texture mapping is UVMapping (texture.rs:896-936). */
typedef enum SgWrapMode { SG_WRAP_REPEAT = 0, SG_WRAP_BLACK = 1, SG_WRAP_CLAMP = 2 } SgWrapMode;       /* image.rs WrapMode */
typedef enum SgFilterFunction { SG_FILTER_POINT = 0, SG_FILTER_BILINEAR = 1, SG_FILTER_TRILINEAR = 2, SG_FILTER_EWA = 3 } SgFilterFunction; /* mipmap.rs:337-345 */
typedef enum SgSpectrumType { SG_SPECTRUM_TYPE_ALBEDO = 0, SG_SPECTRUM_TYPE_UNBOUNDED = 1 } SgSpectrumType; /* texture.rs SpectrumType; Illuminant is not on this path */
typedef struct SgImageLevel {
    uint32_t offset;           /* into SgSceneDesc.texels (in floats)    */
    int32_t  res[2];           /* resolution x, y                        */
    uint32_t pad;
} SgImageLevel;
typedef struct SgTexture {
    int32_t  n_channels;       /* 1 or 3                                 */
    int32_t  n_levels;
    uint32_t first_level;      /* into SgSceneDesc.image_levels          */
    int32_t  wrap;             /* SgWrapMode, `wrap` default repeat      */
    int32_t  filter;           /* SgFilterFunction, default bilinear (texture.rs:741) */
    float    max_anisotropy;   /* default 8                              */
    float    scale;            /* default 1                              */
    int32_t  invert;
    float    su, sv, du, dv;   /* UVMapping (texture.rs:896-936)         */
    int32_t  spectrum_type;    /* SgSpectrumType (three-channel spectrum textures) */
    int32_t  mapping;          /* index into SgSceneDesc.texture_mappings, or -1 = UVMapping with su, sv, du, dv above */
    int32_t  kind;             /* SgTextureKind; 0 = image texture (every field above), else only n_channels and `node` are read */
    int32_t  node;             /* index into SgSceneDesc.texture_nodes for the non-image kinds                           */
} SgTexture;
/* The non-image members of `enum FloatTexture` / `enum SpectrumTexture` (texture.rs:88-94,411-417).  A texture row with
 * n_channels == 1 is a FloatTexture, any other channel count a SpectrumTexture; operands are texture ids of the required
 * type (spectrum operands may also be one-channel rows: they evaluate to a constant spectrum like a one-channel image,
 * texture.rs:803-807).  Nesting depth (operand of an operand ...) is at most SG_MAX_TEXTURE_DEPTH below the root. */
typedef enum SgTextureKind {
    SG_TEXTURE_IMAGE = 0,          /* FloatImageTexture :393-404 / SpectrumImageTexture :777-808                         */
    SG_TEXTURE_CONSTANT = 1,       /* FloatConstantTexture :138-177 (`value`) / SpectrumConstantTexture :485-535 (`spectrum`) */
    SG_TEXTURE_SCALED = 2,         /* Float/SpectrumScaledTexture :180-213,:537-583: scale == 0 ? 0 : tex * scale        */
    SG_TEXTURE_MIX = 3,            /* Float/SpectrumMixTexture :215-262,:585-651: tex1 skipped when amount == 1, tex2 when amount == 0 */
    SG_TEXTURE_DIRECTION_MIX = 4   /* Float/SpectrumDirectionMixTexture :264-310,:653-...: amount = dot(ctx.n, dir), NOT normalised or clamped */
} SgTextureKind;
#define SG_MAX_TEXTURE_DEPTH 3
typedef struct SgTextureNode {
    int32_t tex1;        /* SCALED: `tex`;  MIX / DIRECTION_MIX: `tex1`                          */
    int32_t tex2;        /* SCALED: `scale` (float texture);  MIX / DIRECTION_MIX: `tex2`        */
    int32_t amount;      /* MIX: `amount` (float texture)                                         */
    int32_t spectrum;    /* CONSTANT spectrum texture: spectrum id                                */
    float   value;       /* CONSTANT float texture                                                */
    float   dir[3];      /* DIRECTION_MIX: `dir` as given (default 0 1 0, render space)           */
} SgTextureNode;
/* `SphericalMapping`, `CylindricalMapping`, `PlanarMapping` (texture.rs:938-1035) as written there -- including the spherical
 * mapping's st = (theta/pi, theta/2pi) (both from theta, :960-963) and the cylindrical s = pi + atan2(y, x)/2pi (:991). */
typedef enum SgTextureMappingKind { SG_MAPPING_SPHERICAL = 1, SG_MAPPING_CYLINDRICAL = 2, SG_MAPPING_PLANAR = 3 } SgTextureMappingKind;
typedef struct SgTextureMapping {
    int32_t kind;                     /* SgTextureMappingKind                                   */
    float   texture_from_render[16];  /* Transform::m, row-major (affine)                       */
    float   vs[3], vt[3];             /* planar: `v1`, `v2`                                     */
    float   ds, dt;                   /* planar: `udelta`, `vdelta`                             */
    int32_t pad[3];
} SgTextureMapping;

/* ---- lights (src/light.rs) ------------------------------------------------- */
typedef enum SgLightKind {
    SG_LIGHT_DIFFUSE_AREA = 0,     /* DiffuseAreaLight light.rs:524-694 over one Triangle */
    SG_LIGHT_POINT = 1,            /* PointLight light.rs:403-519                         */
    SG_LIGHT_UNIFORM_INFINITE = 2, /* UniformInfiniteLight light.rs:697-803               */
    SG_LIGHT_DIFFUSE_AREA_SPHERE = 3, /* DiffuseAreaLight over a Sphere (sphere.rs:299-457): tri = index into spheres, area = Sphere::area() */
    SG_LIGHT_DIFFUSE_AREA_PATCH = 5, /* DiffuseAreaLight over a BilinearPatch (bilinear_patch.rs:517-783): mesh = the SG_MESH_BILINEAR mesh, tri = patch
                                        index, area = BilinearPatch::new's area (:40-76).  Rectangular patches are sampled by solid angle
                                        (sample_spherical_rectangle, sampling.rs:501-579), others by area with the bilinear warp */
    SG_LIGHT_IMAGE_INFINITE = 4    /* ImageInfinitelight light.rs:805-981: tri = index into env_maps, spectrum = the image colour space's
                                      illuminant (dense), scale as computed by Light::create (light.rs:181-223) */
} SgLightKind;
typedef struct SgLight {
    int32_t kind;
    int32_t spectrum;     /* dense 360..830 table: l_emit / i                              */
    float   scale;        /* already divided by spectrum_to_photometric (light.rs:583)     */
    int32_t two_sided;
    uint32_t mesh, tri;   /* area: the Triangle the light samples (scene.rs:609-622)       */
    float   area;         /* Shape::area() cached at construction (light.rs:546)           */
    float   pos[3];       /* point: render_from_light(0,0,0)                               */
    float   scene_center[3];
    float   scene_radius; /* infinite: preprocess() result (light.rs:797-802)              */
    float   pad[2];
} SgLight;

/* `PiecewiseConstant2D` over [0,1]^2 (sampling.rs:101-179) exactly as `new` builds it: per row v the conditional
 * `func` (|f|, nu floats) and `cdf` (nu + 1 floats), then the marginal over the row integrals (func nv floats, cdf nv + 1 floats,
 * `func_int`).  All arrays live in SgSceneDesc.spectrum_pool (float offsets).  The row integral `conditional_v[v].func_int`
 * equals marg_func[v]. */
typedef struct SgDistribution2D {
    int32_t  nu, nv;
    uint32_t func_off;        /* nv rows of nu floats          */
    uint32_t cdf_off;         /* nv rows of nu + 1 floats      */
    uint32_t marg_func_off;   /* nv floats                     */
    uint32_t marg_cdf_off;    /* nv + 1 floats                 */
    float    marg_integral;   /* marginal.func_int             */
    uint32_t pad;
} SgDistribution2D;
/* The environment map of an `ImageInfinitelight` (light.rs:805-981): a square RGB image in the equal-area octahedral
 * parameterisation, res x res x 3 linear f32 texels at texel_offset in SgSceneDesc.texels (row 0 first), looked up with
 * `lookup_nearest_channel_wrapped(.., OctahedralSphere)` (image.rs:134-162,590-601) and turned into a spectrum per lookup by
 * RgbIlluminantSpectrum::new (spectrum.rs:566-606; needs the rgb2spec table).  `distribution` is built from
 * Image::get_default_sampling_distribution (image.rs:1379-1405: channel average per pixel), `compensated` from the same values
 * minus their mean, clamped at 0 (light.rs:941-948); the path integrator samples the compensated one (allow_incomplete_pdf),
 * SimplePath the plain one. */
typedef struct SgEnvMap {
    float    render_from_light[16];
    float    light_from_render[16];
    uint64_t texel_offset;
    int32_t  res;
    int32_t  pad;
    SgDistribution2D distribution;
    SgDistribution2D compensated;
} SgEnvMap;

/* ---- camera (src/camera.rs:830-1114 PerspectiveCamera) --------------------- */
/* SG_CAMERA_ORTHOGRAPHIC = `OrthographicCamera` (camera.rs:657-827): ray origin camera_from_raster(p_film), direction +z,
 * auxiliary origins shifted by dx_camera / dy_camera, no depth of field (the reference has a TODO there).  Reproduced as
 * written: `generate_ray_differential` (:760-784), the entry the integrator calls, returns the ray in CAMERA space -- it
 * never applies render_from_camera (generate_ray :737-758 does) -- so images are only right when render space == camera
 * space up to translation-free axes; host code that wants the intended camera passes render_from_camera = identity scenes. */
typedef enum SgCameraKind { SG_CAMERA_PERSPECTIVE = 0, SG_CAMERA_ORTHOGRAPHIC = 1 } SgCameraKind;
typedef struct SgCamera {
    float camera_from_raster[16];     /* row-major 4x4, ProjectiveCameraBase camera.rs:632 */
    float render_from_camera[16];     /* CameraTransform camera.rs:518                     */
    float camera_from_render[16];     /* its inverse (Transform::m_inv)                     */
    float dx_camera[3];
    float dy_camera[3];
    float lens_radius;
    float focal_distance;
    float shutter_open;
    float shutter_close;
    float min_pos_differential_x[3], min_pos_differential_y[3];
    float min_dir_differential_x[3], min_dir_differential_y[3];
    int32_t kind;                     /* SgCameraKind                                       */
    int32_t pad;
} SgCamera;

/* ---- film (src/film.rs RgbFilm + PixelSensor) ------------------------------ */
typedef struct SgFilm {
    int32_t full_resolution[2];
    int32_t pixel_bounds[4];      /* min.x, min.y, max.x, max.y (max exclusive)            */
    float   filter_radius[2];     /* BoxFilter radius, filter.rs:64-106                    */
    int32_t r_bar, g_bar, b_bar;  /* dense spectra ids, PixelSensor film.rs:754-765         */
    float   imaging_ratio;
    float   max_component_value;  /* RgbFilm::max_component_value (film.rs:462), +inf default */
    float   output_rgb_from_sensor_rgb[9]; /* used only by sg_film_develop                  */
} SgFilm;

typedef struct SgSceneDesc {
    uint32_t abi_version;         /* SG_ABI_VERSION */
    uint32_t n_nodes;      const SgBvhNode*   nodes;
    uint32_t n_primitives; const SgPrimitive* primitives;
    uint32_t n_top_nodes;                                   /* top-level BVH = nodes[0, n_top_nodes); 0 = n_nodes           */
    uint32_t n_top_primitives;                              /* top-level primitives;                  0 = n_primitives      */
    uint32_t n_objects;    const SgObject*    objects;
    uint32_t n_instances;  const SgInstance*  instances;
    uint32_t n_spheres;    const SgSphere*    spheres;     /* referenced by SG_PRIM_SPHERE primitives (top level or inside objects) */
    uint32_t scene_flags;                                   /* SG_SCENE_*                                                   */
    uint32_t n_meshes;     const SgMesh*      meshes;
    uint32_t n_indices;    const uint32_t*    indices;     /* 3 per triangle               */
    uint32_t n_vertices;   const float*       p;           /* xyz per vertex               */
                           const float*       n;           /* xyz per vertex or NULL       */
                           const float*       uv;          /* uv per vertex or NULL        */
                           const float*       s;           /* xyz per vertex or NULL       */
    uint32_t n_spectra;    const SgSpectrum*  spectra;
    uint32_t n_pool;       const float*       spectrum_pool;
    uint32_t n_materials;  const SgMaterial*  materials;
    uint32_t n_lights;     const SgLight*     lights;      /* order = light sampler order  */
    uint32_t n_textures;   const SgTexture*   textures;
    uint32_t n_image_levels; const SgImageLevel* image_levels;
    uint64_t n_texels;     const float*       texels;
    const float* mip_filter_lut;                            /* MIP_FILTER_LUT[128], mipmap.rs:388-518 (EWA) or NULL */
    /* rgb2spec coefficient table of the scene colour space (rgb_to_spectra.rs:16-45; third-party `rgb2spec` 0.1.1,
     * RGB2Spec{res, scale[res], data[3*res^3*3]}); required iff a three-channel texture exists */
    uint32_t rgb2spec_res; const float* rgb2spec_scale; const float* rgb2spec_data;
    uint32_t n_texture_mappings; const SgTextureMapping* texture_mappings;
    uint32_t n_env_maps;   const SgEnvMap*    env_maps;     /* texels / spectrum_pool hold their data */
    uint32_t n_texture_nodes; const SgTextureNode* texture_nodes;   /* operands of the non-image textures */
    const SgMaterialTextures* material_textures;            /* n_materials rows or NULL (no texture-valued parameters)      */
    SgCamera camera;
    SgFilm   film;
} SgSceneDesc;

/* ---- render parameters: Options (options.rs:15-36), sampler (sampler.rs:95-99),
 *      integrator parameters (integrator.rs:188-192) ----------------------------- */
enum {
    SG_OPT_DISABLE_PIXEL_JITTER = 1,
    SG_OPT_DISABLE_WAVELENGTH_JITTER = 2,
    SG_OPT_DISABLE_TEXTURE_FILTERING = 4,
    SG_OPT_FORCE_DIFFUSE = 8      /* interaction.rs:258-273: every BSDF -> DiffuseBxDF(rho_hd(wo, 1 sample)); path integrator only */
};
typedef struct SgRenderParams {
    uint64_t seed;              /* IndependentSampler seed                                  */
    int32_t  samples_per_pixel; /* `pixelsamples`: used for ray-differential scale          */
    int32_t  sample_begin;      /* this call renders sample indices [begin, end) of every   */
    int32_t  sample_end;        /*   pixel: the multi-GPU split (DESIGN.md section e)       */
    int32_t  max_depth;         /* `maxdepth`, default 5                                    */
    int32_t  regularize;        /* `regularize`, default false                              */
    uint32_t option_flags;      /* SG_OPT_*                                                 */
    int32_t  max_paths_in_flight; /* wavefront width; 0 = library default (64 Mi paths, 18.5 GB) */
    int32_t  flags;             /* SG_RENDER_*                                              */
    int32_t  integrator;        /* SgIntegratorKind: the `Integrator` registry name (integrator.rs:16-42) */
    int32_t  integrator_flags;  /* SG_SIMPLEPATH_* (`samplelights`, `samplebsdf`, both default true: integrator.rs:24-31) */
} SgRenderParams;
/* "path" = PathIntegrator (integrator.rs:730-963), "simplepath" = SimplePathIntegrator (:570-728: no MIS, no Russian roulette,
 * light sampling with complete pdfs), "randomwalk" = RandomWalkIntegrator (:458-568: uniform sphere sampling, no light sampling). */
typedef enum SgIntegratorKind { SG_INTEGRATOR_PATH = 0, SG_INTEGRATOR_SIMPLE_PATH = 1, SG_INTEGRATOR_RANDOM_WALK = 2 } SgIntegratorKind;
enum { SG_SIMPLEPATH_SAMPLE_LIGHTS = 1, SG_SIMPLEPATH_SAMPLE_BSDF = 2 };
enum {
    SG_RENDER_COUNT_VISITS = 1,   /* count BVH nodes / triangles tested (slower; roofline accounting) */
    SG_RENDER_TIME_KERNELS = 2,   /* CUDA events around every traversal launch -> closest_ms/shadow_ms */
    SG_RENDER_OVERWRITE_FILM = 4, /* sg_render: store the film instead of adding to the caller's sums  */
    /* multi-GPU (the tile fan-out of integrator.rs:235-245, by sample range instead of by tile): */
    SG_RENDER_SPLIT_SAMPLES = 8,  /* with a process communicator (sg_comm_init_rank): this rank renders ITS share of
                                     [sample_begin, sample_end) -- rank r of n gets a contiguous range, remainder to the low
                                     ranks (sg_sample_range_for_rank).  Implied for sg_render after sg_init_multi.             */
    SG_RENDER_REDUCE_FILM = 16    /* with a process communicator: after the render, ONE ncclReduce (f64 sum) of the film onto
                                     rank 0, in stream order.  sg_render: only rank 0's `film` is written (other ranks may pass
                                     NULL); sg_render_device: d_film is reduced in place.  Implied for sg_render after
                                     sg_init_multi.                                                                            */
};

/* `RgbFilmPixel` without the (unused on this path) splat: film.rs:470-479. */
typedef struct SgFilmPixel {
    double rgb_sum[3];
    double weight_sum;
} SgFilmPixel;

typedef struct SgStats {
    uint64_t camera_paths;
    uint64_t closest_hit_rays;
    uint64_t shadow_rays;
    uint64_t nodes_visited;     /* bounds tests, counted in reference traversal order      */
    uint64_t tris_tested;       /* intersect_triangle entries                              */
    uint64_t kernel_launches;
    double   render_ms;         /* device time of the wavefront loop                       */
    double   trace_ms;          /* device time inside closest-hit + any-hit kernels        */
    uint64_t closest_nodes;     /* nodes_visited / tris_tested of the closest-hit rays only */
    uint64_t closest_tris;
    uint64_t closest_launches;  /* closest-hit / any-hit traversal kernel launches          */
    uint64_t shadow_launches;
    double   closest_ms;        /* device time of the closest-hit kernels (reserved bit 1)  */
    double   shadow_ms;         /* device time of the any-hit kernels     (reserved bit 1)  */
    /* multi-GPU: render_ms is the slowest device's wavefront loop; ray / path counters are summed over the devices this
     * process drives (one process per GPU: this rank's only) */
    double   reduce_ms;         /* device time of the NCCL film reduce on the root (includes waiting for the slowest rank) */
    double   d2h_ms;            /* sg_render: device -> pinned host copy of the (reduced) film                              */
    uint32_t n_devices;         /* devices that rendered in this call (single process) or ranks of the communicator         */
    uint32_t rank;              /* this process's rank in the communicator (0 without one)                                  */
} SgStats;

/* `ShapeIntersection` reduced to what parity needs (shape.rs:221-225 +
 * TriangleIntersection triangle.rs:748-755 + geometric normal triangle.rs:407-412). */
typedef struct SgHit {
    int32_t prim;     /* index into SgSceneDesc.primitives (for an instanced hit: the primitive inside the object), -1 = miss */
    float   t;
    float   b0, b1, b2;
    float   ng[3];
} SgHit;

typedef struct SgScene SgScene;

/* Selects the CUDA device of this process and creates the library's stream.  Calling it again with another device moves
 * the library there (scenes created before stay on their device and must be destroyed first). */
int sg_init(int device);
/* Single process, n GPUs (`sg_init(const int* devices, int n)`): one stream per device and one in-process NCCL
 * communicator (ncclCommInitAll).  Afterwards sg_scene_create replicates the scene on every device and sg_render splits
 * [sample_begin, sample_end) across them (rayon's tile fan-out, integrator.rs:235-245, becomes a sample-range fan-out),
 * sums the films onto devices[0] with one ncclReduce and copies the result to the caller's host film.  Every other entry
 * point (sg_render_device, sg_trace, ...) runs on devices[0].  n == 1 is sg_init(devices[0]). */
int sg_init_multi(const int* devices, int n);
int sg_device_count(void);      /* devices this process drives (0 before sg_init) */
int sg_shutdown(void);

/* One process per GPU: a communicator across processes.  Rank 0 calls sg_comm_get_unique_id and hands the
 * SG_COMM_ID_BYTES bytes to the other ranks by any means (a file, MPI, torch.distributed broadcast, a socket);
 * then EVERY rank calls sg_comm_init_rank (collective).  sg_comm_destroy is collective too. */
#define SG_COMM_ID_BYTES 128
int sg_comm_get_unique_id(void* id_out);
int sg_comm_init_rank(const void* id, int rank, int n_ranks);
int sg_comm_destroy(void);
int sg_comm_rank(int* rank, int* n_ranks);   /* 0 / 1 without a communicator */
/* The sample-index range rank `rank` of `n_ranks` renders out of [begin, end): contiguous, remainder to the low ranks. */
int sg_sample_range_for_rank(int32_t begin, int32_t end, int rank, int n_ranks, int32_t* out_begin, int32_t* out_end);
/* In-place ncclReduce (sum of f64) of a device film onto rank 0 of the process communicator, on `stream`
 * (a no-op without a communicator or with one rank).  n_pixels SgFilmPixel records. */
int sg_film_reduce_device(void* d_film, int64_t n_pixels, void* stream);
const char* sg_last_error(void);
int sg_abi_version(void);

/* Replaces the object graph `render_cpu` builds before `integrator.render`
 * (render.rs:33-52): flattens nothing itself, uploads the host-flattened scene. */
int sg_scene_create(const SgSceneDesc* desc, SgScene** out);
int sg_scene_destroy(SgScene* scene);

/* Replaces `ImageTileIntegrator::render` (integrator.rs:227-321) for the `path`
 * integrator: evaluates samples [sample_begin,sample_end) of every pixel and
 * ADDS them into `film` (row-major (y-y0)*W+(x-x0), vec2d.rs:24-28).
 * sg_render: `film` is host memory (copied back inside the call).
 * sg_render_device: `film` is device memory of this process's GPU, zeroed by
 * the caller; `stream` is the cudaStream_t the caller's own work on d_film is ordered on -- every kernel (and the
 * optional reduce) is enqueued there.  NULL is the CUDA legacy default stream (what a handle of 0 means everywhere
 * else, e.g. torch.cuda.current_stream().cuda_stream on torch's default stream), NOT a private library stream. */
int sg_render(SgScene* scene, const SgRenderParams* params, SgFilmPixel* film, SgStats* stats);
int sg_render_device(SgScene* scene, const SgRenderParams* params, void* d_film, SgStats* stats, void* stream);

/* Replaces `BvhAggregate::intersect` / `intersect_predicate`
 * (aggregate.rs:71-203) for a batch of rays: the ray-cast parity entry.
 * o,d: xyz per ray; any_hit!=0 -> predicate (out[i].prim is 0 on hit, -1 on miss). */
int sg_trace(SgScene* scene, int64_t n, const float* o, const float* d, const float* t_max,
             int any_hit, SgHit* out, SgStats* stats);
/* Same with every buffer in device memory (bench: inputs resident in HBM). */
int sg_trace_device(SgScene* scene, int64_t n, const void* d_o, const void* d_d, const void* d_t_max,
                    int any_hit, void* d_out, SgStats* stats, void* stream);

/* Replaces `IndependentSampler::get_1d` (sampler.rs:123-125) over the stream
 * assigned to (pixel_index, sample_index): RNG known-answer entry.
 * raw!=0 -> seed the generator with `seed` directly (SmallRng::seed_from_u64). */
int sg_sampler_fill(uint64_t seed, int raw, uint32_t pixel_index, uint32_t sample_index,
                    int64_t n, float* out);

/* Replaces `evaluate_pixel_sample`'s camera stage (integrator.rs:339-362):
 * wavelengths + camera ray for n (pixel,sample) pairs; out_rays: o.xyz d.xyz per
 * sample, out_lambda: 4 lambdas + 4 pdfs per sample. */
int sg_camera_rays(SgScene* scene, const SgRenderParams* params, int64_t n,
                   const int32_t* pixel_xy, const int32_t* sample_index,
                   float* out_rays, float* out_lambda);

/* Replaces `SpectrumImageTexture::evaluate` (texture.rs:777-808; as_float!=0: `FloatImageTexture::evaluate`
 * :393-404, value replicated) for n lookups of texture `tex`: the texture-filtering parity entry.
 * q: u v dudx dudy dvdx dvdy per lookup (TextureEvalContext); lambda: 4 wavelengths per lookup; out: 4 floats. */
int sg_texture_eval(SgScene* scene, int tex, int as_float, int64_t n, const float* q, const float* lambda, float* out);
/* Same with the full `TextureEvalContext` (texture.rs): pdp = p, dpdx, dpdy in render space (9 floats per lookup), which the
 * spherical / cylindrical / planar mappings read (texture.rs:938-1035). */
int sg_texture_eval_p(SgScene* scene, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* lambda, float* out);
/* Same with `TextureEvalContext::n` as well (3 floats per lookup), which the direction-mix textures read (texture.rs:299,:690). */
int sg_texture_eval_ctx(SgScene* scene, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* nrm, const float* lambda, float* out);

/* Replaces `RgbFilm::get_pixel_rgb` (film.rs:720-738): rgb_sum/weight_sum then
 * output_rgb_from_sensor_rgb; out: 3 floats per pixel. */
int sg_film_develop(SgScene* scene, const SgFilmPixel* film, int64_t n_pixels, float* out_rgb);

/* Replaces `RgbFilm::get_image` (film.rs:647-707): get_pixel_rgb per pixel, the fp16 clamp exactly as written there
 * (when any channel exceeds 65504: r and b are clamped, a too-large g clamps r AGAIN and stays -- film.rs:683-685),
 * then the pixel store of `Image::set_channel` (image.rs:648-661: NaN -> 0; PixelFormat::Half rounds through
 * `half::f16::from_f32`, IEEE round-to-nearest-even) and the read-back `Image::write_pfm` does (image.rs:1352-1358:
 * f16 -> f32).  SG_IMAGE_FP16 = `savefp16` (film.rs:491, default true); SG_IMAGE_BOTTOM_UP emits rows in PFM raster
 * order (bottom row first, image.rs:1350) so the buffer can be written straight after the "PF\nW H\n-1\n" header.
 * film / out_rgb: width*height pixels, row-major; out: 3 floats per pixel. */
enum SgImageFlags { SG_IMAGE_FP16 = 1, SG_IMAGE_BOTTOM_UP = 2 };
int sg_film_get_image(SgScene* scene, const SgFilmPixel* film, int32_t width, int32_t height, uint32_t flags, float* out_rgb);

/* Replaces `Image::generate_pyramid` (image.rs:699-787; called by MIPMap::new, mipmap.rs:19-106) including `Image::float_resize_up`
 * (:1007-1111) for images whose resolution is not a power of two: the stage right before the texture lookups of the hot path.
 * sg_image_pyramid_layout fills `levels` (capacity >= 32; offsets in floats from the start of the output buffer) and the total
 * texel count; sg_image_generate_pyramid takes width x height x n_channels LINEAR f32 texels (row 0 first; what
 * Image::convert_to_format(Float) holds) and writes every level back to back into `out_texels` (host memory).
 * wrap: SG_WRAP_REPEAT or SG_WRAP_CLAMP (the reference asserts on `black` while resizing, image.rs:826-828).  Resizing requires
 * BOTH dimensions to grow (image.rs:1009-1010): e.g. 64 x 50 is rejected like the reference's assert.  The layout can be copied
 * into SgImageLevel rows (add the texture's base offset) and the texels into SgSceneDesc.texels. */
int sg_image_pyramid_layout(int32_t width, int32_t height, int32_t n_channels, int32_t* n_levels, SgImageLevel* levels, uint64_t* n_texels);
int sg_image_generate_pyramid(const float* image, int32_t width, int32_t height, int32_t n_channels, int32_t wrap, float* out_texels);

#ifdef __cplusplus
}
#endif
#endif /* SHIMMER_GPU_H */
