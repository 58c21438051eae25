// Device-resident scene layout + BVH traversal / watertight triangle test.
//
// HBM layout (DESIGN.md "Data layout"):
//   nodes     : SgBvhNode[n_nodes], 32 B each, 2 x LDG.128 per visit (reference order:
//               depth-first, first child = index+1, aggregate.rs:425-467)
//   tri_verts : float4[3*n_prims] in BVH-leaf (ordered-primitive) order -- the three
//               render-space vertices pre-gathered through vertex_indices so a triangle
//               test costs 3 x LDG.128 instead of the reference's 5 dependent loads
//               (Arc<Primitive> -> Arc<Shape> -> Arc<TriangleMesh> -> indices -> p).
//               .w lanes: [0] material | mesh flags<<23 | kind<<28, [1] area-light id,
//               [2] mesh id | last-in-leaf<<31.
//   everything else is the host arrays copied verbatim.
#pragma once
#include "sg_math.cuh"
#include "../../include/shimmer_gpu.h"

namespace sg {

// One `TransformedPrimitive` (primitive.rs:136-176) as the kernels want it: both 3x4 matrices, the root of the object's
// BvhAggregate (Node64 ref, or a leaf ref for a bare primitive) and that root's bounds.
struct DInstance {
    float m[12];            // render_from_primitive, rows 0..2
    float mi[12];           // primitive_from_render, rows 0..2
    float bmin[3], bmax[3]; // bounds of the object's root node (tested on entry like aggregate.rs:92-98)
    uint32_t root_ref;
    uint32_t root_has_bounds;   // 0: bare primitive without an aggregate (scene.rs:831-833)
    uint32_t pad[2];
};
static constexpr uint32_t kKindInstance = 7u;     // tri_verts[3*i].w kind field of an instance primitive

struct DScene {
    const float4* nodes;        // 2 float4 per node
    const float4* tri_verts;    // 3 float4 per primitive
    const float4* light_verts;  // 3 float4 per light (area lights: the emitter triangle; .w[0] = mesh flags)
    const SgPrimitive* prims;
    const SgMesh* meshes;
    const uint32_t* indices;
    const float* p;
    const float* n;
    const float* uv;
    const float* s;
    const SgSpectrum* spectra;          // device copy: `pad` holds 1 + the spectrum's offset into spec_lut (0: none)
    const float* pool;
    const uint16_t* spec_lut;           // piecewise-linear spectra: interval index at every integer wavelength 360..830 (sg_shading.cuh spectrum_get)
    const SgMaterial* materials;
    const SgLight* lights;
    const SgTexture* textures;          // image textures (sg_texture.cuh)
    const SgImageLevel* image_levels;
    const float* texels;
    const float* mip_lut;               // MIP_FILTER_LUT[128]
    const float* rgb2spec_scale;        // rgb2spec table of the scene colour space
    const float* rgb2spec_data;
    uint32_t rgb2spec_res, n_textures;
    const SgTextureMapping* texture_mappings;   // spherical / cylindrical / planar mappings (sg_texture.cuh)
    const SgTextureNode* texture_nodes;         // operands of the constant / scaled / mix / direction-mix textures (sg_texture.cuh)
    const SgMaterialTextures* material_textures; // texture-valued material parameters, one row per material, or null
    const SgEnvMap* env_maps;           // ImageInfinitelight images + sampling distributions (sg_envmap.cuh)
    const DInstance* instances;         // object instancing
    const struct DSphere* spheres;      // sphere shapes (sg_sphere.cuh)
    const float4* patch_verts;          // bilinear patches: 4 float4 per patch primitive (sg_patch.cuh)
    uint32_t n_instances, scene_flags;
    uint32_t n_nodes, n_prims, n_lights, n_materials;
    int32_t n_infinite;          // number of infinite lights (uniform + image)
    int32_t infinite_ids[4];
    SgCamera camera;
    SgFilm film;
};

struct RayPre {                 // per-ray constants of the watertight test, triangle.rs:197-215
    int kx, ky, kz;
    float sx, sy, sz;
};
SGD RayPre ray_precompute(float3 d) {
    RayPre r;
    r.kz = maxcomp_index(abs3(d));
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    float dx = comp3(d, r.kx), dy = comp3(d, r.ky), dz = comp3(d, r.kz);
    r.sx = -dx / dz; r.sy = -dy / dz; r.sz = 1.0f / dz;
    return r;
}
SGD float3 permute3(float3 v, int kx, int ky, int kz) { return f3(comp3(v, kx), comp3(v, ky), comp3(v, kz)); }

// Triangle::intersect_triangle, triangle.rs:173-302.  Returns true on an accepted hit.
// The degenerate-triangle test of triangle.rs:181 depends on the triangle only: the traversal kernels read it from a flag that
// k_mark_degenerate (below) sets once per scene with this very expression (kDegenerateBit in tri_verts[3*i].w) and call
// intersect_triangle<false>; every other caller evaluates it in place.
static constexpr uint32_t kDegenerateBit = 0x80000000u;
SGD bool triangle_is_degenerate(float3 p0, float3 p1, float3 p2) { return len2(cross3(p2 - p0, p1 - p0)) == 0.0f; }
template <bool CHECK_DEGENERATE = true>
SGD bool intersect_triangle(float3 o, const RayPre& rp, float t_max, float3 p0, float3 p1, float3 p2,
                            float& b0, float& b1, float& b2, float& t) {
    if (CHECK_DEGENERATE && triangle_is_degenerate(p0, p1, p2)) return false;   // degenerate :181
    float3 p0t = permute3(p0 - o, rp.kx, rp.ky, rp.kz);
    float3 p1t = permute3(p1 - o, rp.kx, rp.ky, rp.kz);
    float3 p2t = permute3(p2 - o, rp.kx, rp.ky, rp.kz);
    p0t.x += rp.sx * p0t.z; p0t.y += rp.sy * p0t.z;
    p1t.x += rp.sx * p1t.z; p1t.y += rp.sy * p1t.z;
    p2t.x += rp.sx * p2t.z; p2t.y += rp.sy * p2t.z;
    float e0 = dop(p1t.x, p2t.y, p1t.y, p2t.x);
    float e1 = dop(p2t.x, p0t.y, p2t.y, p0t.x);
    float e2 = dop(p0t.x, p1t.y, p0t.y, p1t.x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {                                // f64 fallback :232-242
        double p2txp1ty = (double)p2t.x * (double)p1t.y;
        double p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y;
        double p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y;
        double p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= rp.sz; p1t.z *= rp.sz; p2t.z *= rp.sz;
    float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
    float inv_det = 1.0f / det;
    float tt = t_scaled * inv_det;
    // conservative t > delta_t test :272-299
    float max_zt = maxcomp(abs3(f3(p0t.z, p1t.z, p2t.z)));
    float delta_z = gamma_n(3) * max_zt;
    float max_xt = maxcomp(abs3(f3(p0t.x, p1t.x, p2t.x)));
    float max_yt = maxcomp(abs3(f3(p0t.y, p1t.y, p2t.y)));
    float delta_x = gamma_n(5) * (max_xt + max_zt);
    float delta_y = gamma_n(5) * (max_yt + max_zt);
    float delta_e = 2.0f * (gamma_n(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    float max_e = maxcomp(abs3(f3(e0, e1, e2)));
    float delta_t = 3.0f * (gamma_n(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * fabsf(inv_det);
    if (tt <= delta_t) return false;
    b0 = e0 * inv_det; b1 = e1 * inv_det; b2 = e2 * inv_det; t = tt;
    return true;
}

// Bounds3f::intersect_p_cached, bounding_box.rs:520-564.  `lo`/`hi` are the node's
// min/max; NaN comparisons evaluate false exactly as on the CPU (0*inf slabs).
SGD bool slab_test(float3 bmin, float3 bmax, float3 o, float3 inv_dir, int nx, int ny, int nz, float ray_t_max) {
    const float k = 1.0f + 2.0f * gamma_n(3);
    float t_min = ((nx ? bmax.x : bmin.x) - o.x) * inv_dir.x;
    float t_max = ((nx ? bmin.x : bmax.x) - o.x) * inv_dir.x;
    float ty_min = ((ny ? bmax.y : bmin.y) - o.y) * inv_dir.y;
    float ty_max = ((ny ? bmin.y : bmax.y) - o.y) * inv_dir.y;
    t_max *= k; ty_max *= k;
    if (t_min > ty_max || ty_min > t_max) return false;
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    float tz_min = ((nz ? bmax.z : bmin.z) - o.z) * inv_dir.z;
    float tz_max = ((nz ? bmin.z : bmax.z) - o.z) * inv_dir.z;
    tz_max *= k;
    if (t_min > tz_max || tz_min > t_max) return false;
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    return t_min < ray_t_max && t_max > 0.0f;
}

struct HitRec { int prim; float t, b0, b1, b2; int inst; };

// BvhAggregate::intersect (ANY=false, aggregate.rs:71-139) / intersect_predicate (ANY=true,
// :141-203): same visiting order (near child by dir_is_neg[axis], far child pushed), same
// strict comparisons, so ties resolve exactly as in the reference.
// `stack` points at this thread's 64-entry stack (stride `sstride` words between levels).
template <bool ANY, bool COUNT>
SGD bool traverse(const DScene& sc, float3 o, float3 d, float t_max, HitRec& hit,
                  uint32_t* stack, int sstride, uint32_t& n_nodes, uint32_t& n_tris) {
    hit.prim = -1;
    if (sc.n_nodes == 0) return false;
    float3 inv_dir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    const int nx = inv_dir.x < 0.0f, ny = inv_dir.y < 0.0f, nz = inv_dir.z < 0.0f;
    const RayPre rp = ray_precompute(d);
    int sp = 0;
    uint32_t cur = 0;
    bool found = false;
    for (;;) {
        const float4 n0 = __ldg(sc.nodes + 2 * (size_t)cur);
        const float4 n1 = __ldg(sc.nodes + 2 * (size_t)cur + 1);
        if (COUNT) n_nodes++;
        // n0 = (min.x, min.y, min.z, max.x)  n1 = (max.y, max.z, offset, n_prims|axis<<16)
        const uint32_t meta = __float_as_uint(n1.w);
        const uint32_t offset = __float_as_uint(n1.z);
        bool descend = slab_test(f3(n0.x, n0.y, n0.z), f3(n0.w, n1.x, n1.y), o, inv_dir, nx, ny, nz, t_max);
        if (descend) {
            const uint32_t n_prims = meta & 0xffffu;
            if (n_prims > 0) {
                for (uint32_t i = 0; i < n_prims; ++i) {
                    const uint32_t pi = offset + i;
                    const float4 v0 = __ldg(sc.tri_verts + 3 * (size_t)pi);
                    const float4 v1 = __ldg(sc.tri_verts + 3 * (size_t)pi + 1);
                    const float4 v2 = __ldg(sc.tri_verts + 3 * (size_t)pi + 2);
                    if (COUNT) n_tris++;
                    float b0, b1, b2, t;
                    if (intersect_triangle(o, rp, t_max, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), b0, b1, b2, t)) {
                        hit.prim = (int)pi; hit.t = t; hit.b0 = b0; hit.b1 = b1; hit.b2 = b2;
                        if (ANY) return true;
                        t_max = t; found = true;
                    }
                }
                if (sp == 0) break;
                cur = stack[(--sp) * sstride];
            } else {
                const uint32_t axis = (meta >> 16) & 0xffu;
                const int neg = axis == 0 ? nx : (axis == 1 ? ny : nz);
                if (neg) { stack[(sp++) * sstride] = cur + 1; cur = offset; }
                else { stack[(sp++) * sstride] = offset; cur = cur + 1; }
            }
        } else {
            if (sp == 0) break;
            cur = stack[(--sp) * sstride];
        }
    }
    return found;
}

}  // namespace sg
