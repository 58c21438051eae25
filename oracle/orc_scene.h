// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).  parity status: see orc_math.h.
// Geometry layer of the CPU restatement: bounds, watertight triangle test, BVH build +
// traversal, surface interaction construction.  Consumes the same flattened
// SgSceneDesc the C ABI takes (include/shimmer_gpu.h).
#pragma once
#include "orc_math.h"
#include "../include/shimmer_gpu.h"
#include <vector>

namespace orc {

struct Ray { V3 o, d; };

struct Counters { uint64_t nodes = 0, tris = 0, closest = 0, shadow = 0; };

// bounding_box.rs:520-564
inline bool bounds_intersect_p_cached(const SgBvhNode& b, V3 o, Float ray_t_max, V3 inv_dir, const int dir_is_neg[3]) {
    const Float* lohi[2] = {b.bmin, b.bmax};
    const Float k = 1.0f + 2.0f * gamma_n(3);
    Float t_min = (lohi[dir_is_neg[0]][0] - o.x) * inv_dir.x;
    Float t_max = (lohi[1 - dir_is_neg[0]][0] - o.x) * inv_dir.x;
    Float ty_min = (lohi[dir_is_neg[1]][1] - o.y) * inv_dir.y;
    Float ty_max = (lohi[1 - dir_is_neg[1]][1] - o.y) * inv_dir.y;
    t_max *= k;
    ty_max *= k;
    if (t_min > ty_max || ty_min > t_max) return false;
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    Float tz_min = (lohi[dir_is_neg[2]][2] - o.z) * inv_dir.z;
    Float tz_max = (lohi[1 - dir_is_neg[2]][2] - o.z) * inv_dir.z;
    tz_max *= k;
    if (t_min > tz_max || tz_min > t_max) return false;
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    return t_min < ray_t_max && t_max > 0.0f;
}

struct TriHit { Float b0, b1, b2, t; };

// triangle.rs:173-302
inline bool intersect_triangle(const Ray& ray, Float t_max, V3 p0, V3 p1, V3 p2, TriHit* out) {
    if (length_squared(cross(p2 - p0, p1 - p0)) == 0.0f) return false;
    V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
    int kz = max_component_index(vabs(ray.d));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    V3 d = permute(ray.d, kx, ky, kz);
    p0t = permute(p0t, kx, ky, kz);
    p1t = permute(p1t, kx, ky, kz);
    p2t = permute(p2t, kx, ky, kz);
    Float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    Float e0 = difference_of_products(p1t.x, p2t.y, p1t.y, p2t.x);
    Float e1 = difference_of_products(p2t.x, p0t.y, p2t.y, p0t.x);
    Float e2 = difference_of_products(p0t.x, p1t.y, p0t.y, p1t.x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {             // :232-242 double fallback
        double p2txp1ty = (double)p2t.x * (double)p1t.y;
        double p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (Float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y;
        double p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (Float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y;
        double p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (Float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    Float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    Float t_scaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) return false;
    else if (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)) return false;
    Float inv_det = 1.0f / det;
    Float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    Float t = t_scaled * inv_det;
    Float max_zt = max_component_value(vabs(v3(p0t.z, p1t.z, p2t.z)));
    Float delta_z = gamma_n(3) * max_zt;
    Float max_xt = max_component_value(vabs(v3(p0t.x, p1t.x, p2t.x)));
    Float max_yt = max_component_value(vabs(v3(p0t.y, p1t.y, p2t.y)));
    Float delta_x = gamma_n(5) * (max_xt + max_zt);
    Float delta_y = gamma_n(5) * (max_yt + max_zt);
    Float delta_e = 2.0f * (gamma_n(2) * max_xt * max_yt + delta_y * max_xt + delta_x * max_yt);
    Float max_e = max_component_value(vabs(v3(e0, e1, e2)));
    Float delta_t = 3.0f * (gamma_n(3) * max_e * max_zt + delta_e * max_zt + delta_z * max_e) * std::fabs(inv_det);
    if (t <= delta_t) return false;
    out->b0 = b0; out->b1 = b1; out->b2 = b2; out->t = t;
    return true;
}

// Read-only view over the flattened scene.
struct Scene {
    const SgSceneDesc* d;
    explicit Scene(const SgSceneDesc* desc) : d(desc) {}
    inline V3 vertex(const SgMesh& m, uint32_t local) const {
        const float* q = d->p + 3 * (size_t)(m.first_vertex + local);
        return v3(q[0], q[1], q[2]);
    }
    inline V3 normal(const SgMesh& m, uint32_t local) const {
        const float* q = d->n + 3 * (size_t)(m.first_vertex + local);
        return v3(q[0], q[1], q[2]);
    }
    inline V2 uv(const SgMesh& m, uint32_t local) const {
        const float* q = d->uv + 2 * (size_t)(m.first_vertex + local);
        V2 r = {q[0], q[1]}; return r;
    }
    inline void tri_indices(uint32_t mesh, uint32_t tri, uint32_t v[3]) const {
        const SgMesh& m = d->meshes[mesh];
        const uint32_t* ix = d->indices + m.first_index + 3 * (size_t)tri;
        v[0] = ix[0]; v[1] = ix[1]; v[2] = ix[2];
    }
    inline void patch_indices(uint32_t mesh, uint32_t patch, uint32_t v[4]) const {      // bilinear_patch.rs:87-106
        const SgMesh& m = d->meshes[mesh];
        const uint32_t* ix = d->indices + m.first_index + 4 * (size_t)patch;
        v[0] = ix[0]; v[1] = ix[1]; v[2] = ix[2]; v[3] = ix[3];
    }
    inline void patch_points(uint32_t mesh, uint32_t patch, V3 q[4]) const {            // p00, p10, p01, p11
        uint32_t v[4]; patch_indices(mesh, patch, v);
        const SgMesh& m = d->meshes[mesh];
        for (int k = 0; k < 4; ++k) q[k] = vertex(m, v[k]);
    }
    inline void tri_points(uint32_t mesh, uint32_t tri, V3* p0, V3* p1, V3* p2) const {   // triangle.rs:148-159
        uint32_t v[3]; tri_indices(mesh, tri, v);
        const SgMesh& m = d->meshes[mesh];
        *p0 = vertex(m, v[0]); *p1 = vertex(m, v[1]); *p2 = vertex(m, v[2]);
    }
};

struct Hit { int32_t prim; int32_t inst; TriHit th; };   // sphere hits: th.b0..b2 = QuadricIntersection::p_obj, th.t = t_hit
}  // namespace orc
#include "orc_sphere.h"
#include "orc_patch.h"
namespace orc {

// Transform::apply_ray_inverse transform.rs:701-723 (inverse = true) / Transform::apply_ray :515-532 (inverse = false)
// with Some(t_max): the inverse variant sends the origin through the Point3fi transform of an EXACT point (:631-700; its error
// term omits the translation column) and shifts it to the edge of its error bounds, t_max reduced by dt; the forward variant
// transforms a plain Point3f, so its interval has zero width and dt == 0.
inline Ray instance_ray(const SgInstance& I, const Ray& r, bool inverse, Float* t_max) {
    const float* m = inverse ? I.primitive_from_render : I.render_from_primitive;
    Float x = r.o.x, y = r.o.y, z = r.o.z;
    P3fi o;
    if (inverse) {               // apply_ray_inverse: apply_inverse(Point3fi::from(o)), error term without the translation column
        Float xp = (m[0] * x + m[1] * y) + (m[2] * z + m[3]);
        Float yp = (m[4] * x + m[5] * y) + (m[6] * z + m[7]);
        Float zp = (m[8] * x + m[9] * y) + (m[10] * z + m[11]);
        V3 err = v3(gamma_n(3) * (std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z)),
                    gamma_n(3) * (std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z)),
                    gamma_n(3) * (std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z)));
        o = p3fi_from_value_and_error(v3(xp, yp, zp), err);
    } else {                     // apply_ray (:516-517): `self.apply(val.o)` is the Point3f overload (apply_point_helper :753-767, summed left
                                 // to right); `.into()` gives a ZERO-width interval, so dt below is 0 and only the interval add + midpoint remain
        V3 pp = v3(((m[0] * x + m[1] * y) + m[2] * z) + m[3], ((m[4] * x + m[5] * y) + m[6] * z) + m[7], ((m[8] * x + m[9] * y) + m[10] * z) + m[11]);
        o = p3fi_from_value_and_error(pp, v3(0.0f, 0.0f, 0.0f));
    }
    // instance transforms are affine (last row 0 0 0 1): the `/ wp` branches (:453-457, :761-766) are never taken
    V3 d = v3(m[0] * r.d.x + m[1] * r.d.y + m[2] * r.d.z, m[4] * r.d.x + m[5] * r.d.y + m[6] * r.d.z, m[8] * r.d.x + m[9] * r.d.y + m[10] * r.d.z);
    Float ls = length_squared(d);
    if (ls > 0.0f) {
        Float dt = dot(vabs(d), p3fi_error(o)) / ls;
        *t_max = *t_max - dt;
        V3 off = d * dt;
        o.lo = v3(next_float_down(o.lo.x + off.x), next_float_down(o.lo.y + off.y), next_float_down(o.lo.z + off.z));   // interval.rs:353-356
        o.hi = v3(next_float_up(o.hi.x + off.x), next_float_up(o.hi.y + off.y), next_float_up(o.hi.z + off.z));
    }
    Ray out; out.o = p3fi_mid(o); out.d = d; return out;
}

inline bool bvh_intersect_range(const Scene& sc, uint32_t node_base, uint32_t n_nodes, uint32_t prim_base, uint32_t n_prims,
                                const Ray& ray, Float t_max, bool any_hit, Hit* hit, Counters* ctr);

// One primitive: a Triangle behind Geometric/SimplePrimitive (primitive.rs:65-131) or a TransformedPrimitive (:136-176).
inline bool primitive_intersect(const Scene& sc, uint32_t pi, const Ray& ray, Float t_max, bool any_hit, Hit* h, Counters* ctr) {
    const SgSceneDesc* D = sc.d;
    const SgPrimitive& pr = D->primitives[pi];
    if (pr.mesh == SG_PRIM_INSTANCE) {
        const SgInstance& I = D->instances[pr.tri];
        const SgObject& O = D->objects[I.object];
        // closest hit: inverse transform (:159-163); predicate: FORWARD transform in the reference (:172-175)
        bool inverse = !any_hit || (D->scene_flags & SG_SCENE_FIX_INSTANCING);
        Float tm = t_max;
        Ray r2 = instance_ray(I, ray, inverse, &tm);
        Hit h2;
        if (!bvh_intersect_range(sc, O.first_node, O.n_nodes, O.first_prim, O.n_prims, r2, tm, any_hit, &h2, ctr)) return false;
        h->prim = h2.prim; h->inst = (int32_t)pr.tri; h->th = h2.th;
        return true;
    }
    if (pr.mesh == SG_PRIM_SPHERE) {                          // Shape::Sphere behind a Simple/GeometricPrimitive
        QuadricHit q;
        if (ctr) ctr->tris++;
        if (!sphere_basic_intersect(D->spheres[pr.tri], ray, t_max, &q)) return false;
        h->prim = (int32_t)pi; h->inst = -1;
        h->th.t = q.t; h->th.b0 = q.p_obj.x; h->th.b1 = q.p_obj.y; h->th.b2 = q.p_obj.z;
        return true;
    }
    if (D->meshes[pr.mesh].flags & SG_MESH_BILINEAR) {        // Shape::BilinearPatch: th.b0 = u, th.b1 = v
        V3 q[4]; sc.patch_points(pr.mesh, pr.tri, q);
        if (ctr) ctr->tris++;
        Float u, v, t;
        if (!intersect_blp(ray.o, ray.d, t_max, q[0], q[1], q[2], q[3], &u, &v, &t)) return false;
        h->prim = (int32_t)pi; h->inst = -1; h->th.t = t; h->th.b0 = u; h->th.b1 = v; h->th.b2 = 0.0f;
        return true;
    }
    V3 p0, p1, p2; sc.tri_points(pr.mesh, pr.tri, &p0, &p1, &p2);
    if (ctr) ctr->tris++;
    if (!intersect_triangle(ray, t_max, p0, p1, p2, &h->th)) return false;
    h->prim = (int32_t)pi; h->inst = -1;
    return true;
}

// aggregate.rs:71-139 (any_hit = false) and :141-203 (any_hit = true) over one BvhAggregate.
// The reference builds the full SurfaceInteraction for every accepted candidate
// (triangle.rs:529-535); only the last one survives, so it is built once by the caller.
inline bool bvh_intersect_range(const Scene& sc, uint32_t node_base, uint32_t n_nodes, uint32_t prim_base, uint32_t n_prims,
                                const Ray& ray, Float t_max, bool any_hit, Hit* hit, Counters* ctr) {
    const SgSceneDesc* D = sc.d;
    hit->prim = -1; hit->inst = -1;
    bool found = false;
    if (n_nodes == 0) {                                       // bare primitive(s) without an aggregate
        for (uint32_t i = 0; i < n_prims; ++i) {
            Hit h;
            if (primitive_intersect(sc, prim_base + i, ray, t_max, any_hit, &h, ctr)) {
                *hit = h; if (any_hit) return true;
                t_max = h.th.t; found = true;
            }
        }
        return found;
    }
    V3 inv_dir = v3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
    int dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    uint32_t to_visit = 0, cur = 0;
    uint32_t stack[64];
    for (;;) {
        const SgBvhNode& node = D->nodes[node_base + cur];
        if (ctr) ctr->nodes++;
        if (bounds_intersect_p_cached(node, ray.o, t_max, inv_dir, dir_is_neg)) {
            if (node.n_prims > 0) {
                for (uint32_t i = 0; i < node.n_prims; ++i) {
                    Hit h;
                    if (primitive_intersect(sc, prim_base + node.offset + i, ray, t_max, any_hit, &h, ctr)) {
                        *hit = h; if (any_hit) return true;
                        t_max = h.th.t; found = true;
                    }
                }
                if (to_visit == 0) break;
                cur = stack[--to_visit];
            } else {
                if (dir_is_neg[node.axis]) { stack[to_visit++] = cur + 1; cur = node.offset; }
                else { stack[to_visit++] = node.offset; cur = cur + 1; }
            }
        } else {
            if (to_visit == 0) break;
            cur = stack[--to_visit];
        }
    }
    return found;
}
inline bool bvh_intersect(const Scene& sc, const Ray& ray, Float t_max, bool any_hit, Hit* hit, Counters* ctr) {
    const SgSceneDesc* D = sc.d;
    hit->prim = -1; hit->inst = -1;
    if (D->n_nodes == 0) return false;
    uint32_t tn = D->n_top_nodes ? D->n_top_nodes : D->n_nodes, tp = D->n_top_primitives ? D->n_top_primitives : D->n_primitives;
    return bvh_intersect_range(sc, 0, tn, 0, tp, ray, t_max, any_hit, hit, ctr);
}

// ---- BVH build: aggregate.rs:207-468 ----------------------------------------
struct BuildPrim { uint32_t index; float bmin[3], bmax[3]; };
inline Float centroid_axis(const BuildPrim& p, int a) { return 0.5f * p.bmin[a] + p.bmax[a] * 0.5f; }  // aggregate.rs:490-492

struct BvhBuilder {
    std::vector<SgBvhNode> nodes;
    std::vector<uint32_t> order;
    // Iterative restatement of build_recursive + flatten_bvh: depth-first emission gives
    // the same linear order as flatten_bvh (:425-467) because the tree is emitted
    // node, left subtree, right subtree.
    uint32_t build(BuildPrim* prims, size_t n) {
        uint32_t my = (uint32_t)nodes.size();
        nodes.push_back(SgBvhNode());
        float bmin[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
        float bmax[3] = {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
        for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {       // :320-324 Bounds3::union via f32::min/max
            bmin[a] = fmin_(bmin[a], prims[i].bmin[a]);
            bmax[a] = fmax_(bmax[a], prims[i].bmax[a]);
        }
        Float dx = bmax[0] - bmin[0], dy = bmax[1] - bmin[1], dz = bmax[2] - bmin[2];
        Float sa = 2.0f * (dx * dy + dx * dz + dy * dz);                    // bounding_box.rs:394-397
        auto make_leaf = [&]() {
            SgBvhNode nd; std::memset(&nd, 0, sizeof nd);
            for (int a = 0; a < 3; ++a) { nd.bmin[a] = bmin[a]; nd.bmax[a] = bmax[a]; }
            nd.offset = (uint32_t)order.size(); nd.n_prims = (uint16_t)n; nd.axis = 0;
            for (size_t i = 0; i < n; ++i) order.push_back(prims[i].index);
            nodes[my] = nd;
        };
        if (sa == 0.0f || n == 1) { make_leaf(); return my; }              // :326-337
        float cmin[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
        float cmax[3] = {std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest(), std::numeric_limits<float>::lowest()};
        for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) {
            Float c = centroid_axis(prims[i], a);
            cmin[a] = fmin_(cmin[a], c); cmax[a] = fmax_(cmax[a], c);
        }
        Float ex = cmax[0] - cmin[0], ey = cmax[1] - cmin[1], ez = cmax[2] - cmin[2];
        int dim = (ex > ey && ex > ez) ? 0 : (ey > ez ? 1 : 2);            // bounding_box.rs:404-415
        if (cmax[dim] == cmin[dim]) { make_leaf(); return my; }            // :345-355
        Float pmid = (cmin[dim] + cmax[dim]) / 2.0f;                       // :360
        // itertools::partition (two-pointer swap partition), :361-364
        size_t split = 0;
        {
            size_t front = 0, back = n;
            for (;;) {
                if (front == back) break;
                if (!(centroid_axis(prims[front], dim) < pmid)) {
                    bool swapped = false;
                    while (back > front + 1) {
                        --back;
                        if (centroid_axis(prims[back], dim) < pmid) { std::swap(prims[front], prims[back]); swapped = true; break; }
                    }
                    if (!swapped) break;
                }
                ++front; ++split;
            }
        }
        if (split == 0 || split == n) {                                     // :366-374 pdqselect fallback
            split = n / 2;
            std::nth_element(prims, prims + split, prims + n, [dim](const BuildPrim& a, const BuildPrim& b) {
                return centroid_axis(a, dim) < centroid_axis(b, dim);
            });
        }
        build(prims, split);
        uint32_t second = build(prims + split, n - split);
        SgBvhNode nd; std::memset(&nd, 0, sizeof nd);
        const SgBvhNode& l = nodes[my + 1]; const SgBvhNode& r = nodes[second];   // init_interior :540-545
        for (int a = 0; a < 3; ++a) { nd.bmin[a] = fmin_(l.bmin[a], r.bmin[a]); nd.bmax[a] = fmax_(l.bmax[a], r.bmax[a]); }
        nd.offset = second; nd.n_prims = 0; nd.axis = (uint8_t)dim;
        nodes[my] = nd;
        return my;
    }
};

// ---- surface interaction -----------------------------------------------------
struct SurfaceInteraction {
    P3fi pi; V3 wo; V3 n; V2 uv;
    V3 dpdu, dpdv;
    V3 sn, sdpdu, sdpdv;     // shading.n / shading.dpdu / shading.dpdv
    V3 sdndu = {0, 0, 0}, sdndv = {0, 0, 0};                    // shading.dndu / shading.dndv
    Float dudx = 0, dudy = 0, dvdx = 0, dvdy = 0;               // compute_differentials interaction.rs:280-366
    V3 dpdx = {0, 0, 0}, dpdy = {0, 0, 0};
    int32_t material, light;
    inline V3 p() const { return p3fi_mid(pi); }
};

// triangle.rs:305-504 (dndu/dndv :451-498 feed bump mapping and specular ray differentials).
inline SurfaceInteraction interaction_from_intersection(const Scene& sc, uint32_t mesh_id, uint32_t tri, const TriHit& ti, V3 wo) {
    const SgMesh& m = sc.d->meshes[mesh_id];
    uint32_t v[3]; sc.tri_indices(mesh_id, tri, v);
    V3 p0 = sc.vertex(m, v[0]), p1 = sc.vertex(m, v[1]), p2 = sc.vertex(m, v[2]);
    V2 uv[3];
    if (!(m.flags & SG_MESH_HAS_UV)) { uv[0] = {0.0f, 0.0f}; uv[1] = {1.0f, 0.0f}; uv[2] = {1.0f, 1.0f}; }
    else { uv[0] = sc.uv(m, v[0]); uv[1] = sc.uv(m, v[1]); uv[2] = sc.uv(m, v[2]); }
    V2 duv02 = {uv[0].x - uv[2].x, uv[0].y - uv[2].y}, duv12 = {uv[1].x - uv[2].x, uv[1].y - uv[2].y};
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    Float determinant = difference_of_products(duv02.x, duv12.y, duv02.y, duv12.x);
    bool degenerate_uv = std::fabs(determinant) < 1e-9f;
    V3 dpdu = v3(0, 0, 0), dpdv = v3(0, 0, 0);
    if (!degenerate_uv) {
        Float inv_det = 1.0f / determinant;
        // difference_of_products_float_vec, math.rs:214-219 (unfused)
        auto dopv = [](Float a, V3 b, Float c, V3 d) { V3 cd = c * d; V3 diff = a * b - cd; V3 err = (-c) * d + cd; return diff + err; };
        dpdu = dopv(duv12.y, dp02, duv02.y, dp12) * inv_det;
        dpdv = dopv(duv02.x, dp12, duv12.x, dp02) * inv_det;
    }
    if (degenerate_uv || length_squared(cross(dpdu, dpdv)) == 0.0f) {
        V3 ng = cross(p2 - p0, p1 - p0);
        if (length_squared(ng) == 0.0f) {
            V3 v1 = p2 - p0, v2 = p1 - p0;
            ng = v3((Float)difference_of_products_d(v1.y, v2.z, v1.z, v2.y),
                    (Float)difference_of_products_d(v1.z, v2.x, v1.x, v2.z),
                    (Float)difference_of_products_d(v1.x, v2.y, v1.y, v2.x));
        }
        coordinate_system(normalize(ng), &dpdu, &dpdv);
    }
    V3 p_hit = ti.b0 * p0 + ti.b1 * p1 + ti.b2 * p2;
    V2 uv_hit = {ti.b0 * uv[0].x + ti.b1 * uv[1].x + ti.b2 * uv[2].x, ti.b0 * uv[0].y + ti.b1 * uv[1].y + ti.b2 * uv[2].y};
    bool flip = ((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    V3 p_abs_sum = vabs(ti.b0 * p0) + vabs(ti.b1 * p1) + vabs(ti.b2 * p2);
    V3 p_error = gamma_n(7) * p_abs_sum;
    SurfaceInteraction si;
    si.pi = p3fi_from_value_and_error(p_hit, p_error);
    si.uv = uv_hit; si.wo = wo; si.dpdu = dpdu; si.dpdv = dpdv;
    si.sdpdu = dpdu; si.sdpdv = dpdv;
    si.n = normalize(cross(dp02, dp12));                                  // :407-412
    if (flip) si.n = -si.n;
    si.sn = si.n;
    si.material = -1; si.light = -1;
    if (m.flags & (SG_MESH_HAS_N | SG_MESH_HAS_S)) {                      // :414-501
        V3 ns;
        if (!(m.flags & SG_MESH_HAS_N)) ns = si.n;
        else {
            V3 nn = ti.b0 * sc.normal(m, v[0]) + ti.b1 * sc.normal(m, v[1]) + ti.b2 * sc.normal(m, v[2]);
            ns = length_squared(nn) > 0.0f ? normalize(nn) : si.n;
        }
        V3 ss = si.dpdu;
        if (m.flags & SG_MESH_HAS_S) {
            const float* S = sc.d->s;
            auto sv = [&](uint32_t l) { const float* q = S + 3 * (size_t)(m.first_vertex + l); return v3(q[0], q[1], q[2]); };
            V3 s = ti.b0 * sv(v[0]) + ti.b1 * sv(v[1]) + ti.b2 * sv(v[2]);
            if (length_squared(s) != 0.0f) ss = s;
        }
        V3 ts = cross(ns, ss);
        if (length_squared(ts) > 0.0f) ss = cross(ts, ns);
        else coordinate_system(ns, &ss, &ts);
        // set_shading_geometry(ns, ss, ts, .., orientation_is_authoritative = true) interaction.rs:379-405
        si.sn = ns;
        si.n = face_forward(si.n, si.sn);
        si.sdpdu = ss; si.sdpdv = ts;
        if (m.flags & SG_MESH_HAS_N) {                                     // dndu, dndv :451-498
            V3 n0 = sc.normal(m, v[0]), n1 = sc.normal(m, v[1]), n2 = sc.normal(m, v[2]);
            if (degenerate_uv) {
                V3 dn = cross(n2 - n0, n1 - n0);
                if (length_squared(dn) != 0.0f) coordinate_system(dn, &si.sdndu, &si.sdndv);
            } else {
                Float inv_det = 1.0f / determinant;
                V3 dn1 = n0 - n2, dn2 = n1 - n2;
                auto dopv = [](Float a, V3 b, Float c, V3 d) { V3 cd = c * d; V3 diff = a * b - cd; V3 err = (-c) * d + cd; return diff + err; };
                si.sdndu = dopv(duv12.y, dn1, duv02.y, dn2) * inv_det;
                si.sdndv = dopv(duv02.x, dn2, duv12.x, dn1) * inv_det;
            }
        }
        while (length_squared(si.sdpdu) > 1e16f || length_squared(si.sdpdv) > 1e16f) { si.sdpdu = si.sdpdu / 1e8f; si.sdpdv = si.sdpdv / 1e8f; }
    }
    return si;
}


// Transform::apply(SurfaceInteraction) transform.rs:573-609 for an instanced hit.  The reference maps every vector and
// normal through `t = self.inverse()`: vectors by M^-1, normals by apply_normal_helper(t.m_inv = M) = M^T.  With
// SG_SCENE_FIX_INSTANCING vectors go through M and normals through (M^-1)^T (pbrt).  pi: forward Point3fi transform of an
// inexact point (:385-457).
inline void transform_interaction_m(const SgSceneDesc* D, const float* M, const float* Mi, SurfaceInteraction& si) {
    bool fix = (D->scene_flags & SG_SCENE_FIX_INSTANCING) != 0;
    auto vec = [&](V3 v) { const float* m = fix ? M : Mi; return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z); };
    auto nrm = [&](V3 n) { const float* m = fix ? Mi : M; return v3(m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z, m[2] * n.x + m[6] * n.y + m[10] * n.z); };
    V3 p = p3fi_mid(si.pi), e = p3fi_error(si.pi);
    Float x = p.x, y = p.y, z = p.z;
    Float xp = (M[0] * x + M[1] * y) + (M[2] * z + M[3]);
    Float yp = (M[4] * x + M[5] * y) + (M[6] * z + M[7]);
    Float zp = (M[8] * x + M[9] * y) + (M[10] * z + M[11]);
    V3 err;
    bool exact = si.pi.lo.x == si.pi.hi.x && si.pi.lo.y == si.pi.hi.y && si.pi.lo.z == si.pi.hi.z;
    auto row_err = [&](int r) {
        Float a = gamma_n(3) * (std::fabs(M[4 * r] * x) + std::fabs(M[4 * r + 1] * y) + std::fabs(M[4 * r + 2] * z) + std::fabs(M[4 * r + 3]));
        if (exact) return a;
        return (gamma_n(3) + 1.0f) * (std::fabs(M[4 * r]) * e.x + std::fabs(M[4 * r + 1]) * e.y + std::fabs(M[4 * r + 2]) * e.z) + a;
    };
    err = v3(row_err(0), row_err(1), row_err(2));
    si.pi = p3fi_from_value_and_error(v3(xp, yp, zp), err);
    V3 n = normalize(nrm(si.n));
    si.n = n;
    si.wo = normalize(vec(si.wo));
    si.dpdu = vec(si.dpdu); si.dpdv = vec(si.dpdv);
    si.sn = face_forward(normalize(nrm(si.sn)), n);
    si.sdpdu = vec(si.sdpdu); si.sdpdv = vec(si.sdpdv);
    si.sdndu = nrm(si.sdndu); si.sdndv = nrm(si.sdndv);
}
inline void transform_interaction(const SgSceneDesc* D, const SgInstance& I, SurfaceInteraction& si) {
    transform_interaction_m(D, I.render_from_primitive, I.primitive_from_render, si);
}

}  // namespace orc
#include "orc_sphere_surface.h"
