// Sphere shape on the device: Sphere::basic_intersect (shape/sphere.rs:95-186) in the reference's interval arithmetic
// (interval.rs, float.rs:92-130: every operation rounded outwards with next_float_up/down) on the ray transformed to
// object space as Point3fi / Vector3fi (transform.rs:385-513).  Pure IEEE f32 +,-,*,/,sqrt (this translation unit is
// built with -fmad=false), so t and p_obj are bit-identical to the CPU oracle; only atan2f (phi, used for the phimax
// clip and the u coordinate) comes from a different libm.
#pragma once
#include "sg_scene.cuh"

namespace sg {

static constexpr uint32_t kSphereBit = 0x40000000u;    // flag in tri_verts[3*i+2].w next to kLastInLeaf: the low bits index `spheres`

struct DSphere {
    float m[12];            // render_from_object rows 0..2
    float mi[12];           // object_from_render rows 0..2
    float radius, z_min, z_max, theta_z_min, theta_z_max, phi_max;
    uint32_t flags;         // SG_MESH_REVERSE_ORIENTATION | SG_MESH_SWAPS_HANDEDNESS
    uint32_t pad;
};

struct Ival { float lo, hi; };
SGD Ival iv(float v) { Ival r; r.lo = v; r.hi = v; return r; }
SGD Ival iv_new(float a, float b) { Ival r; r.lo = fminf(a, b); r.hi = fmaxf(a, b); return r; }                 // interval.rs:34-41
SGD Ival iv_ve(float v, float e) { Ival r; if (e == 0.0f) { r.lo = v; r.hi = v; } else { r.lo = next_down_n(v - e); r.hi = next_up_n(v + e); } return r; }                                 // :47-56
SGD float iv_mid(Ival a) { return (a.lo + a.hi) / 2.0f; }
SGD bool iv_in_range(Ival a, float v) { return v >= a.lo && v <= a.hi; }
SGD Ival iv_add(Ival a, Ival b) { Ival r; r.lo = next_down_n(a.lo + b.lo); r.hi = next_up_n(a.hi + b.hi); return r; }   // :353-355
SGD Ival iv_sub(Ival a, Ival b) { Ival r; r.lo = next_down_n(a.lo - b.lo); r.hi = next_up_n(a.hi - b.hi); return r; }   // :362-367 (sic: low-low, high-high)
SGD Ival iv_mul(Ival a, Ival b) {                                                                              // :374-392, fold(NAN, min/max)
    const float p0 = a.lo * b.lo, p1 = a.hi * b.lo, p2 = a.lo * b.hi, p3 = a.hi * b.hi;
    Ival r;
    r.lo = fminf(fminf(fminf(next_down_n(p0), next_down_n(p1)), next_down_n(p2)), next_down_n(p3));
    r.hi = fmaxf(fmaxf(fmaxf(next_up_n(p0), next_up_n(p1)), next_up_n(p2)), next_up_n(p3));
    return r;
}
SGD Ival iv_div(Ival a, Ival b) {                                                                              // :399-425
    Ival r;
    if (iv_in_range(b, 0.0f)) { r.lo = -INFINITY; r.hi = INFINITY; return r; }
    const float q0 = a.lo / b.lo, q1 = a.hi / b.lo, q2 = a.lo / b.hi, q3 = a.hi / b.hi;
    r.lo = fminf(fminf(fminf(next_down_n(q0), next_down_n(q1)), next_down_n(q2)), next_down_n(q3));
    r.hi = fmaxf(fmaxf(fmaxf(next_up_n(q0), next_up_n(q1)), next_up_n(q2)), next_up_n(q3));
    return r;
}
SGD Ival iv_scale(float f, Ival a) {                                                                           // Float * Interval :451-457
    if (f > 0.0f) return iv_new(next_down_n(f * a.lo), next_up_n(f * a.hi));
    return iv_new(next_down_n(f * a.hi), next_up_n(f * a.lo));
}
SGD Ival iv_sqr(Ival a) {                                                                                      // :99-117
    float alow = fabsf(a.lo), ahigh = fabsf(a.hi);
    if (alow > ahigh) { const float t = alow; alow = ahigh; ahigh = t; }
    Ival r;
    r.lo = iv_in_range(a, 0.0f) ? 0.0f : next_down_n(alow * alow);
    r.hi = next_up_n(ahigh * ahigh);
    return r;
}
SGD Ival iv_sqrt(Ival a) { Ival r; r.lo = next_down_n(sqrtf(a.lo)); r.hi = next_up_n(sqrtf(a.hi)); return r; }      // :498-505

struct V3i { Ival x, y, z; };
SGD float3 v3i_mid(const V3i& v) { return f3(iv_mid(v.x), iv_mid(v.y), iv_mid(v.z)); }

// Transform::apply(Point3fi) / apply(Vector3fi) of an EXACT value, affine matrix (transform.rs:385-457, 459-513)
SGD V3i sph_point_fi(const float* m, float3 p) {
    V3i r;
    r.x = iv_ve((m[0] * p.x + m[1] * p.y) + (m[2] * p.z + m[3]), gamma_n(3) * (fabsf(m[0] * p.x) + fabsf(m[1] * p.y) + fabsf(m[2] * p.z) + fabsf(m[3])));
    r.y = iv_ve((m[4] * p.x + m[5] * p.y) + (m[6] * p.z + m[7]), gamma_n(3) * (fabsf(m[4] * p.x) + fabsf(m[5] * p.y) + fabsf(m[6] * p.z) + fabsf(m[7])));
    r.z = iv_ve((m[8] * p.x + m[9] * p.y) + (m[10] * p.z + m[11]), gamma_n(3) * (fabsf(m[8] * p.x) + fabsf(m[9] * p.y) + fabsf(m[10] * p.z) + fabsf(m[11])));
    return r;
}
SGD V3i sph_vector_fi(const float* m, float3 v) {
    V3i r;
    r.x = iv_ve(m[0] * v.x + m[1] * v.y + m[2] * v.z, gamma_n(3) * (fabsf(m[0] * v.x) + fabsf(m[1] * v.y) + fabsf(m[2] * v.z)));
    r.y = iv_ve(m[4] * v.x + m[5] * v.y + m[6] * v.z, gamma_n(3) * (fabsf(m[4] * v.x) + fabsf(m[5] * v.y) + fabsf(m[6] * v.z)));
    r.z = iv_ve(m[8] * v.x + m[9] * v.y + m[10] * v.z, gamma_n(3) * (fabsf(m[8] * v.x) + fabsf(m[9] * v.y) + fabsf(m[10] * v.z)));
    return r;
}
SGD float sphere_phi(float3 p) { float ph = atan2f(p.y, p.x); if (ph < 0.0f) ph += 2.0f * kPi; return ph; }      // sphere.rs:138-141
SGD float3 sphere_hit_point(const DSphere& S, const V3i& oi, const V3i& di, Ival t) {                           // :127-137
    float3 p = v3i_mid(oi) + iv_mid(t) * v3i_mid(di);
    const float s = S.radius / len3(p);
    p = f3(p.x * s, p.y * s, p.z * s);
    if (p.x == 0.0f && p.y == 0.0f) p.x = 1e-5f * S.radius;
    return p;
}
SGD bool sphere_clipped(const DSphere& S, float3 p) {                                                           // :143-146
    return (S.z_min > -S.radius && p.z < S.z_min) || (S.z_max < S.radius && p.z > S.z_max) || sphere_phi(p) > S.phi_max;
}
// Sphere::basic_intersect sphere.rs:95-186; (o, d) is the render-space ray.  Out: QuadricIntersection{t_hit, p_obj}.
static __device__ __noinline__ bool sphere_basic_intersect(const DSphere& S, float3 o, float3 d, float t_max, float3& p_obj, float& t_hit) {
    const V3i oi = sph_point_fi(S.mi, o), di = sph_vector_fi(S.mi, d);
    const Ival a = iv_add(iv_add(iv_sqr(di.x), iv_sqr(di.y)), iv_sqr(di.z));
    const Ival b = iv_scale(2.0f, iv_add(iv_add(iv_mul(di.x, oi.x), iv_mul(di.y, oi.y)), iv_mul(di.z, oi.z)));
    const Ival rr = iv(S.radius);
    const Ival c = iv_sub(iv_add(iv_add(iv_sqr(oi.x), iv_sqr(oi.y)), iv_sqr(oi.z)), iv_sqr(rr));
    const Ival f = iv_div(b, iv_scale(2.0f, a));
    const Ival vx = iv_sub(oi.x, iv_mul(di.x, f)), vy = iv_sub(oi.y, iv_mul(di.y, f)), vz = iv_sub(oi.z, iv_mul(di.z, f));
    const Ival len = iv_sqrt(iv_add(iv_add(iv_mul(vx, vx), iv_mul(vy, vy)), iv_mul(vz, vz)));                   // length_fns.rs:6-21
    const Ival discrim = iv_mul(iv_mul(iv_scale(4.0f, a), iv_add(rr, len)), iv_sub(rr, len));
    if (discrim.lo < 0.0f) return false;
    const Ival root = iv_sqrt(discrim);
    const Ival q = iv_mid(b) < 0.0f ? iv_scale(-0.5f, iv_sub(b, root)) : iv_scale(-0.5f, iv_add(b, root));
    Ival t0 = iv_div(q, a), t1 = iv_div(c, q);
    if (t0.lo > t1.lo) { const Ival tmp = t0; t0 = t1; t1 = tmp; }
    if (t0.hi > t_max || t1.lo <= 0.0f) return false;
    Ival ts = t0;
    if (ts.lo <= 0.0f) { ts = t1; if (ts.hi > t_max) return false; }
    float3 p = sphere_hit_point(S, oi, di, ts);
    if (sphere_clipped(S, p)) {
        if (ts.lo == t1.lo && ts.hi == t1.hi) return false;
        if (t1.hi > t_max) return false;
        ts = t1;
        p = sphere_hit_point(S, oi, di, ts);
        if (sphere_clipped(S, p)) return false;
    }
    t_hit = iv_mid(ts); p_obj = p;
    return true;
}

}  // namespace sg
