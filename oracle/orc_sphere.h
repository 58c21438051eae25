// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h).
//
// Sphere shape: src/shape/sphere.rs restated, with the interval arithmetic it runs in (src/interval.rs,
// src/float.rs:92-130) and the Point3fi / Vector3fi transforms (src/transform.rs:385-513).
#pragma once
#include "orc_math.h"
#include "../include/shimmer_gpu.h"

namespace orc {

// ---- interval.rs -------------------------------------------------------------------------------------------------
struct Ival { Float lo, hi; };
inline Ival iv(Float v) { Ival r = {v, v}; return r; }                                    // from_val :43-45
inline Ival iv_new(Float a, Float b) { Ival r = {fmin_(a, b), fmax_(a, b)}; return r; }   // new :34-41
inline Ival iv_from_value_and_error(Float v, Float e) { Ival r; interval_from_value_and_error(v, e, &r.lo, &r.hi); return r; }   // :47-56
inline Float iv_mid(Ival a) { return (a.lo + a.hi) / 2.0f; }                              // :66-68
inline bool iv_in_range(Ival a, Float v) { return v >= a.lo && v <= a.hi; }               // :87-89
inline Ival iv_add(Ival a, Ival b) { Ival r = {next_float_down(a.lo + b.lo), next_float_up(a.hi + b.hi)}; return r; }   // :353-355
inline Ival iv_sub(Ival a, Ival b) { Ival r = {next_float_down(a.lo - b.lo), next_float_up(a.hi - b.hi)}; return r; }   // :362-367 (sic: low-low, high-high)
inline Float fold_min4(const Float v[4]) { Float a = NAN; for (int i = 0; i < 4; ++i) a = fmin_(a, v[i]); return a; }   // fold(NAN, min)
inline Float fold_max4(const Float v[4]) { Float a = NAN; for (int i = 0; i < 4; ++i) a = fmax_(a, v[i]); return a; }
inline Ival iv_mul(Ival a, Ival b) {                                                      // :374-392
    const Float p[4] = {a.lo * b.lo, a.hi * b.lo, a.lo * b.hi, a.hi * b.hi};
    Float lp[4], hp[4];
    for (int i = 0; i < 4; ++i) { lp[i] = next_float_down(p[i]); hp[i] = next_float_up(p[i]); }
    Ival r = {fold_min4(lp), fold_max4(hp)}; return r;
}
inline Ival iv_div(Ival a, Ival b) {                                                      // :399-425
    if (iv_in_range(b, 0.0f)) { Ival r = {-F_INF, F_INF}; return r; }
    const Float q[4] = {a.lo / b.lo, a.hi / b.lo, a.lo / b.hi, a.hi / b.hi};
    Float lq[4], hq[4];
    for (int i = 0; i < 4; ++i) { lq[i] = next_float_down(q[i]); hq[i] = next_float_up(q[i]); }
    Ival r = {fold_min4(lq), fold_max4(hq)}; return r;
}
inline Ival iv_scale(Float f, Ival a) {                                                   // Float * Interval :451-457
    if (f > 0.0f) return iv_new(next_float_down(f * a.lo), next_float_up(f * a.hi));
    return iv_new(next_float_down(f * a.hi), next_float_up(f * a.lo));
}
inline Ival iv_sqr(Ival a) {                                                              // :99-117
    Float alow = std::fabs(a.lo), ahigh = std::fabs(a.hi);
    if (alow > ahigh) std::swap(alow, ahigh);
    if (iv_in_range(a, 0.0f)) { Ival r = {0.0f, next_float_up(ahigh * ahigh)}; return r; }
    Ival r = {next_float_down(alow * alow), next_float_up(ahigh * ahigh)}; return r;
}
inline Ival iv_sqrt(Ival a) { Ival r = {next_float_down(std::sqrt(a.lo)), next_float_up(std::sqrt(a.hi))}; return r; }   // :498-505
inline bool iv_eq(Ival a, Ival b) { return a.lo == b.lo && a.hi == b.hi; }                // derived PartialEq

struct V3i { Ival x, y, z; };
inline V3 v3i_mid(const V3i& v) { return v3(iv_mid(v.x), iv_mid(v.y), iv_mid(v.z)); }

// Transform::apply(Point3fi) transform.rs:385-457 for an EXACT input point and an affine matrix (wp == 1)
inline V3i sph_xform_point_fi(const float* m, V3 p) {
    const Float x = p.x, y = p.y, z = p.z;
    const Float xp = (m[0] * x + m[1] * y) + (m[2] * z + m[3]);
    const Float yp = (m[4] * x + m[5] * y) + (m[6] * z + m[7]);
    const Float zp = (m[8] * x + m[9] * y) + (m[10] * z + m[11]);
    const Float ex = gamma_n(3) * (std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z) + std::fabs(m[3]));
    const Float ey = gamma_n(3) * (std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z) + std::fabs(m[7]));
    const Float ez = gamma_n(3) * (std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z) + std::fabs(m[11]));
    V3i r = {iv_from_value_and_error(xp, ex), iv_from_value_and_error(yp, ey), iv_from_value_and_error(zp, ez)}; return r;
}
// Transform::apply(Vector3fi) transform.rs:459-513 for an EXACT input vector
inline V3i sph_xform_vector_fi(const float* m, V3 v) {
    const Float x = v.x, y = v.y, z = v.z;
    const Float ex = gamma_n(3) * (std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z));
    const Float ey = gamma_n(3) * (std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z));
    const Float ez = gamma_n(3) * (std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z));
    const Float xp = m[0] * x + m[1] * y + m[2] * z, yp = m[4] * x + m[5] * y + m[6] * z, zp = m[8] * x + m[9] * y + m[10] * z;
    V3i r = {iv_from_value_and_error(xp, ex), iv_from_value_and_error(yp, ey), iv_from_value_and_error(zp, ez)}; return r;
}

// QuadricIntersection (shape.rs): t_hit, p_obj, phi
struct QuadricHit { Float t, phi; V3 p_obj; };

// sphere.rs:127-141 / :152-166: hit point on the sphere and its phi
inline void sphere_hit_point(const SgSphere& S, const V3i& oi, const V3i& di, Ival t, V3* p_hit, Float* phi) {
    V3 p = v3i_mid(oi) + iv_mid(t) * v3i_mid(di);
    const Float s = S.radius / length(p);                               // p_hit.distance(Point3f::ZERO)
    p = v3(p.x * s, p.y * s, p.z * s);
    if (p.x == 0.0f && p.y == 0.0f) p.x = 1e-5f * S.radius;
    Float ph = std::atan2(p.y, p.x);
    if (ph < 0.0f) ph += 2.0f * PI_F;
    *p_hit = p; *phi = ph;
}
inline bool sphere_clipped(const SgSphere& S, V3 p, Float phi) {        // :143-146
    return (S.z_min > -S.radius && p.z < S.z_min) || (S.z_max < S.radius && p.z > S.z_max) || phi > S.phi_max;
}
// Sphere::basic_intersect sphere.rs:95-186
inline bool sphere_basic_intersect(const SgSphere& S, const Ray& ray, Float t_max, QuadricHit* out) {
    const V3i oi = sph_xform_point_fi(S.object_from_render, ray.o);
    const V3i di = sph_xform_vector_fi(S.object_from_render, ray.d);
    const Ival a = iv_add(iv_add(iv_sqr(di.x), iv_sqr(di.y)), iv_sqr(di.z));
    const Ival b = iv_scale(2.0f, iv_add(iv_add(iv_mul(di.x, oi.x), iv_mul(di.y, oi.y)), iv_mul(di.z, oi.z)));
    const Ival rr = iv(S.radius);
    const Ival c = iv_sub(iv_add(iv_add(iv_sqr(oi.x), iv_sqr(oi.y)), iv_sqr(oi.z)), iv_sqr(rr));
    // v = oi - b / (2 a) * di ; length via x*x + y*y + z*z (length_fns.rs:6-13: general products, not sqr)
    const Ival f = iv_div(b, iv_scale(2.0f, a));
    const V3i v = {iv_sub(oi.x, iv_mul(di.x, f)), iv_sub(oi.y, iv_mul(di.y, f)), iv_sub(oi.z, iv_mul(di.z, f))};
    const Ival len = iv_sqrt(iv_add(iv_add(iv_mul(v.x, v.x), iv_mul(v.y, v.y)), iv_mul(v.z, v.z)));
    const Ival discrim = iv_mul(iv_mul(iv_scale(4.0f, a), iv_add(rr, len)), iv_sub(rr, len));
    if (discrim.lo < 0.0f) return false;
    const Ival root = iv_sqrt(discrim);
    const Ival q = iv_mid(b) < 0.0f ? iv_scale(-0.5f, iv_sub(b, root)) : iv_scale(-0.5f, iv_add(b, root));
    Ival t0 = iv_div(q, a), t1 = iv_div(c, q);
    if (t0.lo > t1.lo) std::swap(t0, t1);
    if (t0.hi > t_max || t1.lo <= 0.0f) return false;
    Ival t_shape = t0;
    if (t_shape.lo <= 0.0f) { t_shape = t1; if (t_shape.hi > t_max) return false; }
    V3 p_hit; Float phi;
    sphere_hit_point(S, oi, di, t_shape, &p_hit, &phi);
    if (sphere_clipped(S, p_hit, phi)) {
        if (iv_eq(t_shape, t1)) return false;
        if (t1.hi > t_max) return false;
        t_shape = t1;
        sphere_hit_point(S, oi, di, t_shape, &p_hit, &phi);
        if (sphere_clipped(S, p_hit, phi)) return false;
    }
    out->t = iv_mid(t_shape); out->p_obj = p_hit; out->phi = phi;
    return true;
}

}  // namespace orc
