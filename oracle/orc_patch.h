// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h).
//
// Bilinear patch: BilinearPatch::intersect_blp (src/shape/bilinear_patch.rs:144-236), quadratic (math.rs:377-410),
// SquareMatrix<3>::determinant (square_matrix.rs:281-292).
#pragma once
#include "orc_math.h"

namespace orc {

inline bool quadratic(Float a, Float b, Float c, Float* t0, Float* t1) {            // math.rs:377-410
    if (a == 0.0f) { if (b == 0.0f) return false; *t0 = -c / b; *t1 = -c / b; return true; }
    const Float discrim = difference_of_products(b, b, 4.0f * a, c);
    if (discrim < 0.0f) return false;
    const Float root = std::sqrt(discrim);
    const Float q = -0.5f * (b + std::copysign(root, b));
    Float x0 = q / a, x1 = c / q;
    if (x0 > x1) std::swap(x0, x1);
    *t0 = x0; *t1 = x1; return true;
}
// determinant of the 3x3 matrix whose COLUMNS are c0, c1, c2 (rows [c0.x c1.x c2.x] ...), square_matrix.rs:281-292
inline Float det3_cols(V3 c0, V3 c1, V3 c2) {
    const Float m00 = c0.x, m01 = c1.x, m02 = c2.x, m10 = c0.y, m11 = c1.y, m12 = c2.y, m20 = c0.z, m21 = c1.z, m22 = c2.z;
    const Float minor12 = difference_of_products(m11, m22, m12, m21);
    const Float minor02 = difference_of_products(m10, m22, m12, m20);
    const Float minor01 = difference_of_products(m10, m21, m11, m20);
    return std::fma(m02, minor01, difference_of_products(m00, minor12, m01, minor02));
}
inline V3 lerp3(Float t, V3 a, V3 b) { return a * (1.0f - t) + b * t; }             // math.rs:246-252
inline Float max_abs_comp(V3 v) { return fmax_(std::fabs(v.x), fmax_(std::fabs(v.y), std::fabs(v.z))); }   // tuple.rs:179-181 on .abs()

// bilinear_patch.rs:144-236.  Out: (u, v, t).
inline bool intersect_blp(V3 ro, V3 rd, Float t_max, V3 p00, V3 p10, V3 p01, V3 p11, Float* u_out, Float* v_out, Float* t_out) {
    const Float a = dot(cross(p10 - p00, p01 - p11), rd);
    const Float c = dot(cross(p00 - ro, rd), p01 - p00);
    const Float b = dot(cross(p10 - ro, rd), p11 - p10) - (a + c);
    Float u1, u2;
    if (!quadratic(a, b, c, &u1, &u2)) return false;
    const Float eps = gamma_n(10) * (max_abs_comp(ro) + max_abs_comp(rd) + max_abs_comp(p00) + max_abs_comp(p10) + max_abs_comp(p01) + max_abs_comp(p11));
    Float t = t_max, u = 0.0f, v = 0.0f;
    if (0.0f <= u1 && u1 <= 1.0f) {
        const V3 uo = lerp3(u1, p00, p10);
        const V3 ud = lerp3(u1, p01, p11) - uo;
        const V3 deltao = uo - ro;
        const V3 perp = cross(rd, ud);
        const Float p2 = length_squared(perp);
        const Float v1 = det3_cols(deltao, rd, perp);
        const Float t1 = det3_cols(deltao, ud, perp);
        if (t1 > p2 * eps && 0.0f <= v1 && v1 <= p2) { u = u1; v = v1 / p2; t = t1 / p2; }
    }
    if (0.0f <= u2 && u2 <= 1.0f && u2 != u1) {
        const V3 uo = lerp3(u2, p00, p10);
        const V3 ud = lerp3(u2, p01, p11) - uo;
        const V3 deltao = uo - ro;
        const V3 perp = cross(rd, ud);
        const Float p2 = length_squared(perp);
        const Float v2 = det3_cols(deltao, rd, perp);
        Float t2 = det3_cols(deltao, ud, perp);
        t2 /= p2;
        if (0.0f <= v2 && v2 <= p2 && t > t2 && t2 > eps) { t = t2; u = u2; v = v2 / p2; }
    }
    if (t >= t_max) return false;
    *u_out = u; *v_out = v; *t_out = t;
    return true;
}

}  // namespace orc
