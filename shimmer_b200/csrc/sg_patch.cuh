// Bilinear patch on the device: BilinearPatch::intersect_blp (shape/bilinear_patch.rs:144-236), quadratic (math.rs:377-410),
// SquareMatrix<3>::determinant (square_matrix.rs:281-292).  Pure f32 arithmetic with the reference's explicit FMAs only.
#pragma once
#include "sg_scene.cuh"

namespace sg {

static constexpr uint32_t kPatchBit = 0x20000000u;     // flag in tri_verts[3*i+2].w: the low bits index `patch_verts` (4 float4 per patch)

SGD bool quadratic(float a, float b, float c, float& t0, float& t1) {
    if (a == 0.0f) { if (b == 0.0f) return false; t0 = -c / b; t1 = -c / b; return true; }
    const float discrim = dop(b, b, 4.0f * a, c);
    if (discrim < 0.0f) return false;
    const float root = sqrtf(discrim);
    const float q = -0.5f * (b + copysignf(root, b));
    float x0 = q / a, x1 = c / q;
    if (x0 > x1) { const float t = x0; x0 = x1; x1 = t; }
    t0 = x0; t1 = x1; return true;
}
SGD float det3_cols(float3 c0, float3 c1, float3 c2) {           // matrix rows [c0.x c1.x c2.x], [c0.y ...], [c0.z ...]
    const float minor12 = dop(c1.y, c2.z, c2.y, c1.z);
    const float minor02 = dop(c0.y, c2.z, c2.y, c0.z);
    const float minor01 = dop(c0.y, c1.z, c1.y, c0.z);
    return fmaf(c2.x, minor01, dop(c0.x, minor12, c1.x, minor02));
}
SGD float3 lerp3(float t, float3 a, float3 b) { return a * (1.0f - t) + b * t; }                               // math.rs:246-252
SGD float max_abs_comp(float3 v) { return fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z))); }

static __device__ __noinline__ bool intersect_blp(float3 ro, float3 rd, float t_max, float3 p00, float3 p10, float3 p01, float3 p11, float& u_out, float& v_out, float& t_out) {
    const float a = dot3(cross3(p10 - p00, p01 - p11), rd);
    const float c = dot3(cross3(p00 - ro, rd), p01 - p00);
    const float b = dot3(cross3(p10 - ro, rd), p11 - p10) - (a + c);
    float u1, u2;
    if (!quadratic(a, b, c, u1, u2)) return false;
    const float eps = gamma_n(10) * (max_abs_comp(ro) + max_abs_comp(rd) + max_abs_comp(p00) + max_abs_comp(p10) + max_abs_comp(p01) + max_abs_comp(p11));
    float t = t_max, u = 0.0f, v = 0.0f;
    if (0.0f <= u1 && u1 <= 1.0f) {
        const float3 uo = lerp3(u1, p00, p10);
        const float3 ud = lerp3(u1, p01, p11) - uo;
        const float3 deltao = uo - ro;
        const float3 perp = cross3(rd, ud);
        const float p2 = len2(perp);
        const float v1 = det3_cols(deltao, rd, perp);
        const float t1 = det3_cols(deltao, ud, perp);
        if (t1 > p2 * eps && 0.0f <= v1 && v1 <= p2) { u = u1; v = v1 / p2; t = t1 / p2; }
    }
    if (0.0f <= u2 && u2 <= 1.0f && u2 != u1) {
        const float3 uo = lerp3(u2, p00, p10);
        const float3 ud = lerp3(u2, p01, p11) - uo;
        const float3 deltao = uo - ro;
        const float3 perp = cross3(rd, ud);
        const float p2 = len2(perp);
        const float v2 = det3_cols(deltao, rd, perp);
        float t2 = det3_cols(deltao, ud, perp);
        t2 /= p2;
        if (0.0f <= v2 && v2 <= p2 && t > t2 && t2 > eps) { t = t2; u = u2; v = v2 / p2; }
    }
    if (t >= t_max) return false;
    u_out = u; v_out = v; t_out = t;
    return true;
}

}  // namespace sg
