// Sphere::interaction_from_intersection (shape/sphere.rs:188-268): the object-space SurfaceInteraction of a sphere hit.
// The caller maps it to render space with transform_interaction_m (Transform::apply(SurfaceInteraction), transform.rs:573-609).
#pragma once
#include "sg_sphere.cuh"
#include "sg_patch.cuh"
#include "sg_texture.cuh"

namespace sg {

template <bool TEX>
SGD Surf make_surface_sphere(const DSphere& S, float3 p_hit, SurfTex* x) {
    const float phi = sphere_phi(p_hit);
    const float u = phi / S.phi_max;
    const float cos_theta = p_hit.z / S.radius;
    const float theta = safe_asin(cos_theta);                                   // sic: math.rs:272-274 `safe_acos` calls asin
    const float v = (theta - S.theta_z_min) / (S.theta_z_max - S.theta_z_min);
    const float z_radius = sqrtf(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    const float cos_phi = p_hit.x / z_radius, sin_phi = p_hit.y / z_radius;
    const float3 dpdu = f3(-S.phi_max * p_hit.y, S.phi_max * p_hit.x, 0.0f);
    const float sin_theta = safe_sqrt(1.0f - cos_theta * cos_theta);
    const float dth = S.theta_z_max - S.theta_z_min;
    const float3 dpdv = dth * f3(p_hit.z * cos_phi, p_hit.z * sin_phi, -S.radius * sin_theta);
    Surf s;
    s.pi = p3fi_make(p_hit, gamma_n(5) * abs3(p_hit));
    const float3 n = normalize3(cross3(dpdu, dpdv));
    const bool flip = ((S.flags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((S.flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    s.n = flip ? -n : n;                                                        // SurfaceInteraction::new interaction.rs:111-148
    s.sn = s.n; s.sdpdu = dpdu; s.sdpdv = dpdv;
    if (TEX) {
        const float3 d2pduu = (-S.phi_max * S.phi_max) * f3(p_hit.x, p_hit.y, 0.0f);
        const float3 d2pduv = (dth * p_hit.z * S.phi_max) * f3(-sin_phi, cos_phi, 0.0f);
        const float3 d2pdvv = (-(dth * dth)) * f3(p_hit.x, p_hit.y, p_hit.z);
        const float e1 = dot3(dpdu, dpdu), f1 = dot3(dpdu, dpdv), g1 = dot3(dpdv, dpdv);
        const float e = dot3(n, d2pduu), f = dot3(n, d2pduv), g = dot3(n, d2pdvv);
        const float egf2 = dop(e1, g1, f1, f1);
        const float inv = egf2 == 0.0f ? 0.0f : 1.0f / egf2;
        const float3 dndu = ((f * f1 - e * g1) * inv) * dpdu + ((e * f1 - f * e1) * inv) * dpdv;
        const float3 dndv = ((g * f1 - f * g1) * inv) * dpdu + ((f * f1 - g * e1) * inv) * dpdv;
        surf_tex_store<TEX>(x, make_float2(u, v), dpdu, dpdv, dndu, dndv);
    }
    return s;
}

// BilinearPatch::interaction_from_intersection bilinear_patch.rs:238-425.  `rec` indexes patch_verts (4 float4: p00 p10 p01 p11;
// .w = mesh flags, mesh id, patch index in the mesh).
template <bool TEX>
__device__ __noinline__ Surf make_surface_patch(const DScene& sc, uint32_t rec, float u, float v, SurfTex* x) {
    const float4 a0 = __ldg(sc.patch_verts + 4 * (size_t)rec), a1 = __ldg(sc.patch_verts + 4 * (size_t)rec + 1),
                 a2 = __ldg(sc.patch_verts + 4 * (size_t)rec + 2), a3 = __ldg(sc.patch_verts + 4 * (size_t)rec + 3);
    const float3 p00 = f3(a0.x, a0.y, a0.z), p10 = f3(a1.x, a1.y, a1.z), p01 = f3(a2.x, a2.y, a2.z), p11 = f3(a3.x, a3.y, a3.z);
    const uint32_t flags = __float_as_uint(a0.w);
    const SgMesh m = sc.meshes[__float_as_uint(a1.w)];
    const uint32_t* ix = sc.indices + m.first_index + 4 * (size_t)__float_as_uint(a2.w);
    const float3 p = lerp3(u, lerp3(v, p00, p01), lerp3(v, p10, p11));
    float3 dpdu = lerp3(v, p10, p11) - lerp3(v, p00, p01);
    float3 dpdv = lerp3(u, p01, p11) - lerp3(u, p00, p10);
    float2 st = make_float2(u, v);
    float duds = 1.0f, dudt = 0.0f, dvds = 0.0f, dvdt = 1.0f;
    if (flags & SG_MESH_HAS_UV) {
        auto ld2 = [&](uint32_t k) { const float* q = sc.uv + 2 * ((size_t)m.first_vertex + __ldg(ix + k)); return make_float2(__ldg(q), __ldg(q + 1)); };
        auto lerp2 = [](float t, float2 a, float2 b) { return make_float2(a.x * (1.0f - t) + b.x * t, a.y * (1.0f - t) + b.y * t); };
        const float2 uv00 = ld2(0), uv10 = ld2(1), uv01 = ld2(2), uv11 = ld2(3);
        st = lerp2(u, lerp2(v, uv00, uv01), lerp2(v, uv10, uv11));
        const float2 c1 = lerp2(v, uv10, uv11), c0 = lerp2(v, uv00, uv01), e1_ = lerp2(u, uv01, uv11), e0_ = lerp2(u, uv00, uv10);
        const float2 dstdu = make_float2(c1.x - c0.x, c1.y - c0.y), dstdv = make_float2(e1_.x - e0_.x, e1_.y - e0_.y);
        duds = fabsf(dstdu.x) < 1e-8f ? 0.0f : 1.0f / dstdu.x;
        dvds = fabsf(dstdv.x) < 1e-8f ? 0.0f : 1.0f / dstdv.x;
        dudt = fabsf(dstdu.y) < 1e-8f ? 0.0f : 1.0f / dstdu.y;
        dvdt = fabsf(dstdv.y) < 1e-8f ? 0.0f : 1.0f / dstdv.y;
        const float3 dpds = dpdu * duds + dpdv * dvds;
        float3 dpdt = dpdu * dudt + dpdv * dvdt;
        const float3 cx = cross3(dpds, dpdt);
        if (!(cx.x == 0.0f && cx.y == 0.0f && cx.z == 0.0f)) {
            if (dot3(cross3(dpdu, dpdv), cross3(dpds, dpdt)) < 0.0f) dpdt = -dpdt;
            dpdu = dpds; dpdv = dpdt;
        }
    }
    const float3 d2pduv = (p00 - p01) + (p11 - p10), zero = f3(0.0f, 0.0f, 0.0f);
    const float e1 = dot3(dpdu, dpdu), f1 = dot3(dpdu, dpdv), g1 = dot3(dpdv, dpdv);
    const float3 n = normalize3(cross3(dpdu, dpdv));
    const float e2 = dot3(n, zero), f2 = dot3(n, d2pduv), g2 = dot3(n, zero);
    const float egf2 = dop(e1, g1, f1, f1);
    const float inv = egf2 != 0.0f ? 1.0f / egf2 : 0.0f;
    float3 dndu = ((f1 * f2 - e2 * g1) * inv) * dpdu + ((e2 * f1 - f2 * e1) * inv) * dpdv;
    float3 dndv = ((g2 * f1 - f2 * g1) * inv) * dpdu + ((f2 * f1 - g2 * e1) * inv) * dpdv;
    { const float3 dnds = dndu * duds + dndv * dvds, dndt = dndu * dudt + dndv * dvdt; dndu = dnds; dndv = dndt; }
    Surf s;
    s.pi = p3fi_make(p, gamma_n(6) * (abs3(p00) + abs3(p01) + abs3(p10) + abs3(p11)));
    const bool flip = ((flags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    s.n = flip ? -n : n;
    s.sn = s.n; s.sdpdu = dpdu; s.sdpdv = dpdv;
    float3 sdndu = dndu, sdndv = dndv;
    if (flags & SG_MESH_HAS_N) {
        const float3 n00 = ldv3(sc.n, (size_t)m.first_vertex + __ldg(ix)), n10 = ldv3(sc.n, (size_t)m.first_vertex + __ldg(ix + 1)),
                     n01 = ldv3(sc.n, (size_t)m.first_vertex + __ldg(ix + 2)), n11 = ldv3(sc.n, (size_t)m.first_vertex + __ldg(ix + 3));
        float3 ns = lerp3(u, lerp3(v, n00, n01), lerp3(v, n10, n11));
        if (len2(ns) > 0.0f) {
            ns = normalize3(ns);
            const float3 a = lerp3(v, n10, n11) - lerp3(v, n00, n01), b = lerp3(u, n01, n11) - lerp3(u, n00, n10);
            sdndu = a * duds + b * dvds; sdndv = a * dudt + b * dvdt;
            float r[9]; rotate_from_to(s.n, ns, r);
            s.sn = ns;                                                          // set_shading_geometry(.., true) interaction.rs:379-405
            s.n = faceforward3(s.n, s.sn);
            s.sdpdu = mul3(r, dpdu); s.sdpdv = mul3(r, dpdv);
            while (len2(s.sdpdu) > 1e16f || len2(s.sdpdv) > 1e16f) { s.sdpdu = s.sdpdu / 1e8f; s.sdpdv = s.sdpdv / 1e8f; }
        }
    }
    if (TEX) surf_tex_store<TEX>(x, st, dpdu, dpdv, sdndu, sdndv);
    return s;
}

// ---- Sphere as an emitter: sphere.rs:295-457 ----
SGD float sphere_area(const DSphere& S) { return S.phi_max * S.radius * (S.z_max - S.z_min); }
SGD float3 sphere_center(const DSphere& S) { return f3(S.m[3], S.m[7], S.m[11]); }                           // render_from_object.apply(Point3f::ZERO)
// Sphere::sample_with_context :339-420 (inside: Sphere::sample :299-333 + area -> solid angle; outside: cone sampling)
static __device__ __noinline__ bool sphere_sample_with_context(const DSphere& S, const LightCtx& ctx, float2 u, P3fi& out_pi, float3& out_n, float& out_pdf) {
    const float3 pc = sphere_center(S), cp = p3fi_mid(ctx.pi);
    const float3 p_origin = offset_ray_origin(ctx.pi, ctx.n, pc - cp);
    if (dist2(p_origin, pc) <= sqr(S.radius)) {
        const float z = 1.0f - 2.0f * u.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * kPi * u.y;            // sample_uniform_sphere sampling.rs:280-289
        float3 p_obj = f3(0.0f, 0.0f, 0.0f) + f3(r * cosf(phi), r * sinf(phi), z) * S.radius;
        const float sc_ = S.radius / len3(p_obj);
        p_obj = f3(p_obj.x * sc_, p_obj.y * sc_, p_obj.z * sc_);
        const float* Mi = S.mi; const float* M = S.m;
        float3 n = normalize3(f3(Mi[0] * p_obj.x + Mi[4] * p_obj.y + Mi[8] * p_obj.z, Mi[1] * p_obj.x + Mi[5] * p_obj.y + Mi[9] * p_obj.z,
                                 Mi[2] * p_obj.x + Mi[6] * p_obj.y + Mi[10] * p_obj.z));
        if (S.flags & SG_MESH_REVERSE_ORIENTATION) n = n * -1.0f;
        const P3fi pin = p3fi_make(p_obj, gamma_n(5) * abs3(p_obj));
        const float3 pm = p3fi_mid(pin), pe = p3fi_err(pin);
        const bool exact = pin.lo.x == pin.hi.x && pin.lo.y == pin.hi.y && pin.lo.z == pin.hi.z;              // is_exact: zero width
        float qa[3], ea[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            qa[k] = (M[4 * k] * pm.x + M[4 * k + 1] * pm.y) + (M[4 * k + 2] * pm.z + M[4 * k + 3]);
            const float a = gamma_n(3) * (fabsf(M[4 * k] * pm.x) + fabsf(M[4 * k + 1] * pm.y) + fabsf(M[4 * k + 2] * pm.z) + fabsf(M[4 * k + 3]));
            ea[k] = exact ? a : (gamma_n(3) + 1.0f) * (fabsf(M[4 * k]) * pe.x + fabsf(M[4 * k + 1]) * pe.y + fabsf(M[4 * k + 2]) * pe.z) + a;
        }
        out_pi = p3fi_make(f3(qa[0], qa[1], qa[2]), f3(ea[0], ea[1], ea[2])); out_n = n;
        float pdf = 1.0f / sphere_area(S);
        const float3 sp = p3fi_mid(out_pi);
        float3 wi = sp - cp;
        if (len2(wi) == 0.0f) return false;
        wi = normalize3(wi);
        pdf /= absdot3(n, -wi) / dist2(cp, sp);
        if (isinf(pdf)) return false;
        out_pdf = pdf;
        return true;
    }
    const float sin_theta_max = S.radius / sqrtf(dist2(cp, pc));
    const float sin2_theta_max = sqr(sin_theta_max);
    const float cos_theta_max = safe_sqrt(1.0f - sin2_theta_max);
    float one_minus_cos_theta_max = 1.0f - cos_theta_max;
    float cos_theta = (cos_theta_max - 1.0f) * u.x + 1.0f;
    float sin2_theta = 1.0f - sqr(cos_theta);
    if (sin2_theta_max < 0.00068523f) {
        sin2_theta = sin2_theta_max * u.x;
        cos_theta = sqrtf(1.0f - sin2_theta);
        one_minus_cos_theta_max = sin2_theta_max / 2.0f;
    }
    const float cos_alpha = sin2_theta / sin_theta_max + cos_theta * safe_sqrt(1.0f - sin2_theta / sqr(sin_theta_max));
    const float sin_alpha = safe_sqrt(1.0f - sqr(cos_alpha));
    const float phi = u.y * 2.0f * kPi;
    const float3 w = f3(clampf(sin_alpha, -1.0f, 1.0f) * cosf(phi), clampf(sin_alpha, -1.0f, 1.0f) * sinf(phi), clampf(cos_alpha, -1.0f, 1.0f));
    const float3 fz = normalize3(pc - cp); float3 fx, fy; coord_system(fz, fx, fy);                             // Frame::from_z
    const float3 mw = -w;
    float3 n = mw.x * fx + mw.y * fy + mw.z * fz;
    if (S.flags & SG_MESH_REVERSE_ORIENTATION) n = n * -1.0f;                                                    // sic: before the point is placed
    const float3 p = pc + f3(n.x, n.y, n.z) * S.radius;
    out_pi = p3fi_make(p, gamma_n(5) * abs3(p)); out_n = n;
    out_pdf = 1.0f / (2.0f * kPi * one_minus_cos_theta_max);
    return true;
}
// Sphere::pdf_with_context :422-456
static __device__ __noinline__ float sphere_pdf_with_context(const DScene& sc, const DSphere& S, const LightCtx& ctx, float3 wi) {
    const float3 pc = sphere_center(S), cp = p3fi_mid(ctx.pi);
    const float3 p_origin = offset_ray_origin(ctx.pi, ctx.n, pc - cp);
    if (dist2(p_origin, pc) <= S.radius * S.radius) {
        const float3 o = offset_ray_origin(ctx.pi, ctx.n, wi);
        float3 p_obj; float t;
        if (!sphere_basic_intersect(S, o, wi, INFINITY, p_obj, t)) return 0.0f;
        Surf s = make_surface_sphere<false>(S, p_obj, nullptr);
        float3 wo_si;
        transform_interaction<false>(sc, S.m, S.mi, wi, s, nullptr, wo_si);
        const float pdf = (1.0f / sphere_area(S)) / absdot3(s.n, -wi) / dist2(cp, p3fi_mid(s.pi));               // sic: (a / b) / c
        if (isinf(pdf)) return 0.0f;
        return pdf;
    }
    const float sin2_theta_max = S.radius * S.radius / dist2(cp, pc);
    const float cos_theta_max = safe_sqrt(1.0f - sin2_theta_max);
    float one_minus_cos_theta_max = 1.0f - cos_theta_max;
    if (sin2_theta_max < 0.00068523f) one_minus_cos_theta_max = sin2_theta_max / 2.0f;
    return 1.0f / (2.90f * kPi * one_minus_cos_theta_max);                                                      // sic: 2.90
}

}  // namespace sg
