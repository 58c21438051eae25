// Sphere::interaction_from_intersection (shape/sphere.rs:188-268): the object-space SurfaceInteraction of a sphere hit.
// The caller maps it to render space with transform_interaction_m (Transform::apply(SurfaceInteraction), transform.rs:573-609).
#pragma once
#include "sg_sphere.cuh"
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

}  // namespace sg
