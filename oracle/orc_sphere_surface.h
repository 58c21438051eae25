// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h).
// Sphere surface interaction + bounds (second half of orc_sphere.h; needs SurfaceInteraction / transform_interaction_m).
#pragma once
namespace orc {

// Sphere::interaction_from_intersection sphere.rs:188-268 followed by render_from_object.apply(si) (transform.rs:573-609,
// shared with instancing: transform_interaction_m).  `wo` is the render-space -ray.d.
inline SurfaceInteraction sphere_interaction(const SgSceneDesc* D, const SgSphere& S, V3 p_hit, Float phi, V3 wo) {
    const Float u = phi / S.phi_max;
    const Float cos_theta = p_hit.z / S.radius;
    const Float theta = safe_asin(cos_theta);                           // sic: math.rs:272-274 `safe_acos` calls asin
    const Float v = (theta - S.theta_z_min) / (S.theta_z_max - S.theta_z_min);
    const Float z_radius = std::sqrt(p_hit.x * p_hit.x + p_hit.y * p_hit.y);
    const Float cos_phi = p_hit.x / z_radius, sin_phi = p_hit.y / z_radius;
    const V3 dpdu = v3(-S.phi_max * p_hit.y, S.phi_max * p_hit.x, 0.0f);
    const Float sin_theta = safe_sqrt(1.0f - cos_theta * cos_theta);
    const Float dth = S.theta_z_max - S.theta_z_min;
    const V3 dpdv = dth * v3(p_hit.z * cos_phi, p_hit.z * sin_phi, -S.radius * sin_theta);
    const V3 d2pduu = (-S.phi_max * S.phi_max) * v3(p_hit.x, p_hit.y, 0.0f);
    const V3 d2pduv = (dth * p_hit.z * S.phi_max) * v3(-sin_phi, cos_phi, 0.0f);
    const V3 d2pdvv = (-(dth * dth)) * v3(p_hit.x, p_hit.y, p_hit.z);
    const Float e1 = dot(dpdu, dpdu), f1 = dot(dpdu, dpdv), g1 = dot(dpdv, dpdv);
    const V3 n = normalize(cross(dpdu, dpdv));
    const Float e = dot(n, d2pduu), f = dot(n, d2pduv), g = dot(n, d2pdvv);
    const Float egf2 = difference_of_products(e1, g1, f1, f1);
    const Float inv = egf2 == 0.0f ? 0.0f : 1.0f / egf2;
    const V3 dndu = ((f * f1 - e * g1) * inv) * dpdu + ((e * f1 - f * e1) * inv) * dpdv;
    const V3 dndv = ((g * f1 - f * g1) * inv) * dpdu + ((f * f1 - g * e1) * inv) * dpdv;
    const V3 p_error = gamma_n(5) * vabs(p_hit);
    const bool flip = ((S.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((S.flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    // wo_object = object_from_render.apply(wo)
    const float* Mi = S.object_from_render;
    const V3 wo_obj = v3(Mi[0] * wo.x + Mi[1] * wo.y + Mi[2] * wo.z, Mi[4] * wo.x + Mi[5] * wo.y + Mi[6] * wo.z, Mi[8] * wo.x + Mi[9] * wo.y + Mi[10] * wo.z);
    SurfaceInteraction si;                                              // SurfaceInteraction::new interaction.rs:111-148
    si.pi = p3fi_from_value_and_error(p_hit, p_error);
    si.uv.x = u; si.uv.y = v; si.wo = wo_obj;
    si.dpdu = dpdu; si.dpdv = dpdv;
    si.n = flip ? -n : n;
    si.sn = si.n; si.sdpdu = dpdu; si.sdpdv = dpdv; si.sdndu = dndu; si.sdndv = dndv;
    si.material = -1; si.light = -1;
    transform_interaction_m(D, S.render_from_object, S.object_from_render, si);
    return si;
}

// Transform::apply(Bounds3f) transform.rs:557-571 of the object-space box (sphere.rs:273-279): the host-side bounds a
// BVH builder needs; exported for the host tests.
inline void sphere_bounds(const SgSphere& S, float bmin[3], float bmax[3]) {
    const float* M = S.render_from_object;
    const Float lo[3] = {-S.radius, -S.radius, S.z_min}, hi[3] = {S.radius, S.radius, S.z_max};
    bool first = true;
    for (int c = 0; c < 8; ++c) {
        const V3 p = v3((c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2]);
        const V3 q = v3(M[0] * p.x + M[1] * p.y + M[2] * p.z + M[3], M[4] * p.x + M[5] * p.y + M[6] * p.z + M[7], M[8] * p.x + M[9] * p.y + M[10] * p.z + M[11]);
        const Float qa[3] = {q.x, q.y, q.z};
        for (int a = 0; a < 3; ++a) {
            if (first) { bmin[a] = qa[a]; bmax[a] = qa[a]; }
            else { bmin[a] = fmin_(bmin[a], qa[a]); bmax[a] = fmax_(bmax[a], qa[a]); }
        }
        first = false;
    }
}

}  // namespace orc
