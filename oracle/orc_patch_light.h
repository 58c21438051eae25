// ORACLE -- TEST INFRASTRUCTURE ONLY.  BilinearPatch as an emitter: is_rectangle / area (bilinear_patch.rs:40-143), sample (:521-600),
// pdf (:602-636), sample_with_context (:638-737), pdf_with_context (:739-783); sample_spherical_rectangle / invert_spherical_rectangle_sample
// (sampling.rs:501-579, 643-787), invert_bilinear / spherical_quad_area (vecmath/mod.rs:70-140).  Written as the reference has it,
// including Sample's `pu0 = lerp(uv[0], p00, p10)`, `pu1 = lerp(uv[1], p10, p11)` (pbrt interpolates both along v).
// Included from orc_shading.h in the lights section (needs ShapeSample, LightSampleContext, patch_interaction).
#pragma once

namespace orc {

inline Float spherical_quad_area(V3 a, V3 b, V3 c, V3 d) {                       // vecmath/mod.rs:118-140
    V3 axb = cross(a, b), bxc = cross(b, c), cxd = cross(c, d), dxa = cross(d, a);
    if (length_squared(axb) == 0.0f || length_squared(bxc) == 0.0f || length_squared(cxd) == 0.0f || length_squared(dxa) == 0.0f) return 0.0f;
    axb = normalize(axb); bxc = normalize(bxc); cxd = normalize(cxd); dxa = normalize(dxa);
    const Float alpha = angle_between(dxa, -axb), beta = angle_between(axb, -bxc), gamma = angle_between(bxc, -cxd), delta = angle_between(cxd, -dxa);
    return std::fabs(alpha + beta + gamma + delta - 2.0f * PI_F);
}
inline Float cross2d(V2 a, V2 b) { return difference_of_products(a.x, b.y, a.y, b.x); }
inline V2 invert_bilinear(V2 p, const V2 vert[4]) {                               // vecmath/mod.rs:70-116
    const V2 a = vert[0], b = vert[1], c = vert[3], d = vert[2];
    const V2 e = {b.x - a.x, b.y - a.y}, f = {d.x - a.x, d.y - a.y}, g = {(a.x - b.x) + (c.x - d.x), (a.y - b.y) + (c.y - d.y)}, h = {p.x - a.x, p.y - a.y};
    const Float k2 = cross2d(g, f), k1 = cross2d(e, f) + cross2d(h, g), k0 = cross2d(h, e);
    V2 r;
    if (std::fabs(k2) < 0.001f) {
        if (std::fabs(e.x * k1 - g.x * k0) < 1e-5f) { r.x = (h.y * k1 + f.y * k0) / (e.y * k1 - g.y * k0); r.y = -k0 / k1; return r; }
        r.x = (h.x * k1 + f.x * k0) / (e.x * k1 - g.x * k0); r.y = -k0 / k1; return r;
    }
    Float v0, v1;
    if (!quadratic(k2, k1, k0, &v0, &v1)) { r.x = 0.0f; r.y = 0.0f; return r; }
    const Float u = (h.x - f.x * v0) / (e.x + g.x * v0);
    if (u < 0.0f || u > 1.0f || v0 < 0.0f || v0 > 1.0f) { r.x = (h.x - f.x * v1) / (e.x + g.x * v1); r.y = v1; return r; }
    r.x = u; r.y = v0; return r;
}
// sampling.rs:501-579
inline V3 sample_spherical_rectangle(V3 p_ref, V3 s, V3 ex, V3 ey, V2 u, Float* pdf) {
    const Float exl = length(ex), eyl = length(ey);
    const V3 rx = ex / exl, ry = ey / eyl; V3 rz = cross(rx, ry);                  // Frame::from_xy frame.rs:19-22
    const V3 dv = s - p_ref;
    const V3 d_local = v3(dot(dv, rx), dot(dv, ry), dot(dv, rz));
    Float z0 = d_local.z;
    if (z0 > 0.0f) { rz = -rz; z0 *= -1.0f; }
    const Float x0 = d_local.x, y0 = d_local.y, x1 = x0 + exl, y1 = y0 + eyl;
    const V3 v00 = v3(x0, y0, z0), v01 = v3(x0, y1, z0), v10 = v3(x1, y0, z0), v11 = v3(x1, y1, z0);
    const V3 n0 = normalize(cross(v00, v10)), n1 = normalize(cross(v10, v11)), n2 = normalize(cross(v11, v01)), n3 = normalize(cross(v01, v00));
    const Float g0 = angle_between(-n0, n1), g1 = angle_between(-n1, n2), g2 = angle_between(-n2, n3), g3 = angle_between(-n3, n0);
    const Float solid_angle = g0 + g1 + g2 + g3 - 2.0f * PI_F;
    if (solid_angle <= 0.0f) { *pdf = 0.0f; return s + u.x * ex + u.y * ey; }
    *pdf = fmax_(0.0f, 1.0f / solid_angle);
    if (solid_angle < 1e-3f) return s + u.x * ex + u.y * ey;
    const Float b0 = n0.z, b1 = n2.z;
    const Float au = u.x * (g0 + g1 - 2.0f * PI_F) + (u.x - 1.0f) * (g2 + g3);
    const Float fu = (std::cos(au) * b0 - b1) / std::sin(au);
    Float cu = std::copysign(1.0f / std::sqrt(sqr(fu) + sqr(b0)), fu);
    cu = clampf(cu, -(1.0f - 1.1920929e-07f), 1.0f - 1.1920929e-07f);
    Float xu = -(cu * z0) / safe_sqrt(1.0f - sqr(cu));
    xu = clampf(xu, x0, x1);
    const Float dd = std::sqrt(sqr(xu) + sqr(z0));
    const Float h0 = y0 / std::sqrt(sqr(dd) + sqr(y0)), h1 = y1 / std::sqrt(sqr(dd) + sqr(y1));
    const Float hv = h0 + u.y * (h1 - h0), hvsq = sqr(hv);
    const Float yv = hvsq < 1.0f - 1e-6f ? (hv * dd) / std::sqrt(1.0f - hvsq) : y1;
    return p_ref + (xu * rx + yv * ry + z0 * rz);
}
// sampling.rs:643-787
inline V2 invert_spherical_rectangle_sample(V3 p_ref, V3 s, V3 ex, V3 ey, V3 p_rect) {
    const Float exl = length(ex), eyl = length(ey);
    const V3 rx = ex / exl, ry = ey / eyl; V3 rz = cross(rx, ry);
    const V3 dv = s - p_ref;
    const V3 d_local = v3(dot(dv, rx), dot(dv, ry), dot(dv, rz));
    Float z0 = d_local.z;
    if (z0 > 0.0f) { rz = -rz; z0 *= -1.0f; }
    const Float z0sq = sqr(z0);
    const Float x0 = d_local.x, y0 = d_local.y, x1 = x0 + exl, y1 = y0 + eyl, y0sq = sqr(y0), y1sq = sqr(y1);
    const V3 v00 = v3(x0, y0, z0), v01 = v3(x0, y1, z0), v10 = v3(x1, y0, z0), v11 = v3(x1, y1, z0);
    const V3 n0 = normalize(cross(v00, v10)), n1 = normalize(cross(v10, v11)), n2 = normalize(cross(v11, v01)), n3 = normalize(cross(v01, v00));
    const Float g0 = angle_between(-n0, n1), g1 = angle_between(-n1, n2), g2 = angle_between(-n2, n3), g3 = angle_between(-n3, n0);
    const Float b0 = n0.z, b1 = n2.z, b0sq = sqr(b0);
    const Float solid_angle = g0 + g1 + g2 + g3 - 2.0f * PI_F;
    V2 r;
    if (solid_angle < 1e-3f) { const V3 pq = p_rect - s; r.x = dot(pq, ex) / length_squared(ex); r.y = dot(pq, ey) / length_squared(ey); return r; }
    const V3 pv = p_rect - p_ref;
    const V3 v = v3(dot(pv, rx), dot(pv, ry), dot(pv, rz));
    Float xu = v.x; const Float yv = v.y;
    xu = clampf(xu, x0, x1);
    if (xu == 0.0f) xu = 1e-10f;
    const Float invcusq = 1.0f + z0sq / sqr(xu);
    const Float fusq = invcusq - b0sq;
    const Float fu = std::copysign(std::sqrt(fusq), xu);
    const Float sq = safe_sqrt(difference_of_products(b0, b0, b1, b1) + fusq);
    Float au = std::atan2(-(b1 * fu) - std::copysign(b0 * sq, fu * b0), b0 * b1 - sq * std::fabs(fu));
    if (au > 0.0f) au -= 2.0f * PI_F;
    if (fu == 0.0f) au = PI_F;
    const Float u0 = (au + g2 + g3) / solid_angle;
    const Float ddsq = sqr(xu) + z0sq, dd = std::sqrt(ddsq);
    const Float h0 = y0 / std::sqrt(ddsq + y0sq), h1 = y1 / std::sqrt(ddsq + y1sq);
    const Float yvsq = sqr(yv);
    const Float u1[2] = {(difference_of_products(h0, h0, h0, h1) - std::fabs(h0 - h1) * std::sqrt(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / sqr(h0 - h1),
                         (difference_of_products(h0, h0, h0, h1) + std::fabs(h0 - h1) * std::sqrt(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / sqr(h0 - h1)};
    const Float hv[2] = {lerp(u1[0], h0, h1), lerp(u1[1], h0, h1)};
    const Float hvsq[2] = {sqr(hv[0]), sqr(hv[1])};
    const Float yz[2] = {(hv[0] * dd) / std::sqrt(1.0f - hvsq[0]), (hv[1] * dd) / std::sqrt(1.0f - hvsq[1])};
    r.x = clampf(u0, 0.0f, 1.0f);
    r.y = std::fabs(yz[0] - yv) < std::fabs(yz[1] - yv) ? u1[0] : u1[1];
    return r;
}

// BilinearPatch::is_rectangle bilinear_patch.rs:108-143
inline bool patch_is_rectangle(const V3 q[4]) {
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    auto eq = [](V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
    if (eq(p00, p01) || eq(p01, p11) || eq(p11, p10) || eq(p10, p00)) return false;
    const V3 n = normalize(cross(p10 - p00, p01 - p00));
    if (abs_dot(normalize(p11 - p00), n) > 1e-5f) return false;
    const V3 pc = (p00 + p01 + p10 + p11) * 0.25f;
    const Float d2[4] = {length_squared(p00 - pc), length_squared(p01 - pc), length_squared(p10 - pc), length_squared(p11 - pc)};
    for (int i = 1; i < 4; ++i) if (std::fabs(d2[i] - d2[0]) / d2[0] > 1e-4f) return false;
    return true;
}
inline void patch_shading_flip(const Scene& sc, uint32_t mesh_id, uint32_t patch, V2 uv, V3* n) {      // :570-580, :703-713
    const SgMesh& m = sc.d->meshes[mesh_id];
    if (m.flags & SG_MESH_HAS_N) {
        uint32_t vi[4]; sc.patch_indices(mesh_id, patch, vi);
        const V3 ns = lerp3(uv.x, lerp3(uv.y, sc.normal(m, vi[0]), sc.normal(m, vi[2])), lerp3(uv.y, sc.normal(m, vi[1]), sc.normal(m, vi[3])));
        *n = face_forward(*n, ns);
    } else if (((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0)) *n = -*n;
}
// BilinearPatch::sample bilinear_patch.rs:521-600
inline bool patch_sample_area(const Scene& sc, uint32_t mesh_id, uint32_t patch, V2 u, ShapeSample* ss) {
    V3 q[4]; sc.patch_points(mesh_id, patch, q);
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    V2 uv = u; Float pdf = 1.0f;
    if (!patch_is_rectangle(q)) {
        const Float w[4] = {length(cross(p10 - p00, p01 - p00)), length(cross(p10 - p00, p11 - p10)), length(cross(p01 - p00, p11 - p01)), length(cross(p11 - p10, p11 - p01))};
        uv = sample_bilinear(u, w); pdf = bilinear_pdf(uv, w);
    }
    const V3 pu0 = lerp3(uv.x, p00, p10), pu1 = lerp3(uv.y, p10, p11);                // sic (:544-545)
    const V3 p = lerp3(uv.x, pu0, pu1);
    const V3 dpdu = pu1 - pu0;
    const V3 dpdv = lerp3(uv.x, p01, p11) - lerp3(uv.x, p00, p10);
    if (length_squared(dpdu) == 0.0f || length_squared(dpdv) == 0.0f) return false;
    V3 n = normalize(cross(dpdu, dpdv));
    patch_shading_flip(sc, mesh_id, patch, uv, &n);
    const V3 p_abs_sum = vabs(p00) + vabs(p01) + vabs(p10) + vabs(p11);
    ss->pi = p3fi_from_value_and_error(p, gamma_n(6) * p_abs_sum); ss->n = n; ss->pdf = pdf / length(cross(dpdu, dpdv));
    return true;
}
// BilinearPatch::pdf bilinear_patch.rs:602-636 (`st` = Interaction::uv of the intersection)
inline Float patch_pdf_area(const Scene& sc, uint32_t mesh_id, uint32_t patch, V2 st) {
    const SgMesh& m = sc.d->meshes[mesh_id];
    V3 q[4]; sc.patch_points(mesh_id, patch, q);
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    V2 uv = st;
    if (m.flags & SG_MESH_HAS_UV) {
        uint32_t vi[4]; sc.patch_indices(mesh_id, patch, vi);
        const V2 verts[4] = {sc.uv(m, vi[0]), sc.uv(m, vi[1]), sc.uv(m, vi[2]), sc.uv(m, vi[3])};
        uv = invert_bilinear(st, verts);
    }
    Float pdf = 1.0f;
    if (!patch_is_rectangle(q)) {
        const Float w[4] = {length(cross(p10 - p00, p01 - p00)), length(cross(p10 - p00, p11 - p10)), length(cross(p01 - p00, p11 - p01)), length(cross(p11 - p10, p11 - p01))};
        pdf = bilinear_pdf(uv, w);
    }
    const V3 pu0 = lerp3(uv.y, p00, p10), pu1 = lerp3(uv.y, p10, p11);                // sic (:628-629)
    const V3 dpdu = pu1 - pu0;
    const V3 dpdv = lerp3(uv.x, p01, p11) - lerp3(uv.x, p00, p10);
    return pdf / length(cross(dpdu, dpdv));
}
// BilinearPatch::sample_with_context bilinear_patch.rs:638-737
inline bool patch_sample_with_context(const Scene& sc, uint32_t mesh_id, uint32_t patch, const LightSampleContext& ctx, V2 u, ShapeSample* ss) {
    V3 q[4]; sc.patch_points(mesh_id, patch, q);
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    const V3 cp = ctx.p();
    const V3 v00 = normalize(p00 - cp), v10 = normalize(p10 - cp), v01 = normalize(p01 - cp), v11 = normalize(p11 - cp);
    if (!patch_is_rectangle(q) || spherical_quad_area(v00, v10, v11, v01) <= 1e-4f) {
        if (!patch_sample_area(sc, mesh_id, patch, u, ss)) return false;              // `.unwrap()` panics in the reference on a degenerate sample
        V3 wi = p3fi_mid(ss->pi) - cp;
        if (length_squared(wi) == 0.0f) return false;
        wi = normalize(wi);
        ss->pdf /= abs_dot(ss->n, -wi) / length_squared(cp - p3fi_mid(ss->pi));
        if (std::isinf(ss->pdf)) return false;
        return true;
    }
    Float pdf = 1.0f;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        const Float w[4] = {fmax_(0.01f, dot(v00, ctx.ns)), fmax_(0.01f, dot(v10, ctx.ns)), fmax_(0.01f, dot(v01, ctx.ns)), fmax_(0.01f, dot(v11, ctx.ns))};
        u = sample_bilinear(u, w);
        pdf = bilinear_pdf(u, w);
    }
    const V3 eu = p10 - p00, ev = p01 - p00;
    Float quad_pdf = 0.0f;
    const V3 p = sample_spherical_rectangle(cp, p00, eu, ev, u, &quad_pdf);
    pdf *= quad_pdf;
    V2 uv; uv.x = dot(p - p00, eu) / distance_squared(p10, p00); uv.y = dot(p - p00, ev) / distance_squared(p01, p00);
    V3 n = normalize(cross(eu, ev));
    patch_shading_flip(sc, mesh_id, patch, uv, &n);
    ss->pi = p3fi_exact(p); ss->n = n; ss->pdf = pdf;
    return true;
}
// BilinearPatch::pdf_with_context bilinear_patch.rs:739-783
inline Float patch_pdf_with_context(const Scene& sc, uint32_t mesh_id, uint32_t patch, const LightSampleContext& ctx, V3 wi) {
    V3 q[4]; sc.patch_points(mesh_id, patch, q);
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    const V3 cp = ctx.p();
    const V3 ro = offset_ray_origin(ctx.pi, ctx.n, wi);                              // ShapeSampleContext::spawn_ray shape.rs:276-283
    Float bu, bv, bt;
    if (!intersect_blp(ro, wi, F_INF, p00, p10, p01, p11, &bu, &bv, &bt)) return 0.0f;
    const SurfaceInteraction isect = patch_interaction(sc, mesh_id, patch, bu, bv, -wi);
    const V3 v00 = normalize(p00 - cp), v10 = normalize(p10 - cp), v01 = normalize(p01 - cp), v11 = normalize(p11 - cp);
    if (!patch_is_rectangle(q) || spherical_quad_area(v00, v10, v11, v01) <= 1e-4f) {
        const Float pdf = patch_pdf_area(sc, mesh_id, patch, isect.uv) * distance_squared(cp, isect.p()) / abs_dot(isect.n, -wi);
        return std::isinf(pdf) ? 0.0f : pdf;
    }
    const Float pdf = 1.0f / spherical_quad_area(v00, v10, v11, v01);
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        const Float w[4] = {fmax_(0.01f, dot(v00, ctx.ns)), fmax_(0.01f, dot(v10, ctx.ns)), fmax_(0.01f, dot(v01, ctx.ns)), fmax_(0.01f, dot(v11, ctx.ns))};
        const V2 u = invert_spherical_rectangle_sample(cp, p00, p10 - p00, p01 - p00, isect.p());
        return bilinear_pdf(u, w) * pdf;
    }
    return pdf;
}

}  // namespace orc
