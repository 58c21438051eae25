// BilinearPatch as an emitter on the device: is_rectangle (bilinear_patch.rs:108-143), sample (:521-600), pdf (:602-636),
// sample_with_context (:638-737), pdf_with_context (:739-783); sample_spherical_rectangle / invert_spherical_rectangle_sample
// (sampling.rs:501-579, 643-787), invert_bilinear / spherical_quad_area (vecmath/mod.rs:70-140).  As written in the reference,
// including Sample's `pu0 = lerp(uv[0], p00, p10)`, `pu1 = lerp(uv[1], p10, p11)` (pbrt interpolates both along v), which puts
// area-sampled points off a non-rectangular patch.  Out of line: only scenes with emissive patches call these.
// Included after sg_sphere_surface.cuh (make_surface_patch).
#pragma once

namespace sg {

SGD float spherical_quad_area(float3 a, float3 b, float3 c, float3 d) {
    float3 axb = cross3(a, b), bxc = cross3(b, c), cxd = cross3(c, d), dxa = cross3(d, a);
    if (len2(axb) == 0.0f || len2(bxc) == 0.0f || len2(cxd) == 0.0f || len2(dxa) == 0.0f) return 0.0f;
    axb = normalize3(axb); bxc = normalize3(bxc); cxd = normalize3(cxd); dxa = normalize3(dxa);
    const float alpha = angle_between3(dxa, -axb), beta = angle_between3(axb, -bxc), gamma = angle_between3(bxc, -cxd), delta = angle_between3(cxd, -dxa);
    return fabsf(alpha + beta + gamma + delta - 2.0f * kPi);
}
SGD float cross2d(float2 a, float2 b) { return dop(a.x, b.y, a.y, b.x); }
SGD float2 invert_bilinear(float2 p, const float2 vert[4]) {
    const float2 a = vert[0], b = vert[1], c = vert[3], d = vert[2];
    const float2 e = make_float2(b.x - a.x, b.y - a.y), f = make_float2(d.x - a.x, d.y - a.y),
                 g = make_float2((a.x - b.x) + (c.x - d.x), (a.y - b.y) + (c.y - d.y)), h = make_float2(p.x - a.x, p.y - a.y);
    const float k2 = cross2d(g, f), k1 = cross2d(e, f) + cross2d(h, g), k0 = cross2d(h, e);
    if (fabsf(k2) < 0.001f) {
        if (fabsf(e.x * k1 - g.x * k0) < 1e-5f) return make_float2((h.y * k1 + f.y * k0) / (e.y * k1 - g.y * k0), -k0 / k1);
        return make_float2((h.x * k1 + f.x * k0) / (e.x * k1 - g.x * k0), -k0 / k1);
    }
    float v0, v1;
    if (!quadratic(k2, k1, k0, v0, v1)) return make_float2(0.0f, 0.0f);
    const float u = (h.x - f.x * v0) / (e.x + g.x * v0);
    if (u < 0.0f || u > 1.0f || v0 < 0.0f || v0 > 1.0f) return make_float2((h.x - f.x * v1) / (e.x + g.x * v1), v1);
    return make_float2(u, v0);
}
struct SphRect { float3 rx, ry, rz; float x0, y0, x1, y1, z0, g0, g1, g2, g3, b0, b1, solid_angle; };
SGD SphRect sph_rect_init(float3 p_ref, float3 s, float3 ex, float3 ey) {          // the common prologue of sampling.rs:501-540 and :643-700
    SphRect R;
    const float exl = len3(ex), eyl = len3(ey);
    R.rx = ex / exl; R.ry = ey / eyl; R.rz = cross3(R.rx, R.ry);
    const float3 dv = s - p_ref;
    const float3 d_local = f3(dot3(dv, R.rx), dot3(dv, R.ry), dot3(dv, R.rz));
    R.z0 = d_local.z;
    if (R.z0 > 0.0f) { R.rz = -R.rz; R.z0 *= -1.0f; }
    R.x0 = d_local.x; R.y0 = d_local.y; R.x1 = R.x0 + exl; R.y1 = R.y0 + eyl;
    const float3 v00 = f3(R.x0, R.y0, R.z0), v01 = f3(R.x0, R.y1, R.z0), v10 = f3(R.x1, R.y0, R.z0), v11 = f3(R.x1, R.y1, R.z0);
    const float3 n0 = normalize3(cross3(v00, v10)), n1 = normalize3(cross3(v10, v11)), n2 = normalize3(cross3(v11, v01)), n3 = normalize3(cross3(v01, v00));
    R.g0 = angle_between3(-n0, n1); R.g1 = angle_between3(-n1, n2); R.g2 = angle_between3(-n2, n3); R.g3 = angle_between3(-n3, n0);
    R.b0 = n0.z; R.b1 = n2.z;
    R.solid_angle = R.g0 + R.g1 + R.g2 + R.g3 - 2.0f * kPi;
    return R;
}
SGD float3 sample_spherical_rectangle(float3 p_ref, float3 s, float3 ex, float3 ey, float2 u, float& pdf) {
    const SphRect R = sph_rect_init(p_ref, s, ex, ey);
    if (R.solid_angle <= 0.0f) { pdf = 0.0f; return s + u.x * ex + u.y * ey; }
    pdf = fmaxf(0.0f, 1.0f / R.solid_angle);
    if (R.solid_angle < 1e-3f) return s + u.x * ex + u.y * ey;
    const float au = u.x * (R.g0 + R.g1 - 2.0f * kPi) + (u.x - 1.0f) * (R.g2 + R.g3);
    const float fu = (cosf(au) * R.b0 - R.b1) / sinf(au);
    float cu = copysignf(1.0f / sqrtf(sqr(fu) + sqr(R.b0)), fu);
    cu = clampf(cu, -(1.0f - 1.1920929e-07f), 1.0f - 1.1920929e-07f);
    float xu = -(cu * R.z0) / safe_sqrt(1.0f - sqr(cu));
    xu = clampf(xu, R.x0, R.x1);
    const float dd = sqrtf(sqr(xu) + sqr(R.z0));
    const float h0 = R.y0 / sqrtf(sqr(dd) + sqr(R.y0)), h1 = R.y1 / sqrtf(sqr(dd) + sqr(R.y1));
    const float hv = h0 + u.y * (h1 - h0), hvsq = sqr(hv);
    const float yv = hvsq < 1.0f - 1e-6f ? (hv * dd) / sqrtf(1.0f - hvsq) : R.y1;
    return p_ref + (xu * R.rx + yv * R.ry + R.z0 * R.rz);
}
SGD float2 invert_spherical_rectangle_sample(float3 p_ref, float3 s, float3 ex, float3 ey, float3 p_rect) {
    const SphRect R = sph_rect_init(p_ref, s, ex, ey);
    const float z0sq = sqr(R.z0), y0sq = sqr(R.y0), y1sq = sqr(R.y1), b0sq = sqr(R.b0);
    if (R.solid_angle < 1e-3f) { const float3 pq = p_rect - s; return make_float2(dot3(pq, ex) / len2(ex), dot3(pq, ey) / len2(ey)); }
    const float3 pv = p_rect - p_ref;
    float xu = dot3(pv, R.rx); const float yv = dot3(pv, R.ry);
    xu = clampf(xu, R.x0, R.x1);
    if (xu == 0.0f) xu = 1e-10f;
    const float invcusq = 1.0f + z0sq / sqr(xu);
    const float fusq = invcusq - b0sq;
    const float fu = copysignf(sqrtf(fusq), xu);
    const float sq = safe_sqrt(dop(R.b0, R.b0, R.b1, R.b1) + fusq);
    float au = atan2f(-(R.b1 * fu) - copysignf(R.b0 * sq, fu * R.b0), R.b0 * R.b1 - sq * fabsf(fu));
    if (au > 0.0f) au -= 2.0f * kPi;
    if (fu == 0.0f) au = kPi;
    const float u0 = (au + R.g2 + R.g3) / R.solid_angle;
    const float ddsq = sqr(xu) + z0sq, dd = sqrtf(ddsq);
    const float h0 = R.y0 / sqrtf(ddsq + y0sq), h1 = R.y1 / sqrtf(ddsq + y1sq);
    const float yvsq = sqr(yv);
    const float u1a = (dop(h0, h0, h0, h1) - fabsf(h0 - h1) * sqrtf(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / sqr(h0 - h1);
    const float u1b = (dop(h0, h0, h0, h1) + fabsf(h0 - h1) * sqrtf(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / sqr(h0 - h1);
    const float hva = lerpf(u1a, h0, h1), hvb = lerpf(u1b, h0, h1);
    const float yza = (hva * dd) / sqrtf(1.0f - sqr(hva)), yzb = (hvb * dd) / sqrtf(1.0f - sqr(hvb));
    return make_float2(clampf(u0, 0.0f, 1.0f), fabsf(yza - yv) < fabsf(yzb - yv) ? u1a : u1b);
}

struct PatchGeo { float3 p00, p10, p01, p11; uint32_t flags; SgMesh m; const uint32_t* ix; };
SGD PatchGeo patch_geo(const DScene& sc, uint32_t rec) {
    const float4 a0 = __ldg(sc.patch_verts + 4 * (size_t)rec), a1 = __ldg(sc.patch_verts + 4 * (size_t)rec + 1),
                 a2 = __ldg(sc.patch_verts + 4 * (size_t)rec + 2), a3 = __ldg(sc.patch_verts + 4 * (size_t)rec + 3);
    PatchGeo g;
    g.p00 = f3(a0.x, a0.y, a0.z); g.p10 = f3(a1.x, a1.y, a1.z); g.p01 = f3(a2.x, a2.y, a2.z); g.p11 = f3(a3.x, a3.y, a3.z);
    g.flags = __float_as_uint(a0.w); g.m = sc.meshes[__float_as_uint(a1.w)];
    g.ix = sc.indices + g.m.first_index + 4 * (size_t)__float_as_uint(a2.w);
    return g;
}
SGD bool patch_is_rectangle(const PatchGeo& g) {
    const float3 p00 = g.p00, p10 = g.p10, p01 = g.p01, p11 = g.p11;
    auto eq = [](float3 a, float3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
    if (eq(p00, p01) || eq(p01, p11) || eq(p11, p10) || eq(p10, p00)) return false;
    const float3 n = normalize3(cross3(p10 - p00, p01 - p00));
    if (absdot3(normalize3(p11 - p00), n) > 1e-5f) return false;
    const float3 pc = (p00 + p01 + p10 + p11) * 0.25f;
    const float d0 = len2(p00 - pc), d1 = len2(p01 - pc), d2 = len2(p10 - pc), d3 = len2(p11 - pc);
    if (fabsf(d1 - d0) / d0 > 1e-4f || fabsf(d2 - d0) / d0 > 1e-4f || fabsf(d3 - d0) / d0 > 1e-4f) return false;
    return true;
}
SGD float3 patch_shading_flip(const DScene& sc, const PatchGeo& g, float2 uv, float3 n) {          // :570-580, :703-713
    if (g.flags & SG_MESH_HAS_N) {
        const size_t fv = g.m.first_vertex;
        const float3 n00 = ldv3(sc.n, fv + __ldg(g.ix)), n10 = ldv3(sc.n, fv + __ldg(g.ix + 1)), n01 = ldv3(sc.n, fv + __ldg(g.ix + 2)), n11 = ldv3(sc.n, fv + __ldg(g.ix + 3));
        return faceforward3(n, lerp3(uv.x, lerp3(uv.y, n00, n01), lerp3(uv.y, n10, n11)));
    }
    if (((g.flags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((g.flags & SG_MESH_SWAPS_HANDEDNESS) != 0)) return -n;
    return n;
}
SGD void patch_corner_weights(const PatchGeo& g, float w[4]) {
    w[0] = len3(cross3(g.p10 - g.p00, g.p01 - g.p00)); w[1] = len3(cross3(g.p10 - g.p00, g.p11 - g.p10));
    w[2] = len3(cross3(g.p01 - g.p00, g.p11 - g.p01)); w[3] = len3(cross3(g.p11 - g.p10, g.p11 - g.p01));
}
// BilinearPatch::sample bilinear_patch.rs:521-600
SGD bool patch_sample_area(const DScene& sc, const PatchGeo& g, bool rect, float2 u, P3fi& out_pi, float3& out_n, float& out_pdf) {
    float2 uv = u; float pdf = 1.0f;
    if (!rect) { float w[4]; patch_corner_weights(g, w); uv = sample_bilinear(u, w); pdf = bilinear_pdf(uv, w); }
    const float3 pu0 = lerp3(uv.x, g.p00, g.p10), pu1 = lerp3(uv.y, g.p10, g.p11);                 // sic (:544-545)
    const float3 p = lerp3(uv.x, pu0, pu1);
    const float3 dpdu = pu1 - pu0;
    const float3 dpdv = lerp3(uv.x, g.p01, g.p11) - lerp3(uv.x, g.p00, g.p10);
    if (len2(dpdu) == 0.0f || len2(dpdv) == 0.0f) return false;
    const float3 n = patch_shading_flip(sc, g, uv, normalize3(cross3(dpdu, dpdv)));
    out_pi = p3fi_make(p, gamma_n(6) * (abs3(g.p00) + abs3(g.p01) + abs3(g.p10) + abs3(g.p11)));
    out_n = n; out_pdf = pdf / len3(cross3(dpdu, dpdv));
    return true;
}
// BilinearPatch::pdf bilinear_patch.rs:602-636
SGD float patch_pdf_area(const DScene& sc, const PatchGeo& g, bool rect, float2 st) {
    float2 uv = st;
    if (g.flags & SG_MESH_HAS_UV) {
        const size_t fv = g.m.first_vertex;
        float2 verts[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { const float* q = sc.uv + 2 * (fv + __ldg(g.ix + k)); verts[k] = make_float2(__ldg(q), __ldg(q + 1)); }
        uv = invert_bilinear(st, verts);
    }
    float pdf = 1.0f;
    if (!rect) { float w[4]; patch_corner_weights(g, w); pdf = bilinear_pdf(uv, w); }
    const float3 pu0 = lerp3(uv.y, g.p00, g.p10), pu1 = lerp3(uv.y, g.p10, g.p11);                 // sic (:628-629)
    const float3 dpdu = pu1 - pu0;
    const float3 dpdv = lerp3(uv.x, g.p01, g.p11) - lerp3(uv.x, g.p00, g.p10);
    return pdf / len3(cross3(dpdu, dpdv));
}
// BilinearPatch::sample_with_context bilinear_patch.rs:638-737
static __device__ __noinline__ bool patch_sample_with_context(const DScene& sc, uint32_t rec, const LightCtx& ctx, float2 u, P3fi& out_pi, float3& out_n, float& out_pdf) {
    const PatchGeo g = patch_geo(sc, rec);
    const float3 cp = p3fi_mid(ctx.pi);
    const float3 v00 = normalize3(g.p00 - cp), v10 = normalize3(g.p10 - cp), v01 = normalize3(g.p01 - cp), v11 = normalize3(g.p11 - cp);
    const bool rect = patch_is_rectangle(g);
    if (!rect || spherical_quad_area(v00, v10, v11, v01) <= 1e-4f) {
        if (!patch_sample_area(sc, g, rect, u, out_pi, out_n, out_pdf)) return false;
        const float3 sp = p3fi_mid(out_pi);
        float3 wi = sp - cp;
        if (len2(wi) == 0.0f) return false;
        wi = normalize3(wi);
        out_pdf /= absdot3(out_n, -wi) / len2(cp - sp);
        if (isinf(out_pdf)) return false;
        return true;
    }
    float pdf = 1.0f;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        const float w[4] = {fmaxf(0.01f, dot3(v00, ctx.ns)), fmaxf(0.01f, dot3(v10, ctx.ns)), fmaxf(0.01f, dot3(v01, ctx.ns)), fmaxf(0.01f, dot3(v11, ctx.ns))};
        u = sample_bilinear(u, w);
        pdf = bilinear_pdf(u, w);
    }
    const float3 eu = g.p10 - g.p00, ev = g.p01 - g.p00;
    float quad_pdf = 0.0f;
    const float3 p = sample_spherical_rectangle(cp, g.p00, eu, ev, u, quad_pdf);
    pdf *= quad_pdf;
    const float2 uv = make_float2(dot3(p - g.p00, eu) / dist2(g.p10, g.p00), dot3(p - g.p00, ev) / dist2(g.p01, g.p00));
    out_n = patch_shading_flip(sc, g, uv, normalize3(cross3(eu, ev)));
    out_pi = p3fi_exact(p); out_pdf = pdf;
    return true;
}
// BilinearPatch::pdf_with_context bilinear_patch.rs:739-783
static __device__ __noinline__ float patch_pdf_with_context(const DScene& sc, uint32_t rec, const LightCtx& ctx, float3 wi) {
    const PatchGeo g = patch_geo(sc, rec);
    const float3 cp = p3fi_mid(ctx.pi);
    const float3 ro = offset_ray_origin(ctx.pi, ctx.n, wi);
    float bu, bv, bt;
    if (!intersect_blp(ro, wi, INFINITY, g.p00, g.p10, g.p01, g.p11, bu, bv, bt)) return 0.0f;
    SurfTex sx;
    const Surf isect = make_surface_patch<true>(sc, rec, bu, bv, &sx);
    const float3 ip = p3fi_mid(isect.pi);
    const float3 v00 = normalize3(g.p00 - cp), v10 = normalize3(g.p10 - cp), v01 = normalize3(g.p01 - cp), v11 = normalize3(g.p11 - cp);
    const bool rect = patch_is_rectangle(g);
    const float sqa = spherical_quad_area(v00, v10, v11, v01);
    if (!rect || sqa <= 1e-4f) {
        const float pdf = patch_pdf_area(sc, g, rect, sx.uv) * dist2(cp, ip) / absdot3(isect.n, -wi);
        return isinf(pdf) ? 0.0f : pdf;
    }
    const float pdf = 1.0f / sqa;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        const float w[4] = {fmaxf(0.01f, dot3(v00, ctx.ns)), fmaxf(0.01f, dot3(v10, ctx.ns)), fmaxf(0.01f, dot3(v01, ctx.ns)), fmaxf(0.01f, dot3(v11, ctx.ns))};
        const float2 u = invert_spherical_rectangle_sample(cp, g.p00, g.p10 - g.p00, g.p01 - g.p00, ip);
        return bilinear_pdf(u, w) * pdf;
    }
    return pdf;
}

// ---- the non-triangle branches of Light::sample_li / pdf_li, one out-of-line entry each (see sg_shading.cuh) ----
static __device__ __noinline__ bool light_sample_li_other(const DScene& sc, uint32_t light_id, const LightCtx& ctx, float2 u, const Wavelengths& lam,
                                                   bool allow_incomplete, LightSample& ls) {
    const SgLight lt = sc.lights[light_id];
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE || lt.kind == SG_LIGHT_DIFFUSE_AREA_PATCH) {            // light.rs:632-661
        P3fi pi; float3 n; float pdf;
        if (lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE) { if (!sphere_sample_with_context(sc.spheres[lt.tri], ctx, u, pi, n, pdf)) return false; }
        else if (!patch_sample_with_context(sc, __float_as_uint(__ldg(sc.light_verts + 3 * (size_t)light_id).w), ctx, u, pi, n, pdf)) return false;
        const float3 sp = p3fi_mid(pi), cp = p3fi_mid(ctx.pi);
        if (pdf == 0.0f || len2(sp - cp) == 0.0f) return false;
        const float3 wi = normalize3(sp - cp);
        const Spec le = light_l(sc, lt, n, -wi, lam);
        if (spec_zero(le)) return false;
        ls.l = le; ls.wi = wi; ls.pdf = pdf; ls.p_light = pi; ls.n_light = n;
        return true;
    }
    if (lt.kind == SG_LIGHT_POINT) {                                                                    // light.rs:461-484
        const float3 p = f3(lt.pos[0], lt.pos[1], lt.pos[2]), cp = p3fi_mid(ctx.pi);
        ls.wi = normalize3(p - cp);
        ls.l = lt.scale * spectrum_sample(sc, lt.spectrum, lam) / dist2(p, cp);
        ls.pdf = 1.0f; ls.p_light = p3fi_exact(p); ls.n_light = f3(0.0f, 0.0f, 0.0f);
        return true;
    }
    return infinite_sample_li(sc, lt, ctx, u, lam, allow_incomplete, ls);                                // light.rs:740-766, :847-880
}
static __device__ __noinline__ float light_pdf_li_other(const DScene& sc, uint32_t light_id, uint32_t hit_mesh_word, const LightCtx& ctx, float3 wi) {
    const SgLight lt = sc.lights[light_id];
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE) return sphere_pdf_with_context(sc, sc.spheres[lt.tri], ctx, wi);
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA_PATCH) return patch_pdf_with_context(sc, hit_mesh_word & ~kPatchBit, ctx, wi);   // TriGeo::mesh of a patch hit = record | kPatchBit
    return 0.0f;                                                             // light.rs:486-494 (infinite lights are handled by k_shade_miss)
}

}  // namespace sg
