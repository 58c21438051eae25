// Image textures, MIP filtering, screen-space differentials and bump mapping on the device.
//
// Reference: SpectrumImageTexture / FloatImageTexture::evaluate (texture.rs:393-404,777-808), UVMapping::map
// (:918-936), MIPMap::filter / ewa (mipmap.rs:121-293), Image::get_channel_wrapped / bilerp_channel_wrapped /
// remap_pixel_coords (image.rs:134-177,452-475,619-646), RgbAlbedoSpectrum / RgbUnboundedSpectrum
// (spectrum.rs:498-588), RgbSigmoidPolynomial::get (color.rs:352-383), compute_differentials
// (interaction.rs:280-366), Camera::approximate_dp_dxy (camera.rs:308-354), Transform::rotate_from_to
// (transform.rs:227-253), bump_map (material.rs:1477-1509), spawn_ray_with_differentials (interaction.rs:434-502).
//
// Layout: every MIP level is a dense row-major array of f32 texels with interleaved channels in one pool
// (`texels`); a bilinear tap is four scalar __ldg's per channel (two 8-byte-adjacent pairs), served from L1/L2.
// The third-party pieces (rgb2spec fetch, fast_polynomial's Estrin `poly`) follow the published algorithms.
#pragma once
#include "sg_shading.cuh"

namespace sg {

struct AuxRays { bool has; float3 rxo, rxd, ryo, ryd; };

// auxiliary rays of PerspectiveCamera::generate_ray_differential (camera.rs:1036-1079) in render space
// (Transform::apply_ray(RayDifferential) transform.rs:534-556: plain point / vector transforms)
SGD void camera_aux(const DScene& sc, float3 p_camera, float2 p_lens, float3 o_cam, AuxRays* aux) {
    const SgCamera& cam = sc.camera;
    const float3 dxc = f3(cam.dx_camera[0], cam.dx_camera[1], cam.dx_camera[2]), dyc = f3(cam.dy_camera[0], cam.dy_camera[1], cam.dy_camera[2]);
    float3 rxo, rxd, ryo, ryd;
    if (cam.kind == SG_CAMERA_ORTHOGRAPHIC) {                          // camera.rs:775-778: shifted origins, same direction, camera space
        aux->has = true;
        aux->rxo = o_cam + dxc; aux->rxd = f3(0.0f, 0.0f, 1.0f); aux->ryo = o_cam + dyc; aux->ryd = f3(0.0f, 0.0f, 1.0f);
        return;
    }
    if (cam.lens_radius > 0.0f) {
        float2 pl = sample_disk_concentric(p_lens);
        pl.x = cam.lens_radius * pl.x; pl.y = cam.lens_radius * pl.y;
        const float3 dx = normalize3(p_camera + dxc);
        float ft = cam.focal_distance / dx.z;
        float3 pf = f3(0.0f, 0.0f, 0.0f) + ft * dx;
        rxo = f3(pl.x, pl.y, 0.0f); rxd = normalize3(pf - rxo);
        const float3 dy = normalize3(p_camera + dyc);
        ft = cam.focal_distance / dy.z;
        pf = f3(0.0f, 0.0f, 0.0f) + ft * dy;
        ryo = f3(pl.x, pl.y, 0.0f); ryd = normalize3(pf - ryo);
    } else {
        rxo = o_cam; ryo = o_cam;
        rxd = normalize3(p_camera + dxc); ryd = normalize3(p_camera + dyc);
    }
    aux->has = true;
    aux->rxo = xform_point(cam.render_from_camera, rxo); aux->rxd = xform_vector(cam.render_from_camera, rxd);
    aux->ryo = xform_point(cam.render_from_camera, ryo); aux->ryd = xform_vector(cam.render_from_camera, ryd);
}

SGD int modulo_i(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }            // math.rs:439-451
SGD uint32_t f2u_sat(float f) { return __float2uint_rz(f); }                                       // Rust `as usize`: saturating, NaN -> 0

struct TexView {
    const DScene& sc; const SgTexture& t;
    SGD SgImageLevel level(int l) const { return sc.image_levels[t.first_level + l]; }
    // Image::get_channel_wrapped image.rs:452-475 + remap_pixel_coords :134-177.  The wrap is evaluated ONCE per texel for
    // all its channels (the reference redoes it per channel with the same integers): returns false for a black border texel.
    SGD bool remap(const SgImageLevel& L, int& x, int& y) const {
        if (x < 0 || x >= L.res[0]) {
            if (t.wrap == SG_WRAP_BLACK) return false;
            x = t.wrap == SG_WRAP_CLAMP ? min(max(x, 0), L.res[0] - 1) : modulo_i(x, L.res[0]);
        }
        if (y < 0 || y >= L.res[1]) {
            if (t.wrap == SG_WRAP_BLACK) return false;
            y = t.wrap == SG_WRAP_CLAMP ? min(max(y, 0), L.res[1] - 1) : modulo_i(y, L.res[1]);
        }
        return true;
    }
    // texel value: RGB (texel_rgb mipmap.rs:203-219) or Float replicated (texel_float :221-225)
    template <bool RGB> SGD float3 texel(const SgImageLevel& L, int x, int y) const {
        if (!remap(L, x, y)) return f3(0.0f, 0.0f, 0.0f);
        const float* q = sc.texels + (size_t)L.offset + ((size_t)y * L.res[0] + x) * t.n_channels;
        if (RGB && t.n_channels == 3) return f3(__ldg(q), __ldg(q + 1), __ldg(q + 2));
        const float v = __ldg(q); return f3(v, v, v);
    }
    // Image::bilerp_channel_wrapped image.rs:619-646, all channels of the four taps at once (per-channel arithmetic unchanged)
    template <bool RGB> SGD float3 bilerp(const SgImageLevel& L, float2 st) const {
        const float x = st.x * (float)L.res[0] - 0.5f, y = st.y * (float)L.res[1] - 0.5f;
        const int xi = f2i_sat(floorf(x)), yi = f2i_sat(floorf(y));
        const float dx = x - (float)xi, dy = y - (float)yi;
        const float3 v0 = texel<RGB>(L, xi, yi), v1 = texel<RGB>(L, xi + 1, yi), v2 = texel<RGB>(L, xi, yi + 1), v3 = texel<RGB>(L, xi + 1, yi + 1);
        const float w0 = (1.0f - dx) * (1.0f - dy), w1 = dx * (1.0f - dy), w2 = (1.0f - dx) * dy, w3 = dx * dy;
        return f3(w0 * v0.x + w1 * v1.x + w2 * v2.x + w3 * v3.x, w0 * v0.y + w1 * v1.y + w2 * v2.y + w3 * v3.y,
                  w0 * v0.z + w1 * v1.z + w2 * v2.z + w3 * v3.z);
    }
};

SGD float3 tex_lerp(float t, float3 a, float3 b) { return a * (1.0f - t) + b * t; }

template <bool RGB> SGD float3 tex_texel(const TexView& tv, int l, int x, int y) { return tv.texel<RGB>(tv.level(l), x, y); }
// one out-of-line copy per texel type: MIPMap::filter reaches it from four places and the shade kernels are I-cache bound
template <bool RGB> __device__ __noinline__ float3 tex_bilerp(const TexView& tv, int l, float2 st) { return tv.bilerp<RGB>(tv.level(l), st); }   // mipmap.rs:298-331
// TexelType::ewa mipmap.rs:233-293
template <bool RGB> __device__ __noinline__ float3 tex_ewa(const TexView& tv, int l, float2 st, float2 d0, float2 d1) {
    if (l >= tv.t.n_levels) return tex_texel<RGB>(tv, tv.t.n_levels - 1, 0, 0);
    const SgImageLevel L = tv.level(l);
    st.x = st.x * (float)L.res[0] - 0.5f; st.y = st.y * (float)L.res[1] - 0.5f;
    d0.x *= (float)L.res[0]; d0.y *= (float)L.res[1]; d1.x *= (float)L.res[0]; d1.y *= (float)L.res[1];
    float a = sqr(d0.y) + sqr(d1.y) + 1.0f;
    float b = -2.0f * (d0.x * d0.y + d1.x * d1.y);
    float c = sqr(d0.x) + sqr(d1.x) + 1.0f;
    const float inv_f = 1.0f / (a * c - sqr(b) * 0.25f);
    a *= inv_f; b *= inv_f; c *= inv_f;
    const float det = -sqr(b) + 4.0f * a * c;
    const float inv_det = 1.0f / det;
    const float u_sqrt = safe_sqrt(det * c), v_sqrt = safe_sqrt(a * det);
    const int s0 = f2i_sat(ceilf(st.x - 2.0f * inv_det * u_sqrt)), s1 = f2i_sat(floorf(st.x + 2.0f * inv_det * u_sqrt));
    const int t0 = f2i_sat(ceilf(st.y - 2.0f * inv_det * v_sqrt)), t1 = f2i_sat(floorf(st.y + 2.0f * inv_det * v_sqrt));
    float3 sum = f3(0.0f, 0.0f, 0.0f); float sum_wts = 0.0f;
    for (int it = t0; it <= t1; ++it) {
        const float tt = (float)it - st.y;
        for (int is = s0; is <= s1; ++is) {
            const float ss = (float)is - st.x;
            const float r2 = a * sqr(ss) + b * ss * tt + c * sqr(tt);
            if (r2 < 1.0f) {
                const uint32_t index = min(f2u_sat(r2 * 128.0f), 127u);
                const float w = __ldg(tv.sc.mip_lut + index);
                const float3 tx = tv.texel<RGB>(L, is, it);
                sum = sum + tx * w;
                sum_wts += w;
            }
        }
    }
    return sum / sum_wts;
}
// MIPMap::filter mipmap.rs:121-201
template <bool RGB> SGD float3 tex_filter(const TexView& tv, float2 st, float2 dst0, float2 dst1) {
    const int n_levels = tv.t.n_levels;
    if (tv.t.filter == SG_FILTER_EWA) {
        if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { const float2 tmp = dst0; dst0 = dst1; dst1 = tmp; }
        const float longer = sqrtf(dst0.x * dst0.x + dst0.y * dst0.y);
        float shorter = sqrtf(dst1.x * dst1.x + dst1.y * dst1.y);
        if (shorter * tv.t.max_anisotropy < longer && shorter > 0.0f) {
            const float scale = longer / (shorter * tv.t.max_anisotropy);
            dst1.x *= scale; dst1.y *= scale; shorter *= scale;
        }
        if (shorter == 0.0f) return tex_bilerp<RGB>(tv, 0, st);
        const float lod = fmaxf(0.0f, (float)n_levels - 1.0f + log2f(shorter));
        const int ilod = (int)(f2u_sat(floorf(lod)) & 0x7fffffffu);
        return tex_lerp(lod - (float)ilod, tex_ewa<RGB>(tv, ilod, st, dst0, dst1), tex_ewa<RGB>(tv, ilod + 1, st, dst0, dst1));
    }
    const float width = 2.0f * fmaxf(fmaxf(fmaxf(fabsf(dst0.x), fabsf(dst0.y)), fabsf(dst1.x)), fabsf(dst1.y));
    const float level = (float)n_levels - 1.0f + log2f(fmaxf(width, 1e-8f));
    if (level >= (float)n_levels - 1.0f) return tex_texel<RGB>(tv, n_levels - 1, 0, 0);
    const int il = max(0, f2i_sat(floorf(level)));
    if (tv.t.filter == SG_FILTER_POINT) {
        const SgImageLevel L = tv.level(il);
        return tex_texel<RGB>(tv, il, f2i_sat(roundf(st.x * (float)L.res[0] - 0.5f)), f2i_sat(roundf(st.y * (float)L.res[1] - 0.5f)));
    }
    if (tv.t.filter == SG_FILTER_BILINEAR) return tex_bilerp<RGB>(tv, il, st);
    if (il == 0) return tex_bilerp<RGB>(tv, 0, st);                                              // trilinear
    return tex_lerp(level - (float)il, tex_bilerp<RGB>(tv, il, st), tex_bilerp<RGB>(tv, il + 1, st));
}

// rgb2spec 0.1.1 RGB2Spec::fetch (third party; published rgb2spec.c `rgb2spec_fetch`)
SGD void rgb2spec_fetch(const DScene& sc, const float rgb_in[3], float out[3]) {
    const int res = (int)sc.rgb2spec_res;
    float rgb[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) rgb[i] = fmaxf(fminf(rgb_in[i], 1.0f), 0.0f);
    int i = 0;
    if (rgb[1] >= rgb[i]) i = 1;
    if (rgb[2] >= (i == 0 ? rgb[0] : rgb[1])) i = 2;
    const float z = i == 0 ? rgb[0] : (i == 1 ? rgb[1] : rgb[2]);
    const float cx = i == 0 ? rgb[1] : (i == 1 ? rgb[2] : rgb[0]);
    const float cy = i == 0 ? rgb[2] : (i == 1 ? rgb[0] : rgb[1]);
    const float scale = (float)(res - 1) / z, x = cx * scale, y = cy * scale;
    const uint32_t xi = min(f2u_sat(x), (uint32_t)(res - 2)), yi = min(f2u_sat(y), (uint32_t)(res - 2));
    uint32_t left = 0, last = (uint32_t)res - 2, size = last;
    while (size > 0) {
        const uint32_t half = size >> 1, middle = left + half + 1;
        if (__ldg(sc.rgb2spec_scale + middle) <= z) { left = middle; size -= half + 1; } else size = half;
    }
    const uint32_t zi = min(left, last);
    size_t offset = ((((size_t)i * res + zi) * res + yi) * res + xi) * 3;
    const size_t dx = 3, dy = 3 * (size_t)res, dz = 3 * (size_t)res * res;
    const float x1 = x - (float)xi, x0 = 1.0f - x1, y1 = y - (float)yi, y0 = 1.0f - y1;
    const float sz0 = __ldg(sc.rgb2spec_scale + zi), sz1 = __ldg(sc.rgb2spec_scale + zi + 1);
    const float z1 = (z - sz0) / (sz1 - sz0), z0 = 1.0f - z1;
    const float* T = sc.rgb2spec_data;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        out[j] = ((__ldg(T + offset) * x0 + __ldg(T + offset + dx) * x1) * y0 + (__ldg(T + offset + dy) * x0 + __ldg(T + offset + dy + dx) * x1) * y1) * z0 +
                 ((__ldg(T + offset + dz) * x0 + __ldg(T + offset + dz + dx) * x1) * y0 + (__ldg(T + offset + dz + dy) * x0 + __ldg(T + offset + dz + dy + dx) * x1) * y1) * z1;
        offset++;
    }
}
// RgbSigmoidPolynomial::get color.rs:352-383; poly(lambda, [c2, c1, c0]) = c2 + c1 x + c0 x^2 (Estrin: fused)
SGD float sigmoid_poly_get(const float c[3], float lambda) {
    const float x = fmaf(lambda * lambda, c[0], fmaf(lambda, c[1], c[2]));
    if (isinf(x)) return x > 0.0f ? 1.0f : 0.0f;
    return 0.5f + x / (2.0f * sqrtf(1.0f + x * x));
}

// TextureEvalContext (texture.rs): uv + screen-space derivatives.  `pdp` -> {p, dpdx, dpdy, n}: p / dpdx / dpdy for the non-UV
// mappings (texture.rs:938-1035), n for the direction-mix textures (:299,:813); null in scenes without SgTextureMapping or
// SgTextureNode rows, so the UV-only image-texture kernels carry nothing extra.
struct TexCoordCtx { float2 uv; float dudx, dudy, dvdx, dvdy; const float3* pdp = nullptr; };
SGD bool tex_needs_ctx(const DScene& sc) { return sc.texture_mappings != nullptr || sc.texture_nodes != nullptr; }

SGD void uv_map(const SgTexture& t, const TexCoordCtx& c, float2& st, float2& dst0, float2& dst1) {   // texture.rs:918-936
    const float dsdx = t.su * c.dudx, dsdy = t.su * c.dudy, dtdx = t.sv * c.dvdx, dtdy = t.sv * c.dvdy;
    st.x = t.su * c.uv.x + t.du; st.y = t.sv * c.uv.y + t.dv;
    st.y = 1.0f - st.y;
    dst0 = make_float2(dsdx, dtdx); dst1 = make_float2(dsdy, dtdy);
}
// SphericalMapping / CylindricalMapping / PlanarMapping ::map texture.rs:943-1035 (as written there), then the t flip of the
// image textures (:396-399, :780-781)
static __device__ __noinline__ void general_map(const DScene& sc, const SgTexture& t, const TexCoordCtx& c, float2& st, float2& dst0, float2& dst1) {
    const SgTextureMapping& M = sc.texture_mappings[t.mapping];
    const float3 p = c.pdp ? c.pdp[0] : f3(0.0f, 0.0f, 0.0f);
    const float3 pt = xform_point(M.texture_from_render, p);
    const float3 dpdx = xform_vector(M.texture_from_render, c.pdp ? c.pdp[1] : f3(0.0f, 0.0f, 0.0f));
    const float3 dpdy = xform_vector(M.texture_from_render, c.pdp ? c.pdp[2] : f3(0.0f, 0.0f, 0.0f));
    float3 dsdp, dtdp;
    if (M.kind == SG_MAPPING_SPHERICAL) {
        const float x2y2 = sqr(pt.x) + sqr(pt.y), sqrtx2y2 = sqrtf(x2y2);
        dsdp = f3(-pt.y, pt.x, 0.0f) / (2.0f * kPi * x2y2);
        dtdp = 1.0f / (kPi * (x2y2 + sqr(pt.z))) * f3(pt.x * pt.z / sqrtx2y2, pt.y * pt.z / sqrtx2y2, -sqrtx2y2);
        const float3 vec = normalize3(pt - f3(0.0f, 0.0f, 0.0f));
        const float theta = safe_asin(vec.z);                                   // spherical_theta = safe_acos, which calls asin (math.rs:272-274)
        st = make_float2(theta * kInvPi, theta * 0.15915494309189533577f);       // sic: both from theta (texture.rs:960-963)
    } else if (M.kind == SG_MAPPING_CYLINDRICAL) {
        const float x2y2 = sqr(pt.x) + sqr(pt.y);
        dsdp = f3(-pt.y, pt.x, 0.0f) / (2.0f * kPi * x2y2);
        dtdp = f3(0.0f, 0.0f, 1.0f);
        st = make_float2(kPi + atan2f(pt.y, pt.x) * 0.15915494309189533577f, pt.z);   // sic: texture.rs:990-993
    } else {
        dsdp = f3(M.vs[0], M.vs[1], M.vs[2]); dtdp = f3(M.vt[0], M.vt[1], M.vt[2]);
        st = make_float2(M.ds + dot3(pt, dsdp), M.dt + dot3(pt, dtdp));
    }
    dst0 = make_float2(dot3(dsdp, dpdx), dot3(dtdp, dpdx));                     // (dsdx, dtdx)
    dst1 = make_float2(dot3(dsdp, dpdy), dot3(dtdp, dpdy));                     // (dsdy, dtdy)
    st.y = 1.0f - st.y;
}
SGD void tex_map(const DScene& sc, const SgTexture& t, const TexCoordCtx& c, float2& st, float2& dst0, float2& dst1) {
    if (t.mapping < 0) uv_map(t, c, st, dst0, dst1); else general_map(sc, t, c, st, dst0, dst1);
}
// FloatImageTexture::evaluate texture.rs:393-404
static __device__ __noinline__ float eval_float_image(const DScene& sc, int tex, const TexCoordCtx& c) {
    const SgTexture t = sc.textures[tex];
    const TexView tv{sc, t};
    float2 st, d0, d1; tex_map(sc, t, c, st, d0, d1);
    const float v = tex_filter<false>(tv, st, d0, d1).x * t.scale;
    return t.invert ? fmaxf(0.0f, 1.0f - v) : v;
}
// SpectrumImageTexture::evaluate texture.rs:777-808
static __device__ __noinline__ Spec eval_spectrum_image(const DScene& sc, int tex, const TexCoordCtx& c, const Wavelengths& lam) {
    const SgTexture t = sc.textures[tex];
    const TexView tv{sc, t};
    float2 st, d0, d1; tex_map(sc, t, c, st, d0, d1);
    float3 rgb = tex_filter<true>(tv, st, d0, d1) * t.scale;
    if (t.invert) rgb = f3(1.0f - rgb.x, 1.0f - rgb.y, 1.0f - rgb.z);
    rgb = f3(fmaxf(0.0f, rgb.x), fmaxf(0.0f, rgb.y), fmaxf(0.0f, rgb.z));                        // clamp_zero
    if (t.n_channels != 3) return spec1(rgb.x);
    float in[3] = {rgb.x, rgb.y, rgb.z}, coef[3], scale = 1.0f;
    if (t.spectrum_type == SG_SPECTRUM_TYPE_UNBOUNDED) {                                         // spectrum.rs:534-546
        const float m = fmaxf(fmaxf(rgb.x, rgb.y), rgb.z);
        scale = 2.0f * m;
        if (scale != 0.0f) { in[0] = rgb.x / scale; in[1] = rgb.y / scale; in[2] = rgb.z / scale; } else { in[0] = in[1] = in[2] = 0.0f; }
    }
    rgb2spec_fetch(sc, in, coef);
    Spec s = make_float4(sigmoid_poly_get(coef, lam.lambda.x), sigmoid_poly_get(coef, lam.lambda.y), sigmoid_poly_get(coef, lam.lambda.z),
                         sigmoid_poly_get(coef, lam.lambda.w));
    if (t.spectrum_type == SG_SPECTRUM_TYPE_UNBOUNDED) s = scale * s;
    return s;
}

// The non-image members of `enum FloatTexture` / `enum SpectrumTexture` (texture.rs:88-94,411-417): constant, scaled, mix and
// direction-mix textures evaluate their operands recursively in the reference; here the recursion is unrolled at compile time to
// SG_MAX_TEXTURE_DEPTH levels (sg_scene_create rejects deeper trees), so no device stack frames of unknown size exist.
template <int DEPTH> struct TexTree {
    // `impl FloatTextureI for FloatTexture` texture.rs:142-152; members :175-179, :206-213, :246-261, :295-310
    static __device__ __noinline__ float eval_float(const DScene& sc, int tex, const TexCoordCtx& c) {
        const int kind = sc.textures[tex].kind;
        if (kind == SG_TEXTURE_IMAGE) return eval_float_image(sc, tex, c);
        const SgTextureNode nd = sc.texture_nodes[sc.textures[tex].node];
        if (kind == SG_TEXTURE_CONSTANT) return nd.value;
        if constexpr (DEPTH > 0) {
            using Sub = TexTree<DEPTH - 1>;
            if (kind == SG_TEXTURE_SCALED) {
                const float scl = Sub::eval_float(sc, nd.tex2, c);
                if (scl == 0.0f) return 0.0f;
                return Sub::eval_float(sc, nd.tex1, c) * scl;
            }
            float amt, t1 = 0.0f, t2 = 0.0f;
            if (kind == SG_TEXTURE_MIX) {
                amt = Sub::eval_float(sc, nd.amount, c);
                if (amt != 1.0f) t1 = Sub::eval_float(sc, nd.tex1, c);
                if (amt != 0.0f) t2 = Sub::eval_float(sc, nd.tex2, c);
                return t1 * (1.0f - amt) + t2 * amt;
            }
            amt = dot3(c.pdp ? c.pdp[3] : f3(0.0f, 0.0f, 0.0f), f3(nd.dir[0], nd.dir[1], nd.dir[2]));      // DirectionMix: 0 / 1 tests swapped w.r.t. Mix, as written
            if (amt != 0.0f) t1 = Sub::eval_float(sc, nd.tex1, c);
            if (amt != 1.0f) t2 = Sub::eval_float(sc, nd.tex2, c);
            return amt * t1 + (1.0f - amt) * t2;
        }
        return 0.0f;
    }
    // `impl SpectrumTextureI for SpectrumTexture` texture.rs:467-483; members :509-513, :567-583, :631-651, :810-826
    static __device__ __noinline__ Spec eval_spectrum(const DScene& sc, int tex, const TexCoordCtx& c, const Wavelengths& lam) {
        const int kind = sc.textures[tex].kind;
        if (kind == SG_TEXTURE_IMAGE) return eval_spectrum_image(sc, tex, c, lam);
        const SgTextureNode nd = sc.texture_nodes[sc.textures[tex].node];
        if (kind == SG_TEXTURE_CONSTANT) return nd.spectrum >= 0 ? spectrum_sample(sc, nd.spectrum, lam) : spec1(nd.value);
        if constexpr (DEPTH > 0) {
            using Sub = TexTree<DEPTH - 1>;
            if (kind == SG_TEXTURE_SCALED) {
                const float scl = Sub::eval_float(sc, nd.tex2, c);
                if (scl == 0.0f) return spec1(0.0f);
                return Sub::eval_spectrum(sc, nd.tex1, c, lam) * scl;
            }
            float amt; Spec t1 = spec1(0.0f), t2 = spec1(0.0f);
            if (kind == SG_TEXTURE_MIX) {
                amt = Sub::eval_float(sc, nd.amount, c);
                if (amt != 1.0f) t1 = Sub::eval_spectrum(sc, nd.tex1, c, lam);
                if (amt != 0.0f) t2 = Sub::eval_spectrum(sc, nd.tex2, c, lam);
                return t1 * (1.0f - amt) + t2 * amt;
            }
            amt = dot3(c.pdp ? c.pdp[3] : f3(0.0f, 0.0f, 0.0f), f3(nd.dir[0], nd.dir[1], nd.dir[2]));
            if (amt != 0.0f) t1 = Sub::eval_spectrum(sc, nd.tex1, c, lam);
            if (amt != 1.0f) t2 = Sub::eval_spectrum(sc, nd.tex2, c, lam);
            return amt * t1 + (1.0f - amt) * t2;
        }
        return spec1(0.0f);
    }
};
// Scenes without SgTextureNode rows (every BASELINE config) take the image path directly: one uniform branch.  (Measured alternative:
// a single out-of-line entry that dispatches inside -- the image-only call-site shape -- was 1.7 % SLOWER on C4, 183.7 vs 186.9 Mpaths/s.)
SGD float eval_float_texture(const DScene& sc, int tex, const TexCoordCtx& c) {
    if (sc.texture_nodes == nullptr) return eval_float_image(sc, tex, c);
    return TexTree<SG_MAX_TEXTURE_DEPTH>::eval_float(sc, tex, c);
}
SGD Spec eval_spectrum_texture(const DScene& sc, int tex, const TexCoordCtx& c, const Wavelengths& lam) {
    if (sc.texture_nodes == nullptr) return eval_spectrum_image(sc, tex, c, lam);
    return TexTree<SG_MAX_TEXTURE_DEPTH>::eval_spectrum(sc, tex, c, lam);
}

// Texture-valued material parameters (SgMaterialTextures): what the reference's materials read through tex_eval.evaluate_float /
// evaluate_spectrum (material.rs:456-499, 603-635, 917-963, 1188-1260).  The caller fills `v` with the SgMaterial constants; rows
// with a texture id overwrite them.  Out of line and only reached in scenes that carry the table.
struct MatTexValues { float ur, vr, thickness, g, ur2, vr2; Spec a, b, d; uint32_t mask; };      // mask: 1 = a, 2 = b, 4 = d hold texture values
static __device__ __noinline__ void resolve_material_textures(const DScene& sc, uint32_t material_id, const TexCoordCtx& tc, const Wavelengths& lam, MatTexValues& v) {
    const SgMaterialTextures mt = sc.material_textures[material_id];
    if (mt.u_roughness >= 0) v.ur = eval_float_texture(sc, mt.u_roughness, tc);
    if (mt.v_roughness >= 0) v.vr = eval_float_texture(sc, mt.v_roughness, tc);
    if (mt.thickness >= 0) v.thickness = eval_float_texture(sc, mt.thickness, tc);
    if (mt.g >= 0) v.g = eval_float_texture(sc, mt.g, tc);
    if (mt.u_roughness2 >= 0) v.ur2 = eval_float_texture(sc, mt.u_roughness2, tc);
    if (mt.v_roughness2 >= 0) v.vr2 = eval_float_texture(sc, mt.v_roughness2, tc);
    if (mt.spec_a >= 0) { v.a = eval_spectrum_texture(sc, mt.spec_a, tc, lam); v.mask |= 1u; }
    if (mt.spec_b >= 0) { v.b = eval_spectrum_texture(sc, mt.spec_b, tc, lam); v.mask |= 2u; }
    if (mt.spec_d >= 0) { v.d = eval_spectrum_texture(sc, mt.spec_d, tc, lam); v.mask |= 4u; }
}

// ---- screen-space differentials ----
SGD float3 xform_normal_t(const float* m, float3 n) {             // apply_normal_helper transform.rs:779-786
    return f3(m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z, m[2] * n.x + m[6] * n.y + m[10] * n.z);
}
SGD float3 mul3(const float* r, float3 v) { return f3(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z); }
SGD float3 mul3t(const float* r, float3 v) { return f3(r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z); }
SGD float3 ld3(const float* a) { return f3(a[0], a[1], a[2]); }

// Transform::rotate_from_to transform.rs:227-253 (3x3 part, row-major)
SGD void rotate_from_to(float3 from, float3 to, float r[9]) {
    float3 ref1;
    if (fabsf(from.x) < 0.72f && fabsf(to.x) < 0.72f) ref1 = f3(1.0f, 0.0f, 0.0f);
    else if (fabsf(from.y) < 0.72f && fabsf(to.y) < 0.72f) ref1 = f3(0.0f, 1.0f, 0.0f);
    else ref1 = f3(0.0f, 0.0f, 1.0f);
    const float3 u = ref1 - from, v = ref1 - to;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float kron = i == j ? 1.0f : 0.0f;
            r[3 * i + j] = kron - 2.0f / dot3(u, u) * comp3(u, i) * comp3(u, j) - 2.0f / dot3(v, v) * comp3(v, i) * comp3(v, j)
                           + 4.0f * dot3(u, v) / (dot3(u, u) * dot3(v, v)) * comp3(v, i) * comp3(u, j);
        }
}
// Camera::approximate_dp_dxy camera.rs:308-354
static __device__ __noinline__ void approximate_dp_dxy(const DScene& sc, float3 p, float3 n, int spp, uint32_t option_flags, float3& dpdx, float3& dpdy) {
    const SgCamera& cam = sc.camera;
    const float3 p_camera = xform_point(cam.camera_from_render, p);
    float r[9];
    rotate_from_to(normalize3(p_camera), f3(0.0f, 0.0f, 1.0f), r);
    const float3 p_down_z = f3(r[0] * p_camera.x + r[1] * p_camera.y + r[2] * p_camera.z + 0.0f, r[3] * p_camera.x + r[4] * p_camera.y + r[5] * p_camera.z + 0.0f,
                               r[6] * p_camera.x + r[7] * p_camera.y + r[8] * p_camera.z + 0.0f);
    const float3 n_down_z = mul3(r, xform_normal_t(cam.render_from_camera, n));
    const float d = n_down_z.z * p_down_z.z;
    const float3 xo = f3(0.0f, 0.0f, 0.0f) + ld3(cam.min_pos_differential_x), xd = f3(0.0f, 0.0f, 1.0f) + ld3(cam.min_dir_differential_x);
    const float tx = -(dot3(n_down_z, xo) - d) / dot3(n_down_z, xd);
    const float3 yo = f3(0.0f, 0.0f, 0.0f) + ld3(cam.min_pos_differential_y), yd = f3(0.0f, 0.0f, 1.0f) + ld3(cam.min_dir_differential_y);
    const float ty = -(dot3(n_down_z, yo) - d) / dot3(n_down_z, yd);
    const float3 px = xo + xd * tx, py = yo + yd * ty;
    const float spp_scale = (option_flags & SG_OPT_DISABLE_PIXEL_JITTER) ? 1.0f : fmaxf(0.125f, 1.0f / sqrtf((float)spp));
    dpdx = spp_scale * xform_vector(cam.render_from_camera, mul3t(r, px - p_down_z));
    dpdy = spp_scale * xform_vector(cam.render_from_camera, mul3t(r, py - p_down_z));
}

// Extra per-hit geometry the textured path needs (Surf keeps what every path needs).
struct SurfTex {
    float2 uv; float3 dpdu, dpdv;        // geometric parameterisation (interaction.rs:111-148)
    float3 dndu, dndv;                   // shading.dndu / dndv (triangle.rs:451-498)
    float dudx, dudy, dvdx, dvdy; float3 dpdx, dpdy;
};

template <> SGD void surf_tex_store<false>(SurfTex*, float2, float3, float3, float3, float3) {}
template <> SGD void surf_tex_store<true>(SurfTex* x, float2 uv, float3 dpdu, float3 dpdv, float3 dndu, float3 dndv) {
    x->uv = uv; x->dpdu = dpdu; x->dpdv = dpdv; x->dndu = dndu; x->dndv = dndv;
    x->dudx = x->dudy = x->dvdx = x->dvdy = 0.0f; x->dpdx = f3(0.0f, 0.0f, 0.0f); x->dpdy = f3(0.0f, 0.0f, 0.0f);
}

// Transform::apply(SurfaceInteraction) (transform.rs:573-609) for a hit inside an object instance (or a sphere's own
// object space).  `s` / `x` were built in that space; `rd` is the ray direction of the ENCLOSING space, so the interaction's
// wo is -(M^-1 rd) -- unless `wo_in` is given (nested case: a sphere / patch inside an instance already has its wo).  The reference maps vectors through M^-1 and normals through M^T
// (`t = self.inverse()`); with SG_SCENE_FIX_INSTANCING vectors go through M and normals through (M^-1)^T.  pi: forward
// Point3fi transform of an inexact point (:385-457).  Returns interaction.wo in `wo_si`.
template <bool TEX>
__device__ __noinline__ void transform_interaction(const DScene& sc, const float* M, const float* Mi, float3 rd, Surf& s, SurfTex* x, float3& wo_si,
                                                   const float3* wo_in = nullptr) {
    const bool fix = (sc.scene_flags & SG_SCENE_FIX_INSTANCING) != 0;
    const float* mv = fix ? M : Mi; const float* mn = fix ? Mi : M;
    auto vec = [&](float3 v) { return f3(mv[0] * v.x + mv[1] * v.y + mv[2] * v.z, mv[4] * v.x + mv[5] * v.y + mv[6] * v.z, mv[8] * v.x + mv[9] * v.y + mv[10] * v.z); };
    auto nrm = [&](float3 n) { return f3(mn[0] * n.x + mn[4] * n.y + mn[8] * n.z, mn[1] * n.x + mn[5] * n.y + mn[9] * n.z, mn[2] * n.x + mn[6] * n.y + mn[10] * n.z); };
    const float3 d2 = f3(Mi[0] * rd.x + Mi[1] * rd.y + Mi[2] * rd.z, Mi[4] * rd.x + Mi[5] * rd.y + Mi[6] * rd.z, Mi[8] * rd.x + Mi[9] * rd.y + Mi[10] * rd.z);
    wo_si = normalize3(vec(wo_in ? *wo_in : -d2));
    const float3 p = p3fi_mid(s.pi), e = p3fi_err(s.pi);
    const bool exact = s.pi.lo.x == s.pi.hi.x && s.pi.lo.y == s.pi.hi.y && s.pi.lo.z == s.pi.hi.z;
    float pp[3], ee[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        pp[r] = (M[4 * r] * p.x + M[4 * r + 1] * p.y) + (M[4 * r + 2] * p.z + M[4 * r + 3]);
        const float a = gamma_n(3) * (fabsf(M[4 * r] * p.x) + fabsf(M[4 * r + 1] * p.y) + fabsf(M[4 * r + 2] * p.z) + fabsf(M[4 * r + 3]));
        ee[r] = exact ? a : (gamma_n(3) + 1.0f) * (fabsf(M[4 * r]) * e.x + fabsf(M[4 * r + 1]) * e.y + fabsf(M[4 * r + 2]) * e.z) + a;
    }
    s.pi = p3fi_make(f3(pp[0], pp[1], pp[2]), f3(ee[0], ee[1], ee[2]));
    const float3 n = normalize3(nrm(s.n));
    s.n = n;
    s.sn = faceforward3(normalize3(nrm(s.sn)), n);
    s.sdpdu = vec(s.sdpdu); s.sdpdv = vec(s.sdpdv);
    if (TEX) { x->dpdu = vec(x->dpdu); x->dpdv = vec(x->dpdv); x->dndu = nrm(x->dndu); x->dndv = nrm(x->dndv); }
}

// SurfaceInteraction::compute_differentials interaction.rs:280-366
SGD void compute_differentials(const DScene& sc, const Surf& s, SurfTex& x, const AuxRays& aux, int spp, uint32_t option_flags) {
    if (option_flags & SG_OPT_DISABLE_TEXTURE_FILTERING) {
        x.dudx = x.dudy = x.dvdx = x.dvdy = 0.0f; x.dpdx = f3(0.0f, 0.0f, 0.0f); x.dpdy = f3(0.0f, 0.0f, 0.0f);
        return;
    }
    const float3 p = p3fi_mid(s.pi);
    if (aux.has && dot3(s.n, aux.rxd) != 0.0f && dot3(s.n, aux.ryd) != 0.0f) {
        const float d = -dot3(s.n, p);
        const float tx = (-dot3(s.n, aux.rxo) - d) / dot3(s.n, aux.rxd);
        const float3 px = aux.rxo + tx * aux.rxd;
        const float ty = (-dot3(s.n, aux.ryo) - d) / dot3(s.n, aux.ryd);
        const float3 py = aux.ryo + ty * aux.ryd;
        x.dpdx = px - p; x.dpdy = py - p;
    } else {
        approximate_dp_dxy(sc, p, s.n, spp, option_flags, x.dpdx, x.dpdy);
    }
    const float ata00 = dot3(x.dpdu, x.dpdu), ata01 = dot3(x.dpdu, x.dpdv), ata11 = dot3(x.dpdv, x.dpdv);
    float inv_det = 1.0f / dop(ata00, ata11, ata01, ata01);
    if (!isfinite(inv_det)) inv_det = 0.0f;
    const float atb0x = dot3(x.dpdu, x.dpdx), atb1x = dot3(x.dpdv, x.dpdx), atb0y = dot3(x.dpdu, x.dpdy), atb1y = dot3(x.dpdv, x.dpdy);
    x.dudx = dop(ata11, atb0x, ata01, atb1x) * inv_det;
    x.dvdx = dop(ata00, atb1x, ata01, atb0x) * inv_det;
    x.dudy = dop(ata11, atb0y, ata01, atb1y) * inv_det;
    x.dvdy = dop(ata00, atb1y, ata01, atb0y) * inv_det;
    x.dudx = isfinite(x.dudx) ? clampf(x.dudx, -1e8f, 1e8f) : 0.0f;
    x.dvdx = isfinite(x.dvdx) ? clampf(x.dvdx, -1e8f, 1e8f) : 0.0f;
    x.dudy = isfinite(x.dudy) ? clampf(x.dudy, -1e8f, 1e8f) : 0.0f;
    x.dvdy = isfinite(x.dvdy) ? clampf(x.dvdy, -1e8f, 1e8f) : 0.0f;
}

// bump_map material.rs:1477-1509 for a FloatImageTexture (tex >= 0) or the constant displacement `cdisp`; writes the
// displaced shading.dpdu / dpdv (the caller then rebuilds the shading normal, interaction.rs:229-250)
SGD void bump_map(const DScene& sc, int tex, float cdisp, Surf& s, const SurfTex& x) {
    float3 pdp[4];
    const bool mapped = tex_needs_ctx(sc);                                      // shifted_ctx.p only matters to the non-UV mappings, n to direction mixes
    if (mapped) { pdp[0] = p3fi_mid(s.pi); pdp[1] = x.dpdx; pdp[2] = x.dpdy; pdp[3] = s.n; }
    const TexCoordCtx c{x.uv, x.dudx, x.dudy, x.dvdx, x.dvdy, mapped ? pdp : nullptr};
    float du = 0.5f * (fabsf(x.dudx) + fabsf(x.dudy));
    if (du == 0.0f) du = 0.0005f;
    float dv = 0.5f * (fabsf(x.dvdx) + fabsf(x.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    float u_displace, v_displace, displace;
    if (tex >= 0) {
        TexCoordCtx cu = c; cu.uv = make_float2(x.uv.x + du, x.uv.y + 0.0f);
        TexCoordCtx cv = c; cv.uv = make_float2(x.uv.x + 0.0f, x.uv.y + dv);
        displace = eval_float_texture(sc, tex, c);
        const float3 p0 = mapped ? pdp[0] : f3(0.0f, 0.0f, 0.0f);
        if (mapped) pdp[0] = p0 + du * s.sdpdu;                                  // material.rs:1488
        u_displace = eval_float_texture(sc, tex, cu);
        if (mapped) pdp[0] = p0 + dv * s.sdpdv;                                  // :1499
        v_displace = eval_float_texture(sc, tex, cv);
    } else u_displace = v_displace = displace = cdisp;
    const float3 dpdu = s.sdpdu + (u_displace - displace) / du * s.sn + displace * x.dndu;
    const float3 dpdv = s.sdpdv + (v_displace - displace) / dv * s.sn + displace * x.dndv;
    s.sdpdu = dpdu; s.sdpdv = dpdv;
}

// normal_map material.rs:1453-1474: level 0 of a three-channel image through Image::bilerp_channel_wrapped with WrapMode::Repeat
static __device__ __noinline__ void normal_map(const DScene& sc, int tex, Surf& s, const SurfTex& x) {
    SgTexture t = sc.textures[tex]; t.wrap = SG_WRAP_REPEAT;
    const TexView tv{sc, t};
    const float3 px = tex_bilerp<true>(tv, 0, make_float2(x.uv.x, 1.0f - x.uv.y));
    float3 ns = normalize3(f3(2.0f * px.x - 1.0f, 2.0f * px.y - 1.0f, 2.0f * px.z - 1.0f));
    const float3 fx = normalize3(s.sdpdu), fz = s.sn, fy = cross3(fz, fx);      // Frame::from_xz frame.rs:14-17
    ns = ns.x * fx + ns.y * fy + ns.z * fz;
    const float ulen = len3(s.sdpdu), vlen = len3(s.sdpdv);
    const float3 dpdu = normalize3(gram_schmidt3(s.sdpdu, ns)) * ulen;
    s.sdpdu = dpdu; s.sdpdv = normalize3(cross3(ns, dpdu)) * vlen;
}

// SurfaceInteraction::spawn_ray_with_differentials interaction.rs:434-502 (auxiliary rays only)
SGD AuxRays spawn_differentials(const Surf& s, const SurfTex& x, const AuxRays& in, float3 wo, float3 wi, int bx_flags, float eta) {
    AuxRays out; out.has = false;
    out.rxo = out.rxd = out.ryo = out.ryd = f3(0.0f, 0.0f, 0.0f);
    if (!in.has) return out;
    float3 n = s.sn;
    float3 dndx = x.dndu * x.dudx + x.dndv * x.dvdx;
    float3 dndy = x.dndu * x.dudy + x.dndv * x.dvdy;
    const float3 dwodx = -in.rxd - wo, dwody = -in.ryd - wo;
    const float3 p = p3fi_mid(s.pi);
    if (bx_flags == (BX_SPECULAR | BX_REFLECTION)) {
        out.has = true;
        out.rxo = p + x.dpdx; out.ryo = p + x.dpdy;
        const float dwo_dotn_dx = dot3(dwodx, n) + dot3(wo, dndx);
        const float dwo_dotn_dy = dot3(dwody, n) + dot3(wo, dndy);
        out.rxd = wi - dwodx + 2.0f * (dot3(wo, n) * dndx + dwo_dotn_dx * n);
        out.ryd = wi - dwody + 2.0f * (dot3(wo, n) * dndy + dwo_dotn_dy * n);
    } else if (bx_flags == (BX_SPECULAR | BX_TRANSMISSION)) {
        out.has = true;
        out.rxo = p + x.dpdx; out.ryo = p + x.dpdy;
        if (dot3(wo, n) < 0.0f) { n = -n; dndx = -dndx; dndy = -dndy; }
        const float dwo_dotn_dx = dot3(dwodx, n) + dot3(wo, dndx);
        const float dwo_dotn_dy = dot3(dwody, n) + dot3(wo, dndy);
        const float mu = dot3(wo, n) / eta - absdot3(wi, n);
        const float dmudx = dwo_dotn_dx * (1.0f / eta + 1.0f / sqr(eta) * dot3(wo, n) / dot3(wi, n));
        const float dmudy = dwo_dotn_dy * (1.0f / eta + 1.0f / sqr(eta) * dot3(wo, n) / dot3(wi, n));
        out.rxd = wi - eta * dwodx + (mu * dndx + dmudx * n);
        out.ryd = wi - eta * dwody + (mu * dndy + dmudy * n);
    }
    if (out.has && (len2(out.rxd) > 1e16f || len2(out.ryd) > 1e16f || len2(out.rxo) > 1e16f || len2(out.ryo) > 1e16f)) out.has = false;
    return out;
}

}  // namespace sg
