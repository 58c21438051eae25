// ORACLE -- TEST INFRASTRUCTURE ONLY.  Image textures, MIP filtering, screen-space differentials.
// CPU restatement of: SpectrumImageTexture / FloatImageTexture::evaluate (texture.rs:393-404,777-808),
// UVMapping::map (:918-936), MIPMap::filter / ewa (mipmap.rs:121-293), Image::get_channel_wrapped /
// bilerp_channel_wrapped / remap_pixel_coords (image.rs:134-177,452-475,619-646), RgbAlbedoSpectrum /
// RgbUnboundedSpectrum (spectrum.rs:498-588), RgbSigmoidPolynomial::get (color.rs:352-383),
// SurfaceInteraction::compute_differentials (interaction.rs:280-366), Camera::approximate_dp_dxy
// (camera.rs:308-354), Transform::rotate_from_to (transform.rs:227-253), bump_map (material.rs:1477-1509),
// spawn_ray_with_differentials (interaction.rs:434-502).
//
// PARITY STATUS: third-party arithmetic restated from published algorithms => "parity unpinned" for
//   * rgb2spec 0.1.1 `RGB2Spec::fetch` (Jakob & Hanika 2019 reference implementation rgb2spec.c), and
//   * fast_polynomial 0.1.0 `poly` (Estrin evaluation of c2 + c1 x + c0 x^2, color.rs:359).
#pragma once
#include "orc_scene.h"

namespace orc {

struct AuxRays { bool has = false; V3 rxo, rxd, ryo, ryd; };

// ---- image access ---------------------------------------------------------------------------
inline int32_t f2i(Float f) {                    // Rust `as i32`: saturating, NaN -> 0
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
inline int32_t modulo_i(int32_t a, int32_t b) { int32_t r = a - (a / b) * b; return r < 0 ? r + b : r; }   // math.rs:439-451

struct TexView {
    const SgSceneDesc* D; const SgTexture* t;
    int levels() const { return t->n_levels; }
    const SgImageLevel& level(int l) const { return D->image_levels[t->first_level + l]; }
    // Image::get_channel_wrapped image.rs:452-475 with remap_pixel_coords :134-177
    Float channel(int l, int32_t x, int32_t y, int c) const {
        const SgImageLevel& L = level(l);
        int32_t p[2] = {x, y};
        for (int k = 0; k < 2; ++k) {
            if (p[k] >= 0 && p[k] < L.res[k]) continue;
            if (t->wrap == SG_WRAP_BLACK) return 0.0f;
            if (t->wrap == SG_WRAP_CLAMP) p[k] = p[k] < 0 ? 0 : (p[k] > L.res[k] - 1 ? L.res[k] - 1 : p[k]);
            else p[k] = modulo_i(p[k], L.res[k]);
        }
        return D->texels[(size_t)L.offset + ((size_t)p[1] * L.res[0] + p[0]) * t->n_channels + c];
    }
    // Image::bilerp_channel_wrapped image.rs:619-646
    Float bilerp_channel(int l, V2 st, int c) const {
        const SgImageLevel& L = level(l);
        Float x = st.x * (Float)L.res[0] - 0.5f, y = st.y * (Float)L.res[1] - 0.5f;
        int32_t xi = f2i(std::floor(x)), yi = f2i(std::floor(y));
        Float dx = x - (Float)xi, dy = y - (Float)yi;
        Float v0 = channel(l, xi, yi, c), v1 = channel(l, xi + 1, yi, c), v2 = channel(l, xi, yi + 1, c), v3 = channel(l, xi + 1, yi + 1, c);
        return (1.0f - dx) * (1.0f - dy) * v0 + dx * (1.0f - dy) * v1 + (1.0f - dx) * dy * v2 + dx * dy * v3;
    }
};

// A texel value: RGB (mipmap.rs texel_rgb :203-219) or Float in .r (texel_float :221-225).
struct Texel { Float r, g, b; };
inline Texel operator+(Texel a, Texel b) { Texel t = {a.r + b.r, a.g + b.g, a.b + b.b}; return t; }
inline Texel operator*(Texel a, Float s) { Texel t = {a.r * s, a.g * s, a.b * s}; return t; }
inline Texel operator/(Texel a, Float s) { Texel t = {a.r / s, a.g / s, a.b / s}; return t; }
inline Texel tex_lerp(Float t, Texel a, Texel b) { return a * (1.0f - t) + b * t; }                  // math.rs:246-252

template <bool RGB> inline Texel tex_texel(const TexView& tv, int l, int32_t x, int32_t y) {
    if (RGB && tv.t->n_channels == 3) { Texel t = {tv.channel(l, x, y, 0), tv.channel(l, x, y, 1), tv.channel(l, x, y, 2)}; return t; }
    Float v = tv.channel(l, x, y, 0); Texel t = {v, v, v}; return t;
}
template <bool RGB> inline Texel tex_bilerp(const TexView& tv, int l, V2 st) {                        // mipmap.rs:298-331
    if (RGB && tv.t->n_channels == 3) { Texel t = {tv.bilerp_channel(l, st, 0), tv.bilerp_channel(l, st, 1), tv.bilerp_channel(l, st, 2)}; return t; }
    Float v = tv.bilerp_channel(l, st, 0); Texel t = {v, v, v}; return t;
}
// TexelType::ewa mipmap.rs:233-293
template <bool RGB> inline Texel tex_ewa(const TexView& tv, int l, V2 st, V2 d0, V2 d1) {
    if (l >= tv.levels()) return tex_texel<RGB>(tv, tv.levels() - 1, 0, 0);
    const SgImageLevel& L = tv.level(l);
    st.x = st.x * (Float)L.res[0] - 0.5f; st.y = st.y * (Float)L.res[1] - 0.5f;
    d0.x *= (Float)L.res[0]; d0.y *= (Float)L.res[1]; d1.x *= (Float)L.res[0]; d1.y *= (Float)L.res[1];
    Float a = sqr(d0.y) + sqr(d1.y) + 1.0f;
    Float b = -2.0f * (d0.x * d0.y + d1.x * d1.y);
    Float c = sqr(d0.x) + sqr(d1.x) + 1.0f;
    Float inv_f = 1.0f / (a * c - sqr(b) * 0.25f);
    a *= inv_f; b *= inv_f; c *= inv_f;
    Float det = -sqr(b) + 4.0f * a * c;
    Float inv_det = 1.0f / det;
    Float u_sqrt = safe_sqrt(det * c), v_sqrt = safe_sqrt(a * det);
    int32_t s0 = f2i(std::ceil(st.x - 2.0f * inv_det * u_sqrt)), s1 = f2i(std::floor(st.x + 2.0f * inv_det * u_sqrt));
    int32_t t0 = f2i(std::ceil(st.y - 2.0f * inv_det * v_sqrt)), t1 = f2i(std::floor(st.y + 2.0f * inv_det * v_sqrt));
    Texel sum = {0.0f, 0.0f, 0.0f}; Float sum_wts = 0.0f;
    for (int32_t it = t0; it <= t1; ++it) {
        Float tt = (Float)it - st.y;
        for (int32_t is = s0; is <= s1; ++is) {
            Float ss = (Float)is - st.x;
            Float r2 = a * sqr(ss) + b * ss * tt + c * sqr(tt);
            if (r2 < 1.0f) {
                Float fi = r2 * 128.0f;                                        // `as usize` saturates
                uint32_t index = fi != fi || fi <= 0.0f ? 0u : (fi >= 4294967296.0f ? 0xffffffffu : (uint32_t)fi);
                if (index > 127u) index = 127u;
                Float w = tv.D->mip_filter_lut[index];
                sum = sum + tex_texel<RGB>(tv, l, is, it) * w;
                sum_wts += w;
            }
        }
    }
    return sum / sum_wts;
}
// MIPMap::filter mipmap.rs:121-201
template <bool RGB> inline Texel tex_filter(const TexView& tv, V2 st, V2 dst0, V2 dst1) {
    const int n_levels = tv.levels();
    if (tv.t->filter == SG_FILTER_EWA) {
        if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) std::swap(dst0, dst1);
        Float longer = std::sqrt(dst0.x * dst0.x + dst0.y * dst0.y);
        Float shorter = std::sqrt(dst1.x * dst1.x + dst1.y * dst1.y);
        if (shorter * tv.t->max_anisotropy < longer && shorter > 0.0f) {
            Float scale = longer / (shorter * tv.t->max_anisotropy);
            dst1.x *= scale; dst1.y *= scale; shorter *= scale;
        }
        if (shorter == 0.0f) return tex_bilerp<RGB>(tv, 0, st);
        Float lod = fmax_(0.0f, (Float)n_levels - 1.0f + std::log2(shorter));
        Float fl = std::floor(lod);
        int ilod = (int)(fl != fl || fl <= 0.0f ? 0u : (fl >= 4294967296.0f ? 0xffffffffu : (uint32_t)fl) & 0x7fffffffu);
        return tex_lerp(lod - (Float)ilod, tex_ewa<RGB>(tv, ilod, st, dst0, dst1), tex_ewa<RGB>(tv, ilod + 1, st, dst0, dst1));
    }
    Float width = 2.0f * fmax_(fmax_(fmax_(std::fabs(dst0.x), std::fabs(dst0.y)), std::fabs(dst1.x)), std::fabs(dst1.y));
    Float level = (Float)n_levels - 1.0f + std::log2(fmax_(width, 1e-8f));
    if (level >= (Float)n_levels - 1.0f) return tex_texel<RGB>(tv, n_levels - 1, 0, 0);
    int32_t il = f2i(std::floor(level)); if (il < 0) il = 0;
    if (tv.t->filter == SG_FILTER_POINT) {
        const SgImageLevel& L = tv.level(il);
        return tex_texel<RGB>(tv, il, f2i(std::round(st.x * (Float)L.res[0] - 0.5f)), f2i(std::round(st.y * (Float)L.res[1] - 0.5f)));
    }
    if (tv.t->filter == SG_FILTER_BILINEAR) return tex_bilerp<RGB>(tv, il, st);
    if (il == 0) return tex_bilerp<RGB>(tv, 0, st);                                            // trilinear
    return tex_lerp(level - (Float)il, tex_bilerp<RGB>(tv, il, st), tex_bilerp<RGB>(tv, il + 1, st));
}

// ---- RGB -> spectrum --------------------------------------------------------------------------
// rgb2spec 0.1.1 RGB2Spec::fetch (third party; restated from rgb2spec.c `rgb2spec_fetch`) -- parity unpinned
inline void rgb2spec_fetch(const SgSceneDesc* D, const Float rgb_in[3], Float out[3]) {
    const int res = (int)D->rgb2spec_res;
    Float rgb[3];
    for (int i = 0; i < 3; ++i) rgb[i] = fmax_(fmin_(rgb_in[i], 1.0f), 0.0f);
    int i = 0;
    for (int j = 1; j < 3; ++j) if (rgb[j] >= rgb[i]) i = j;
    Float z = rgb[i], scale = (Float)(res - 1) / z, x = rgb[(i + 1) % 3] * scale, y = rgb[(i + 2) % 3] * scale;
    auto tou = [](Float f) { return f != f || f <= 0.0f ? 0u : (f >= 4294967296.0f ? 0xffffffffu : (uint32_t)f); };
    uint32_t xi = std::min(tou(x), (uint32_t)(res - 2)), yi = std::min(tou(y), (uint32_t)(res - 2));
    // rgb2spec_find_interval: largest index with scale[idx] <= z, clamped to [0, res-2]
    uint32_t left = 0, last = (uint32_t)res - 2, size = last;
    while (size > 0) {
        uint32_t half = size >> 1, middle = left + half + 1;
        if (D->rgb2spec_scale[middle] <= z) { left = middle; size -= half + 1; } else size = half;
    }
    uint32_t zi = std::min(left, last);
    size_t offset = ((((size_t)i * res + zi) * res + yi) * res + xi) * 3, dx = 3, dy = 3 * (size_t)res, dz = 3 * (size_t)res * res;
    Float x1 = x - (Float)xi, x0 = 1.0f - x1, y1 = y - (Float)yi, y0 = 1.0f - y1;
    Float z1 = (z - D->rgb2spec_scale[zi]) / (D->rgb2spec_scale[zi + 1] - D->rgb2spec_scale[zi]), z0 = 1.0f - z1;
    const float* T = D->rgb2spec_data;
    for (int j = 0; j < 3; ++j) {
        out[j] = ((T[offset] * x0 + T[offset + dx] * x1) * y0 + (T[offset + dy] * x0 + T[offset + dy + dx] * x1) * y1) * z0 +
                 ((T[offset + dz] * x0 + T[offset + dz + dx] * x1) * y0 + (T[offset + dz + dy] * x0 + T[offset + dz + dy + dx] * x1) * y1) * z1;
        offset++;
    }
}
// RgbSigmoidPolynomial::get color.rs:352-383; poly(lambda, [c2, c1, c0]) = c2 + c1 x + c0 x^2 (fast_polynomial, Estrin)
inline Float sigmoid_poly_get(const Float c[3], Float lambda) {
    Float x = std::fma(lambda * lambda, c[0], std::fma(lambda, c[1], c[2]));
    if (std::isinf(x)) return x > 0.0f ? 1.0f : 0.0f;
    return 0.5f + x / (2.0f * std::sqrt(1.0f + x * x));
}

inline V3 xform_vector3(const float m[16], V3 v) {
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
inline V3 xform_normal_t(const float m[16], V3 n) {              // apply_normal_helper transform.rs:779-786 (transposed 3x3)
    return v3(m[0] * n.x + m[4] * n.y + m[8] * n.z, m[1] * n.x + m[5] * n.y + m[9] * n.z, m[2] * n.x + m[6] * n.y + m[10] * n.z);
}
inline V3 xform_point3(const float m[16], V3 p) {                // apply_point_helper transform.rs:753-767
    Float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    Float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    Float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    Float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1.0f) return v3(xp, yp, zp);
    return v3(xp, yp, zp) / wp;
}
// TextureEvalContext (texture.rs): uv + screen-space derivatives; p / dpdx / dpdy feed the non-UV mappings only
struct TexCoordCtx { V2 uv; Float dudx, dudy, dvdx, dvdy; V3 p = {0, 0, 0}, dpdx = {0, 0, 0}, dpdy = {0, 0, 0}, n = {0, 0, 0}; };

inline void uv_map(const SgTexture& t, const TexCoordCtx& c, V2* st, V2* dst0, V2* dst1) {            // texture.rs:918-936
    Float dsdx = t.su * c.dudx, dsdy = t.su * c.dudy, dtdx = t.sv * c.dvdx, dtdy = t.sv * c.dvdy;
    st->x = t.su * c.uv.x + t.du; st->y = t.sv * c.uv.y + t.dv;
    st->y = 1.0f - st->y;                                                                           // :396-399, :780-781
    dst0->x = dsdx; dst0->y = dtdx; dst1->x = dsdy; dst1->y = dtdy;
}
// SphericalMapping / CylindricalMapping / PlanarMapping ::map texture.rs:943-1035, then the t flip of the image textures
inline void tex_map(const SgSceneDesc* D, const SgTexture& t, const TexCoordCtx& c, V2* st, V2* dst0, V2* dst1) {
    if (t.mapping < 0) { uv_map(t, c, st, dst0, dst1); return; }
    const SgTextureMapping& M = D->texture_mappings[t.mapping];
    const V3 pt = xform_point3(M.texture_from_render, c.p);
    const V3 dpdx = xform_vector3(M.texture_from_render, c.dpdx), dpdy = xform_vector3(M.texture_from_render, c.dpdy);
    V3 dsdp, dtdp;
    if (M.kind == SG_MAPPING_SPHERICAL) {
        const Float x2y2 = sqr(pt.x) + sqr(pt.y), sqrtx2y2 = std::sqrt(x2y2);
        dsdp = v3(-pt.y, pt.x, 0.0f) / (2.0f * PI_F * x2y2);
        dtdp = 1.0f / (PI_F * (x2y2 + sqr(pt.z))) * v3(pt.x * pt.z / sqrtx2y2, pt.y * pt.z / sqrtx2y2, -sqrtx2y2);
        const V3 vec = normalize(pt - v3(0, 0, 0));
        const Float theta = safe_asin(vec.z);                                   // spherical_theta = safe_acos, which calls asin (math.rs:272-274)
        st->x = theta * INV_PI; st->y = theta * INV_2PI;                         // sic: both from theta (texture.rs:960-963)
    } else if (M.kind == SG_MAPPING_CYLINDRICAL) {
        const Float x2y2 = sqr(pt.x) + sqr(pt.y);
        dsdp = v3(-pt.y, pt.x, 0.0f) / (2.0f * PI_F * x2y2);
        dtdp = v3(0, 0, 1);
        st->x = PI_F + std::atan2(pt.y, pt.x) * INV_2PI; st->y = pt.z;          // sic: texture.rs:990-993
    } else {
        dsdp = v3(M.vs[0], M.vs[1], M.vs[2]); dtdp = v3(M.vt[0], M.vt[1], M.vt[2]);
        st->x = M.ds + dot(pt, dsdp); st->y = M.dt + dot(pt, dtdp);
    }
    dst0->x = dot(dsdp, dpdx); dst1->x = dot(dsdp, dpdy);                       // dsdx, dsdy
    dst0->y = dot(dtdp, dpdx); dst1->y = dot(dtdp, dpdy);                       // dtdx, dtdy
    st->y = 1.0f - st->y;                                                        // :396-399, :780-781
}
// FloatImageTexture::evaluate texture.rs:393-404
inline Float eval_float_image(const SgSceneDesc* D, int tex, const TexCoordCtx& c) {
    TexView tv = {D, &D->textures[tex]};
    V2 st, d0, d1; tex_map(D, *tv.t, c, &st, &d0, &d1);
    Float v = tex_filter<false>(tv, st, d0, d1).r * tv.t->scale;
    return tv.t->invert ? fmax_(0.0f, 1.0f - v) : v;
}
// SpectrumImageTexture::evaluate texture.rs:777-808
inline Spec eval_spectrum_image(const SgSceneDesc* D, int tex, const TexCoordCtx& c, const Wavelengths& lambda) {
    TexView tv = {D, &D->textures[tex]};
    V2 st, d0, d1; tex_map(D, *tv.t, c, &st, &d0, &d1);
    Texel rgb = tex_filter<true>(tv, st, d0, d1) * tv.t->scale;
    if (tv.t->invert) { rgb.r = 1.0f - rgb.r; rgb.g = 1.0f - rgb.g; rgb.b = 1.0f - rgb.b; }
    rgb.r = fmax_(0.0f, rgb.r); rgb.g = fmax_(0.0f, rgb.g); rgb.b = fmax_(0.0f, rgb.b);                 // clamp_zero
    if (tv.t->n_channels != 3) return spec_const(rgb.r);
    Float in[3] = {rgb.r, rgb.g, rgb.b}, coef[3], scale = 1.0f;
    if (tv.t->spectrum_type == SG_SPECTRUM_TYPE_UNBOUNDED) {                                        // spectrum.rs:534-546
        Float m = fmax_(fmax_(rgb.r, rgb.g), rgb.b);
        scale = 2.0f * m;
        if (scale != 0.0f) { in[0] = rgb.r / scale; in[1] = rgb.g / scale; in[2] = rgb.b / scale; } else { in[0] = in[1] = in[2] = 0.0f; }
    }
    rgb2spec_fetch(D, in, coef);
    Spec s;
    for (int i = 0; i < 4; ++i) s.v[i] = sigmoid_poly_get(coef, lambda.lambda[i]);
    if (tv.t->spectrum_type == SG_SPECTRUM_TYPE_UNBOUNDED) for (int i = 0; i < 4; ++i) s.v[i] = scale * s.v[i];
    return s;
}

// `impl FloatTextureI for FloatTexture` texture.rs:142-152 and its members :175-310
inline Float eval_float_texture(const SgSceneDesc* D, int tex, const TexCoordCtx& c) {
    const SgTexture& t = D->textures[tex];
    if (t.kind == SG_TEXTURE_IMAGE) return eval_float_image(D, tex, c);
    const SgTextureNode& nd = D->texture_nodes[t.node];
    switch (t.kind) {
    case SG_TEXTURE_CONSTANT: return nd.value;                                   // :175-179
    case SG_TEXTURE_SCALED: {                                                    // :206-213
        const Float sc = eval_float_texture(D, nd.tex2, c);
        if (sc == 0.0f) return 0.0f;
        return eval_float_texture(D, nd.tex1, c) * sc;
    }
    case SG_TEXTURE_MIX: {                                                       // :246-261
        const Float amt = eval_float_texture(D, nd.amount, c);
        Float t1 = 0.0f, t2 = 0.0f;
        if (amt != 1.0f) t1 = eval_float_texture(D, nd.tex1, c);
        if (amt != 0.0f) t2 = eval_float_texture(D, nd.tex2, c);
        return t1 * (1.0f - amt) + t2 * amt;
    }
    default: {                                                                   // DirectionMix :295-310 (note the tests are 0 / 1 swapped w.r.t. Mix)
        const Float amt = dot(c.n, v3(nd.dir[0], nd.dir[1], nd.dir[2]));
        Float t1 = 0.0f, t2 = 0.0f;
        if (amt != 0.0f) t1 = eval_float_texture(D, nd.tex1, c);
        if (amt != 1.0f) t2 = eval_float_texture(D, nd.tex2, c);
        return amt * t1 + (1.0f - amt) * t2;
    }
    }
}
// `impl SpectrumTextureI for SpectrumTexture` texture.rs:467-483 and its members :509-513,:567-583,:631-651,:810-826
inline Spec eval_spectrum_texture(const SgSceneDesc* D, int tex, const TexCoordCtx& c, const Wavelengths& lambda) {
    const SgTexture& t = D->textures[tex];
    if (t.kind == SG_TEXTURE_IMAGE) return eval_spectrum_image(D, tex, c, lambda);
    const SgTextureNode& nd = D->texture_nodes[t.node];
    switch (t.kind) {
    case SG_TEXTURE_CONSTANT: return nd.spectrum >= 0 ? spectrum_sample(D, nd.spectrum, lambda) : spec_const(nd.value);
    case SG_TEXTURE_SCALED: {
        const Float sc = eval_float_texture(D, nd.tex2, c);
        if (sc == 0.0f) return spec_const(0.0f);
        return eval_spectrum_texture(D, nd.tex1, c, lambda) * sc;
    }
    case SG_TEXTURE_MIX: {
        const Float amt = eval_float_texture(D, nd.amount, c);
        Spec t1 = spec_const(0.0f), t2 = spec_const(0.0f);
        if (amt != 1.0f) t1 = eval_spectrum_texture(D, nd.tex1, c, lambda);
        if (amt != 0.0f) t2 = eval_spectrum_texture(D, nd.tex2, c, lambda);
        return t1 * (1.0f - amt) + t2 * amt;
    }
    default: {
        const Float amt = dot(c.n, v3(nd.dir[0], nd.dir[1], nd.dir[2]));
        Spec t1 = spec_const(0.0f), t2 = spec_const(0.0f);
        if (amt != 0.0f) t1 = eval_spectrum_texture(D, nd.tex1, c, lambda);
        if (amt != 1.0f) t2 = eval_spectrum_texture(D, nd.tex2, c, lambda);
        return amt * t1 + (1.0f - amt) * t2;
    }
    }
}

// ---- screen-space differentials -------------------------------------------------------------------
// Transform::rotate_from_to transform.rs:227-253 (3x3 part, row-major)
inline void rotate_from_to(V3 from, V3 to, Float r[9]) {
    V3 ref1;
    if (std::fabs(from.x) < 0.72f && std::fabs(to.x) < 0.72f) ref1 = v3(1, 0, 0);
    else if (std::fabs(from.y) < 0.72f && std::fabs(to.y) < 0.72f) ref1 = v3(0, 1, 0);
    else ref1 = v3(0, 0, 1);
    V3 u = ref1 - from, v = ref1 - to;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        Float kron = i == j ? 1.0f : 0.0f;
        r[3 * i + j] = kron - 2.0f / dot(u, u) * comp(u, i) * comp(u, j) - 2.0f / dot(v, v) * comp(v, i) * comp(v, j)
                       + 4.0f * dot(u, v) / (dot(u, u) * dot(v, v)) * comp(v, i) * comp(u, j);
    }
}
inline V3 mul3(const Float r[9], V3 v) { return v3(r[0] * v.x + r[1] * v.y + r[2] * v.z, r[3] * v.x + r[4] * v.y + r[5] * v.z, r[6] * v.x + r[7] * v.y + r[8] * v.z); }
inline V3 mul3t(const Float r[9], V3 v) { return v3(r[0] * v.x + r[3] * v.y + r[6] * v.z, r[1] * v.x + r[4] * v.y + r[7] * v.z, r[2] * v.x + r[5] * v.y + r[8] * v.z); }

// BilinearPatch::interaction_from_intersection bilinear_patch.rs:238-425
inline SurfaceInteraction patch_interaction(const Scene& sc, uint32_t mesh_id, uint32_t patch, Float u, Float v, V3 wo) {
    const SgMesh& m = sc.d->meshes[mesh_id];
    uint32_t vi[4]; sc.patch_indices(mesh_id, patch, vi);
    V3 q[4]; sc.patch_points(mesh_id, patch, q);
    const V3 p00 = q[0], p10 = q[1], p01 = q[2], p11 = q[3];
    const V3 p = lerp3(u, lerp3(v, p00, p01), lerp3(v, p10, p11));
    V3 dpdu = lerp3(v, p10, p11) - lerp3(v, p00, p01);
    V3 dpdv = lerp3(u, p01, p11) - lerp3(u, p00, p10);
    V2 st = {u, v};
    Float duds = 1.0f, dudt = 0.0f, dvds = 0.0f, dvdt = 1.0f;
    auto lerp2 = [](Float t, V2 a, V2 b) { V2 r = {a.x * (1.0f - t) + b.x * t, a.y * (1.0f - t) + b.y * t}; return r; };
    if (m.flags & SG_MESH_HAS_UV) {
        const V2 uv00 = sc.uv(m, vi[0]), uv10 = sc.uv(m, vi[1]), uv01 = sc.uv(m, vi[2]), uv11 = sc.uv(m, vi[3]);
        st = lerp2(u, lerp2(v, uv00, uv01), lerp2(v, uv10, uv11));
        const V2 a1 = lerp2(v, uv10, uv11), a0 = lerp2(v, uv00, uv01), b1 = lerp2(u, uv01, uv11), b0 = lerp2(u, uv00, uv10);
        const V2 dstdu = {a1.x - a0.x, a1.y - a0.y}, dstdv = {b1.x - b0.x, b1.y - b0.y};
        duds = std::fabs(dstdu.x) < 1e-8f ? 0.0f : 1.0f / dstdu.x;
        dvds = std::fabs(dstdv.x) < 1e-8f ? 0.0f : 1.0f / dstdv.x;
        dudt = std::fabs(dstdu.y) < 1e-8f ? 0.0f : 1.0f / dstdu.y;
        dvdt = std::fabs(dstdv.y) < 1e-8f ? 0.0f : 1.0f / dstdv.y;
        const V3 dpds = dpdu * duds + dpdv * dvds;
        V3 dpdt = dpdu * dudt + dpdv * dvdt;
        const V3 cx = cross(dpds, dpdt);
        if (!(cx.x == 0.0f && cx.y == 0.0f && cx.z == 0.0f)) {
            if (dot(cross(dpdu, dpdv), cross(dpds, dpdt)) < 0.0f) dpdt = -dpdt;
            dpdu = dpds; dpdv = dpdt;
        }
    }
    const V3 d2pduu = v3(0, 0, 0), d2pdvv = v3(0, 0, 0), d2pduv = (p00 - p01) + (p11 - p10);
    const Float e1 = dot(dpdu, dpdu), f1 = dot(dpdu, dpdv), g1 = dot(dpdv, dpdv);
    const V3 n = normalize(cross(dpdu, dpdv));
    const Float e2 = dot(n, d2pduu), f2 = dot(n, d2pduv), g2 = dot(n, d2pdvv);
    const Float egf2 = difference_of_products(e1, g1, f1, f1);
    const Float inv = egf2 != 0.0f ? 1.0f / egf2 : 0.0f;
    V3 dndu = ((f1 * f2 - e2 * g1) * inv) * dpdu + ((e2 * f1 - f2 * e1) * inv) * dpdv;
    V3 dndv = ((g2 * f1 - f2 * g1) * inv) * dpdu + ((f2 * f1 - g2 * e1) * inv) * dpdv;
    const V3 dnds = dndu * duds + dndv * dvds, dndt = dndu * dudt + dndv * dvdt;
    dndu = dnds; dndv = dndt;
    const V3 p_abs_sum = vabs(p00) + vabs(p01) + vabs(p10) + vabs(p11);
    const bool flip = ((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    SurfaceInteraction si;                                           // SurfaceInteraction::new_with_face_index interaction.rs:111-176
    si.pi = p3fi_from_value_and_error(p, gamma_n(6) * p_abs_sum);
    si.uv = st; si.wo = wo; si.dpdu = dpdu; si.dpdv = dpdv;
    si.n = flip ? -n : n;
    si.sn = si.n; si.sdpdu = dpdu; si.sdpdv = dpdv; si.sdndu = dndu; si.sdndv = dndv;
    si.material = -1; si.light = -1;
    if (m.flags & SG_MESH_HAS_N) {
        const V3 n00 = sc.normal(m, vi[0]), n10 = sc.normal(m, vi[1]), n01 = sc.normal(m, vi[2]), n11 = sc.normal(m, vi[3]);
        V3 ns = lerp3(u, lerp3(v, n00, n01), lerp3(v, n10, n11));
        if (length_squared(ns) > 0.0f) {
            ns = normalize(ns);
            V3 sdndu = lerp3(v, n10, n11) - lerp3(v, n00, n01);
            V3 sdndv = lerp3(u, n01, n11) - lerp3(u, n00, n10);
            const V3 s_dnds = sdndu * duds + sdndv * dvds, s_dndt = sdndu * dudt + sdndv * dvdt;
            Float r[9]; rotate_from_to(si.n, ns, r);
            // set_shading_geometry(ns, r(dpdu), r(dpdv), dnds, dndt, true) interaction.rs:379-405
            si.sn = ns;
            si.n = face_forward(si.n, si.sn);
            si.sdpdu = mul3(r, dpdu); si.sdpdv = mul3(r, dpdv); si.sdndu = s_dnds; si.sdndv = s_dndt;
            while (length_squared(si.sdpdu) > 1e16f || length_squared(si.sdpdv) > 1e16f) { si.sdpdu = si.sdpdu / 1e8f; si.sdpdv = si.sdpdv / 1e8f; }
        }
    }
    return si;
}

// Camera::approximate_dp_dxy camera.rs:308-354
inline void approximate_dp_dxy(const SgCamera& cam, V3 p, V3 n, int spp, uint32_t option_flags, V3* dpdx, V3* dpdy) {
    V3 p_camera = xform_point3(cam.camera_from_render, p);
    Float r[9];
    rotate_from_to(normalize(p_camera), v3(0, 0, 1), r);
    // Transform::apply(Point3f) through the 4x4 helper: w = 1 exactly, translation 0
    V3 p_down_z = v3(r[0] * p_camera.x + r[1] * p_camera.y + r[2] * p_camera.z + 0.0f, r[3] * p_camera.x + r[4] * p_camera.y + r[5] * p_camera.z + 0.0f,
                     r[6] * p_camera.x + r[7] * p_camera.y + r[8] * p_camera.z + 0.0f);
    // camera_from_render_n = render_from_camera.apply_inverse(Normal) = apply_normal_helper(m = render_from_camera) (transform.rs:623-629)
    V3 n_cam = xform_normal_t(cam.render_from_camera, n);
    // down_z.apply(Normal) = apply_normal_helper(m_inv = r^T) -> transpose of the transpose = r
    V3 n_down_z = mul3(r, n_cam);
    Float d = n_down_z.z * p_down_z.z;
    V3 xo = v3(0, 0, 0) + v3(cam.min_pos_differential_x[0], cam.min_pos_differential_x[1], cam.min_pos_differential_x[2]);
    V3 xd = v3(0, 0, 1) + v3(cam.min_dir_differential_x[0], cam.min_dir_differential_x[1], cam.min_dir_differential_x[2]);
    Float tx = -(dot(n_down_z, xo) - d) / dot(n_down_z, xd);
    V3 yo = v3(0, 0, 0) + v3(cam.min_pos_differential_y[0], cam.min_pos_differential_y[1], cam.min_pos_differential_y[2]);
    V3 yd = v3(0, 0, 1) + v3(cam.min_dir_differential_y[0], cam.min_dir_differential_y[1], cam.min_dir_differential_y[2]);
    Float ty = -(dot(n_down_z, yo) - d) / dot(n_down_z, yd);
    V3 px = xo + xd * tx, py = yo + yd * ty;
    Float spp_scale = (option_flags & SG_OPT_DISABLE_PIXEL_JITTER) ? 1.0f : fmax_(0.125f, 1.0f / std::sqrt((Float)spp));
    *dpdx = spp_scale * xform_vector3(cam.render_from_camera, mul3t(r, px - p_down_z));
    *dpdy = spp_scale * xform_vector3(cam.render_from_camera, mul3t(r, py - p_down_z));
}

// SurfaceInteraction::compute_differentials interaction.rs:280-366
inline void compute_differentials(const SgSceneDesc* D, SurfaceInteraction& si, const AuxRays& aux, int spp, uint32_t option_flags) {
    if (option_flags & SG_OPT_DISABLE_TEXTURE_FILTERING) {
        si.dudx = si.dudy = si.dvdx = si.dvdy = 0.0f; si.dpdx = v3(0, 0, 0); si.dpdy = v3(0, 0, 0);
        return;
    }
    V3 p = si.p();
    if (aux.has && dot(si.n, aux.rxd) != 0.0f && dot(si.n, aux.ryd) != 0.0f) {
        Float d = -dot(si.n, p);
        Float tx = (-dot(si.n, aux.rxo) - d) / dot(si.n, aux.rxd);
        V3 px = aux.rxo + tx * aux.rxd;
        Float ty = (-dot(si.n, aux.ryo) - d) / dot(si.n, aux.ryd);
        V3 py = aux.ryo + ty * aux.ryd;
        si.dpdx = px - p; si.dpdy = py - p;
    } else {
        approximate_dp_dxy(D->camera, p, si.n, spp, option_flags, &si.dpdx, &si.dpdy);
    }
    Float ata00 = dot(si.dpdu, si.dpdu), ata01 = dot(si.dpdu, si.dpdv), ata11 = dot(si.dpdv, si.dpdv);
    Float inv_det = 1.0f / difference_of_products(ata00, ata11, ata01, ata01);
    if (!std::isfinite(inv_det)) inv_det = 0.0f;
    Float atb0x = dot(si.dpdu, si.dpdx), atb1x = dot(si.dpdv, si.dpdx), atb0y = dot(si.dpdu, si.dpdy), atb1y = dot(si.dpdv, si.dpdy);
    si.dudx = difference_of_products(ata11, atb0x, ata01, atb1x) * inv_det;
    si.dvdx = difference_of_products(ata00, atb1x, ata01, atb0x) * inv_det;
    si.dudy = difference_of_products(ata11, atb0y, ata01, atb1y) * inv_det;
    si.dvdy = difference_of_products(ata00, atb1y, ata01, atb0y) * inv_det;
    auto fix = [](Float v) { return std::isfinite(v) ? clampf(v, -1e8f, 1e8f) : 0.0f; };
    si.dudx = fix(si.dudx); si.dvdx = fix(si.dvdx); si.dudy = fix(si.dudy); si.dvdy = fix(si.dvdy);
}

// bump_map material.rs:1477-1509 for a FloatImageTexture (tex >= 0) or the constant displacement `cdisp`
inline void bump_map(const SgSceneDesc* D, int tex, Float cdisp, const SurfaceInteraction& si, V3* dpdu_out, V3* dpdv_out) {
    TexCoordCtx c = {si.uv, si.dudx, si.dudy, si.dvdx, si.dvdy, si.p(), si.dpdx, si.dpdy, si.n};
    Float du = 0.5f * (std::fabs(si.dudx) + std::fabs(si.dudy));
    if (du == 0.0f) du = 0.0005f;
    Float dv = 0.5f * (std::fabs(si.dvdx) + std::fabs(si.dvdy));
    if (dv == 0.0f) dv = 0.0005f;
    Float u_displace, v_displace, displace;
    if (tex >= 0) {
        TexCoordCtx cu = c; cu.uv.x = si.uv.x + du; cu.uv.y = si.uv.y + 0.0f; cu.p = si.p() + du * si.sdpdu;
        TexCoordCtx cv = c; cv.uv.x = si.uv.x + 0.0f; cv.uv.y = si.uv.y + dv; cv.p = si.p() + dv * si.sdpdv;
        u_displace = eval_float_texture(D, tex, cu); v_displace = eval_float_texture(D, tex, cv); displace = eval_float_texture(D, tex, c);
    } else u_displace = v_displace = displace = cdisp;
    *dpdu_out = si.sdpdu + (u_displace - displace) / du * si.sn + displace * si.sdndu;
    *dpdv_out = si.sdpdv + (v_displace - displace) / dv * si.sn + displace * si.sdndv;
}

// normal_map material.rs:1453-1474: level 0 of a three-channel image, Image::bilerp_channel_wrapped with WrapMode::Repeat
inline void normal_map(const SgSceneDesc* D, int tex, const SurfaceInteraction& si, V3* dpdu_out, V3* dpdv_out) {
    SgTexture t = D->textures[tex]; t.wrap = SG_WRAP_REPEAT;
    TexView tv = {D, &t};
    V2 uv; uv.x = si.uv.x; uv.y = 1.0f - si.uv.y;
    const Texel px = tex_bilerp<true>(tv, 0, uv);
    V3 ns = normalize(v3(2.0f * px.r - 1.0f, 2.0f * px.g - 1.0f, 2.0f * px.b - 1.0f));
    const V3 fx = normalize(si.sdpdu), fz = si.sn, fy = cross(fz, fx);            // Frame::from_xz frame.rs:14-17
    ns = ns.x * fx + ns.y * fy + ns.z * fz;
    const Float ulen = length(si.sdpdu), vlen = length(si.sdpdv);
    const V3 dpdu = normalize(gram_schmidt(si.sdpdu, ns)) * ulen;
    *dpdu_out = dpdu; *dpdv_out = normalize(cross(ns, dpdu)) * vlen;
}

// SurfaceInteraction::spawn_ray_with_differentials interaction.rs:434-502 (auxiliary part)
inline AuxRays spawn_differentials(const SurfaceInteraction& si, const AuxRays& in, V3 wi, int bx_flags, Float eta) {
    AuxRays out;
    if (!in.has) return out;
    V3 n = si.sn;
    V3 dndx = si.sdndu * si.dudx + si.sdndv * si.dvdx;
    V3 dndy = si.sdndu * si.dudy + si.sdndv * si.dvdy;
    V3 dwodx = -in.rxd - si.wo, dwody = -in.ryd - si.wo;
    if (bx_flags == (BX_SPECULAR | BX_REFLECTION)) {
        out.has = true;
        out.rxo = si.p() + si.dpdx; out.ryo = si.p() + si.dpdy;
        Float dwo_dotn_dx = dot(dwodx, n) + dot(si.wo, dndx);
        Float dwo_dotn_dy = dot(dwody, n) + dot(si.wo, dndy);
        out.rxd = wi - dwodx + 2.0f * (dot(si.wo, n) * dndx + dwo_dotn_dx * n);
        out.ryd = wi - dwody + 2.0f * (dot(si.wo, n) * dndy + dwo_dotn_dy * n);
    } else if (bx_flags == (BX_SPECULAR | BX_TRANSMISSION)) {
        out.has = true;
        out.rxo = si.p() + si.dpdx; out.ryo = si.p() + si.dpdy;
        if (dot(si.wo, n) < 0.0f) { n = -n; dndx = -dndx; dndy = -dndy; }
        Float dwo_dotn_dx = dot(dwodx, n) + dot(si.wo, dndx);
        Float dwo_dotn_dy = dot(dwody, n) + dot(si.wo, dndy);
        Float mu = dot(si.wo, n) / eta - abs_dot(wi, n);
        Float dmudx = dwo_dotn_dx * (1.0f / eta + 1.0f / sqr(eta) * dot(si.wo, n) / dot(wi, n));
        Float dmudy = dwo_dotn_dy * (1.0f / eta + 1.0f / sqr(eta) * dot(si.wo, n) / dot(wi, n));
        out.rxd = wi - eta * dwodx + (mu * dndx + dmudx * n);
        out.ryd = wi - eta * dwody + (mu * dndy + dmudy * n);
    }
    if (out.has && (length_squared(out.rxd) > 1e16f || length_squared(out.ryd) > 1e16f || length_squared(out.rxo) > 1e16f || length_squared(out.ryo) > 1e16f))
        out.has = false;
    return out;
}

}  // namespace orc
