// ORACLE -- TEST INFRASTRUCTURE ONLY.  ImageInfinitelight (light.rs:805-981) and what it stands on:
// PiecewiseConstant1D/2D sampling (sampling.rs:11-179), the equal-area square <-> sphere maps (math.rs:453-530, with the
// precedence slip `vp - up / r + 1.0` of :472 kept), Image::lookup_nearest_channel_wrapped with the OctahedralSphere wrap
// (image.rs:134-162, 590-601) and RgbIlluminantSpectrum (spectrum.rs:566-606).
// Third-party: fast_polynomial 0.1.0 `poly_array` with 7 coefficients (math.rs:514) is restated as the crate's Estrin scheme
// fma(x^4, fma(x^2, c6, fma(x, c5, c4)), fma(x^2, fma(x, c3, c2), fma(x, c1, c0))) -- parity unpinned (no reference test
// holds a vector for equal_area_sphere_to_square).
// Included from orc_shading.h after orc_texture.h (needs rgb2spec_fetch / sigmoid_poly_get).
#pragma once

namespace orc {

inline V3 equal_area_square_to_sphere(V2 p) {                                   // math.rs:453-484
    const Float u = 2.0f * p.x - 1.0f, v = 2.0f * p.y - 1.0f;
    const Float up = std::fabs(u), vp = std::fabs(v);
    const Float signed_distance = 1.0f - (up + vp);
    const Float d = std::fabs(signed_distance);
    const Float r = 1.0f - d;
    const Float phi = (r == 0.0f ? 1.0f : vp - up / r + 1.0f) * PI_F / 4.0f;     // sic: (vp - up) / r in pbrt
    const Float z = std::copysign(1.0f - sqr(r), signed_distance);
    const Float cos_phi = std::copysign(std::cos(phi), u), sin_phi = std::copysign(std::sin(phi), v);
    return v3(cos_phi * r * safe_sqrt(2.0f - sqr(r)), sin_phi * r * safe_sqrt(2.0f - sqr(r)), z);
}
inline Float poly7_estrin(Float x, const Float c[7]) {                          // fast_polynomial::poly_array, N = 7
    const Float x2 = x * x, x4 = x2 * x2;
    return std::fma(x4, std::fma(x2, c[6], std::fma(x, c[5], c[4])), std::fma(x2, std::fma(x, c[3], c[2]), std::fma(x, c[1], c[0])));
}
inline V2 equal_area_sphere_to_square(V3 d) {                                   // math.rs:487-530
    const Float x = std::fabs(d.x), y = std::fabs(d.y), z = std::fabs(d.z);
    const Float r = safe_sqrt(1.0f - z);
    const Float a = fmax_(x, y); Float b = fmin_(x, y);
    b = a == 0.0f ? 0.0f : b / a;
    static const Float T[7] = {0.406758566246788489601959989e-5f, 0.636226545274016134946890922156f, 0.61572017898280213493197203466e-2f,
                               -0.247333733281268944196501420480f, 0.881770664775316294736387951347e-1f, 0.419038818029165735901852432784e-1f,
                               -0.251390972343483509333252996350e-1f};
    Float phi = poly7_estrin(b, T);
    if (x < y) phi = 1.0f - phi;
    Float v = phi * r, u = r - v;
    if (d.z < 0.0f) { Float t = u; u = v; v = t; u = 1.0f - u; v = 1.0f - v; }
    u = std::copysign(u, d.x); v = std::copysign(v, d.y);
    V2 o; o.x = 0.5f * (u + 1.0f); o.y = 0.5f * (v + 1.0f);
    return o;
}

// PiecewiseConstant1D::sample sampling.rs:70-91 over `func[n]`, `cdf[n + 1]`, domain [0, 1]
inline Float pc1d_sample(const float* func, const float* cdf, int n, Float func_int, Float u, Float* pdf, int* offset) {
    const int o = find_interval(n + 1, [&](int i) { return cdf[i] <= u; });
    Float du = u - cdf[o];
    if (cdf[o + 1] - cdf[o] > 0.0f) du /= cdf[o + 1] - cdf[o];
    *pdf = func_int > 0.0f ? func[o] / func_int : 0.0f;
    *offset = o;
    return lerp(((Float)o + du) / (Float)n, 0.0f, 1.0f);
}
// PiecewiseConstant2D::sample sampling.rs:153-162
inline V2 pc2d_sample(const SgSceneDesc* D, const SgDistribution2D& d, V2 u, Float* pdf) {
    const float* P = D->spectrum_pool;
    Float pdf1, pdf0; int v, iu;
    const Float d1 = pc1d_sample(P + d.marg_func_off, P + d.marg_cdf_off, d.nv, d.marg_integral, u.y, &pdf1, &v);
    const Float d0 = pc1d_sample(P + d.func_off + (size_t)v * d.nu, P + d.cdf_off + (size_t)v * (d.nu + 1), d.nu, P[d.marg_func_off + v], u.x, &pdf0, &iu);
    *pdf = pdf0 * pdf1;
    V2 o; o.x = d0; o.y = d1; return o;
}
// Rust `as usize` of an f32: saturating, NaN -> 0
inline int f2usize_clamped(Float x, int hi) {
    if (!(x > 0.0f)) return 0;
    if (x >= (Float)hi) return hi;
    return (int)x;
}
// PiecewiseConstant2D::pdf sampling.rs:164-171 (domain.offset is the identity on [0,1]^2: (p - 0) / (1 - 0))
inline Float pc2d_pdf(const SgSceneDesc* D, const SgDistribution2D& d, V2 pr) {
    const float* P = D->spectrum_pool;
    const Float px = (pr.x - 0.0f) / (1.0f - 0.0f), py = (pr.y - 0.0f) / (1.0f - 0.0f);
    const int iu = f2usize_clamped(px * (Float)d.nu, d.nu - 1), iv = f2usize_clamped(py * (Float)d.nv, d.nv - 1);
    return P[d.func_off + (size_t)iv * d.nu + iu] / d.marg_integral;
}

// ImageInfinitelight::image_le light.rs:966-976
inline Spec env_image_le(const SgSceneDesc* D, const SgLight& lt, V2 uv, const Wavelengths& lambda) {
    const SgEnvMap& E = D->env_maps[lt.tri];
    // lookup_nearest_channel_wrapped image.rs:590-601: `(p * res) as i32` truncates toward zero (saturating)
    auto toi = [](Float x) { return x != x ? 0 : (x >= 2147483648.0f ? 2147483647 : (x <= -2147483648.0f ? (int)0x80000000 : (int)x)); };
    int px = toi(uv.x * (Float)E.res), py = toi(uv.y * (Float)E.res);
    const int R = E.res;                                                           // remap_pixel_coords, OctahedralSphere image.rs:135-162
    if (px < 0) { px = -px; py = R - 1 - py; } else if (px >= R) { px = 2 * R - 1 - px; py = R - 1 - py; }
    if (py < 0) { px = R - 1 - px; py = -py; } else if (py >= R) { px = R - 1 - px; py = 2 * R - 1 - py; }
    if (R == 1) { px = 0; py = 0; }
    const float* T = D->texels + E.texel_offset + ((size_t)py * R + px) * 3;
    Float rgb[3] = {fmax_(0.0f, T[0]), fmax_(0.0f, T[1]), fmax_(0.0f, T[2])};         // clamp_zero
    const Float m = fmax_(fmax_(rgb[0], rgb[1]), rgb[2]);                           // RgbIlluminantSpectrum::new spectrum.rs:573-587
    const Float scale = 2.0f * m;
    Float in[3] = {0.0f, 0.0f, 0.0f}, coef[3];
    if (scale != 0.0f) { in[0] = rgb[0] / scale; in[1] = rgb[1] / scale; in[2] = rgb[2] / scale; }
    rgb2spec_fetch(D, in, coef);
    Spec s;
    for (int i = 0; i < 4; ++i) s.v[i] = scale * sigmoid_poly_get(coef, lambda.lambda[i]);   // RgbIlluminantSpectrum::sample :599-605
    return lt.scale * (s * spectrum_sample(D, lt.spectrum, lambda));
}
// ImageInfinitelight::le light.rs:907-911
inline Spec env_le(const SgSceneDesc* D, const SgLight& lt, V3 ray_d, const Wavelengths& lambda) {
    const SgEnvMap& E = D->env_maps[lt.tri];
    const V3 wl = xform_vector3(E.light_from_render, ray_d);                        // apply_inverse(Vector3f)
    return env_image_le(D, lt, equal_area_sphere_to_square(wl), lambda);
}
// ImageInfinitelight::pdf_li light.rs:882-892
inline Float env_pdf_li(const SgSceneDesc* D, const SgLight& lt, V3 wi, bool allow_incomplete) {
    const SgEnvMap& E = D->env_maps[lt.tri];
    const V3 wl = xform_vector3(E.light_from_render, wi);
    const V2 uv = equal_area_sphere_to_square(wl);
    return pc2d_pdf(D, allow_incomplete ? E.compensated : E.distribution, uv) / (4.0f * PI_F);
}

}  // namespace orc
