// ImageInfinitelight on the device (light.rs:805-981): PiecewiseConstant1D/2D sampling (sampling.rs:11-179), the equal-area
// square <-> sphere maps (math.rs:453-530; the precedence slip `vp - up / r + 1.0` of :472 is kept), the nearest-texel lookup
// with the OctahedralSphere wrap (image.rs:134-162, 590-601) and RgbIlluminantSpectrum (spectrum.rs:566-606).
// fast_polynomial's 7-coefficient `poly_array` (math.rs:514) is restated as the crate's Estrin scheme, like the oracle.
// Out of line on purpose: only scenes with an environment map ever call these, and the shade kernels are sensitive to
// instruction-cache footprint.  Included after sg_texture.cuh (rgb2spec_fetch, sigmoid_poly_get).
#pragma once

namespace sg {

SGD float3 equal_area_square_to_sphere(float2 p) {
    const float u = 2.0f * p.x - 1.0f, v = 2.0f * p.y - 1.0f;
    const float up = fabsf(u), vp = fabsf(v);
    const float signed_distance = 1.0f - (up + vp);
    const float d = fabsf(signed_distance);
    const float r = 1.0f - d;
    const float phi = (r == 0.0f ? 1.0f : vp - up / r + 1.0f) * kPi / 4.0f;           // sic
    const float z = copysignf(1.0f - sqr(r), signed_distance);
    const float cos_phi = copysignf(cosf(phi), u), sin_phi = copysignf(sinf(phi), v);
    return f3(cos_phi * r * safe_sqrt(2.0f - sqr(r)), sin_phi * r * safe_sqrt(2.0f - sqr(r)), z);
}
SGD float2 equal_area_sphere_to_square(float3 d) {
    const float x = fabsf(d.x), y = fabsf(d.y), z = fabsf(d.z);
    const float r = safe_sqrt(1.0f - z);
    const float a = fmaxf(x, y); float b = fminf(x, y);
    b = a == 0.0f ? 0.0f : b / a;
    const float t1 = 0.406758566246788489601959989e-5f, t2 = 0.636226545274016134946890922156f, t3 = 0.61572017898280213493197203466e-2f,
                t4 = -0.247333733281268944196501420480f, t5 = 0.881770664775316294736387951347e-1f, t6 = 0.419038818029165735901852432784e-1f,
                t7 = -0.251390972343483509333252996350e-1f;
    const float b2 = b * b, b4 = b2 * b2;
    float phi = fmaf(b4, fmaf(b2, t7, fmaf(b, t6, t5)), fmaf(b2, fmaf(b, t4, t3), fmaf(b, t2, t1)));
    if (x < y) phi = 1.0f - phi;
    float v = phi * r, u = r - v;
    if (d.z < 0.0f) { const float t = u; u = v; v = t; u = 1.0f - u; v = 1.0f - v; }
    u = copysignf(u, d.x); v = copysignf(v, d.y);
    return make_float2(0.5f * (u + 1.0f), 0.5f * (v + 1.0f));
}
// PiecewiseConstant1D::sample sampling.rs:70-91 (find_interval over the n + 1 cdf entries, math.rs:322-333)
SGD float pc1d_sample(const float* func, const float* cdf, int n, float func_int, float u, float& pdf, int& offset) {
    const int o = find_interval_le(cdf, n + 1, u);
    const float c0 = __ldg(cdf + o), c1 = __ldg(cdf + o + 1);
    float du = u - c0;
    if (c1 - c0 > 0.0f) du /= c1 - c0;
    pdf = func_int > 0.0f ? __ldg(func + o) / func_int : 0.0f;
    offset = o;
    return lerpf(((float)o + du) / (float)n, 0.0f, 1.0f);
}
SGD float2 pc2d_sample(const DScene& sc, const SgDistribution2D& d, float2 u, float& pdf) {       // sampling.rs:153-162
    const float* P = sc.pool;
    float pdf1, pdf0; int v, iu;
    const float d1 = pc1d_sample(P + d.marg_func_off, P + d.marg_cdf_off, d.nv, d.marg_integral, u.y, pdf1, v);
    const float d0 = pc1d_sample(P + d.func_off + (size_t)v * d.nu, P + d.cdf_off + (size_t)v * (d.nu + 1), d.nu, __ldg(P + d.marg_func_off + v), u.x, pdf0, iu);
    pdf = pdf0 * pdf1;
    return make_float2(d0, d1);
}
SGD int f2usize_clamped(float x, int hi) {           // Rust `as usize` (saturating, NaN -> 0) then .clamp(0, hi)
    if (!(x > 0.0f)) return 0;
    if (x >= (float)hi) return hi;
    return (int)x;
}
SGD float pc2d_pdf(const DScene& sc, const SgDistribution2D& d, float2 pr) {                        // sampling.rs:164-171
    const float px = (pr.x - 0.0f) / (1.0f - 0.0f), py = (pr.y - 0.0f) / (1.0f - 0.0f);
    const int iu = f2usize_clamped(px * (float)d.nu, d.nu - 1), iv = f2usize_clamped(py * (float)d.nv, d.nv - 1);
    return __ldg(sc.pool + d.func_off + (size_t)iv * d.nu + iu) / d.marg_integral;
}
// ImageInfinitelight::image_le light.rs:966-976
static __device__ __noinline__ Spec env_image_le(const DScene& sc, const SgLight& lt, float2 uv, const Wavelengths& lam) {
    const SgEnvMap& E = sc.env_maps[lt.tri];
    const int R = E.res;
    int px = f2i_sat(uv.x * (float)R), py = f2i_sat(uv.y * (float)R);                                // `as i32`: truncation, saturating, NaN -> 0
    if (px < 0) { px = -px; py = R - 1 - py; } else if (px >= R) { px = 2 * R - 1 - px; py = R - 1 - py; }
    if (py < 0) { px = R - 1 - px; py = -py; } else if (py >= R) { px = R - 1 - px; py = 2 * R - 1 - py; }
    if (R == 1) { px = 0; py = 0; }
    const float* T = sc.texels + E.texel_offset + ((size_t)py * R + px) * 3;
    const float rgb[3] = {fmaxf(0.0f, __ldg(T)), fmaxf(0.0f, __ldg(T + 1)), fmaxf(0.0f, __ldg(T + 2))};
    const float m = fmaxf(fmaxf(rgb[0], rgb[1]), rgb[2]);
    const float scale = 2.0f * m;
    float in[3] = {0.0f, 0.0f, 0.0f}, coef[3];
    if (scale != 0.0f) { in[0] = rgb[0] / scale; in[1] = rgb[1] / scale; in[2] = rgb[2] / scale; }
    rgb2spec_fetch(sc, in, coef);
    const Spec s = make_float4(scale * sigmoid_poly_get(coef, lam.lambda.x), scale * sigmoid_poly_get(coef, lam.lambda.y),
                               scale * sigmoid_poly_get(coef, lam.lambda.z), scale * sigmoid_poly_get(coef, lam.lambda.w));
    return lt.scale * (s * spectrum_sample(sc, lt.spectrum, lam));
}
// Light::le of the infinite lights: light.rs:792-794 (uniform), :907-911 (image)
static __device__ __noinline__ Spec infinite_le(const DScene& sc, const SgLight& lt, float3 ray_d, const Wavelengths& lam) {
    if (lt.kind == SG_LIGHT_IMAGE_INFINITE) {
        const float3 wl = xform_vector(sc.env_maps[lt.tri].light_from_render, ray_d);
        return env_image_le(sc, lt, equal_area_sphere_to_square(wl), lam);
    }
    return lt.scale * spectrum_sample(sc, lt.spectrum, lam);
}
// ImageInfinitelight::pdf_li light.rs:882-892 / UniformInfiniteLight::pdf_li :768-780
static __device__ __noinline__ float infinite_pdf_li(const DScene& sc, const SgLight& lt, float3 wi, bool allow_incomplete) {
    if (lt.kind == SG_LIGHT_IMAGE_INFINITE) {
        const SgEnvMap& E = sc.env_maps[lt.tri];
        const float2 uv = equal_area_sphere_to_square(xform_vector(E.light_from_render, wi));
        return pc2d_pdf(sc, allow_incomplete ? E.compensated : E.distribution, uv) / (4.0f * kPi);
    }
    return allow_incomplete ? 0.0f : kInv4Pi;
}
// ImageInfinitelight::sample_li light.rs:847-880 / UniformInfiniteLight::sample_li :740-766
static __device__ __noinline__ bool infinite_sample_li(const DScene& sc, const SgLight& lt, const LightCtx& ctx, float2 u, const Wavelengths& lam, bool allow_incomplete,
                                                LightSample& ls) {
    const float3 cp = p3fi_mid(ctx.pi);
    if (lt.kind == SG_LIGHT_IMAGE_INFINITE) {
        const SgEnvMap& E = sc.env_maps[lt.tri];
        float map_pdf;
        const float2 uv = pc2d_sample(sc, allow_incomplete ? E.compensated : E.distribution, u, map_pdf);
        if (map_pdf == 0.0f) return false;
        const float3 wi = xform_vector(E.render_from_light, equal_area_square_to_sphere(uv));
        ls.l = env_image_le(sc, lt, uv, lam); ls.wi = wi; ls.pdf = map_pdf / (4.0f * kPi);
        ls.p_light = p3fi_exact(cp + wi * (2.0f * lt.scene_radius)); ls.n_light = f3(0.0f, 0.0f, 0.0f);
        return true;
    }
    if (allow_incomplete) return false;
    const float z = 1.0f - 2.0f * u.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * kPi * u.y;         // sample_uniform_sphere sampling.rs:280-289
    const float3 wi = f3(r * cosf(phi), r * sinf(phi), z);
    ls.l = lt.scale * spectrum_sample(sc, lt.spectrum, lam); ls.wi = wi; ls.pdf = kInv4Pi;           // uniform_hemisphere_pdf() == 1/(4 pi) (sampling.rs:306-308)
    ls.p_light = p3fi_exact(cp + wi * (2.0f * lt.scene_radius)); ls.n_light = f3(0.0f, 0.0f, 0.0f);
    return true;
}

}  // namespace sg
