// Device shading layer: spectra, sampling warps, BxDFs, area/point/infinite lights,
// perspective camera and film accumulation.  Each function cites the reference routine
// (paths relative to /root/reference/src) whose arithmetic it reproduces.
#pragma once
#include "sg_sphere.cuh"
#include "sg_host_tables.h"     // kSpecLutMin / kSpecLutMax (LAMBDA_MIN / LAMBDA_MAX, spectrum.rs)
#include "sg_scene.cuh"

namespace sg {

// ---------------- 4-wide spectra (spectra/mod.rs:17) as float4 ----------------
typedef float4 Spec;
SGD Spec spec1(float c) { return make_float4(c, c, c, c); }
SGD Spec operator+(Spec a, Spec b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SGD Spec operator*(Spec a, Spec b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
SGD Spec operator*(Spec a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
SGD Spec operator*(float s, Spec a) { return a * s; }
SGD Spec operator/(Spec a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
SGD bool spec_zero(Spec a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f && a.w == 0.0f; }
SGD float spec_max(Spec a) { return fmaxf(fmaxf(fmaxf(a.x, a.y), a.z), a.w); }          // fold(NaN, max) sampled_spectrum.rs:113-118
SGD Spec spec_clamp(Spec a, float lo, float hi) { return make_float4(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi), clampf(a.w, lo, hi)); }
SGD float& sref(Spec& s, int i) { return i == 0 ? s.x : (i == 1 ? s.y : (i == 2 ? s.z : s.w)); }
SGD float sget(const Spec& s, int i) { return i == 0 ? s.x : (i == 1 ? s.y : (i == 2 ? s.z : s.w)); }

struct Wavelengths { Spec lambda, pdf; };

// sampling.rs:268-278, sampled_wavelengths.rs:57-71
SGD float sample_visible_wavelengths(float u) { return 538.0f - 138.888889f * atanhf(0.85691062f - 1.82750197f * u); }
SGD float visible_wavelengths_pdf(float l) {
    if (l < 360.0f || l > 830.0f) return 0.0f;
    float x = coshf(0.0072f * (l - 538.0f));
    return 0.0039398042f / (x * x);
}
SGD Wavelengths sample_visible(float u) {
    Wavelengths w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float up = u + (float)i / 4.0f;
        if (up > 1.0f) up -= 1.0f;
        float l = sample_visible_wavelengths(up);
        sref(w.lambda, i) = l;
        sref(w.pdf, i) = visible_wavelengths_pdf(l);
    }
    return w;
}
SGD void terminate_secondary(Wavelengths& w) {                   // sampled_wavelengths.rs:79-96
    if (w.pdf.y == 0.0f && w.pdf.z == 0.0f && w.pdf.w == 0.0f) return;
    w.pdf.y = 0.0f; w.pdf.z = 0.0f; w.pdf.w = 0.0f;
    w.pdf.x /= 4.0f;
}

SGD float blackbody(float lambda, float temperature) {           // spectrum.rs:462-476
    if (temperature < 0.0f) return 0.0f;
    const float c = 299792458.0f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
    float l = lambda * 1e-9f;
    float l2 = l * l; float l5 = l2 * l2 * l;
    return (2.0f * h * c * c) / (l5 * (expf((h * c) / (l * kb * temperature)) - 1.0f));
}
SGD int find_interval_le(const float* L, int size, float lambda) {   // math.rs:322-333 with pred = L[i] <= lambda
    int first = 1, last = size - 2;
    while (last > 0) {
        int half = last >> 1, middle = first + half;
        bool r = __ldg(L + middle) <= lambda;
        first = r ? middle + 1 : first;
        last = r ? last - (half + 1) : half;
    }
    int v = first - 1;
    return v < 0 ? 0 : (v > size - 2 ? size - 2 : v);
}
// Spectrum::get, spectrum.rs:52-62
SGD float spectrum_get(const DScene& sc, int id, float lambda) {
    const SgSpectrum s = sc.spectra[id];
    switch (s.kind) {
    case SG_SPECTRUM_CONSTANT: return s.c;
    case SG_SPECTRUM_DENSE: {
        int off = f2i_sat(lambda) - s.lambda_min;
        return (off < 0 || off >= s.n) ? 0.0f : __ldg(sc.pool + s.off_a + off);
    }
    case SG_SPECTRUM_PIECEWISE_LINEAR: {
        const float* L = sc.pool + s.off_a; const float* V = sc.pool + s.off_b;
        if (s.n == 0 || lambda < __ldg(L) || lambda > __ldg(L + s.n - 1)) return 0.0f;
        // find_interval (math.rs:322-333) = the largest o in [0, n-2] with o == 0 or L[o] <= lambda.  Instead of the binary search
        // (six dependent loads for a 48-knot spectrum, four wavelengths per lookup) start from that index for floor(lambda), tabulated
        // at upload, and step forward over the knots inside the same 1 nm bin: the same index by construction, so the same value.
        int o;
        const int fl = __float2int_rz(lambda);
        if (s.pad != 0u && fl >= kSpecLutMin && fl <= kSpecLutMax) {
            o = (int)__ldg(sc.spec_lut + (s.pad - 1u) + (uint32_t)(fl - kSpecLutMin));
            while (o < s.n - 2 && __ldg(L + o + 1) <= lambda) ++o;
        } else o = find_interval_le(L, s.n, lambda);
        float l0 = __ldg(L + o), l1 = __ldg(L + o + 1);
        float t = (lambda - l0) / (l1 - l0);
        return lerpf(t, __ldg(V + o), __ldg(V + o + 1));
    }
    case SG_SPECTRUM_BLACKBODY: return blackbody(lambda, s.c) * s.scale;
    }
    return 0.0f;
}
// Spectrum::sample, spectrum.rs:76-86; the dense variant ROUNDS (:283), `get` truncates (:265)
SGD Spec spectrum_sample(const DScene& sc, int id, const Wavelengths& w) {
    const SgSpectrum s = sc.spectra[id];
    Spec r;
    if (s.kind == SG_SPECTRUM_DENSE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int off = f2i_sat(roundf(sget(w.lambda, i))) - s.lambda_min;
            sref(r, i) = (off < 0 || off >= s.n) ? 0.0f : __ldg(sc.pool + s.off_a + off);
        }
        return r;
    }
    if (s.kind == SG_SPECTRUM_CONSTANT) return spec1(s.c);
#pragma unroll
    for (int i = 0; i < 4; ++i) sref(r, i) = spectrum_get(sc, id, sget(w.lambda, i));
    return r;
}

// ---------------- sampling.rs ----------------
SGD float power_heuristic(float f_pdf, float g_pdf) {            // :187-194, nf = ng = 1
    float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
    if (isinf(sqr(f))) return 1.0f;
    return (f * f) / (f * f + g * g);
}
SGD float sample_linear(float u, float a, float b) {             // :250-257
    if (u == 0.0f && a == 0.0f) return 0.0f;
    float x = u * (a + b) / (a + sqrtf(lerpf(u, a * a, b * b)));
    return fminf(x, 1.0f - 1.1920929e-07f);
}
SGD float2 sample_bilinear(float2 u, const float w[4]) {         // :386-393
    float2 p;
    p.y = sample_linear(u.y, w[0] + w[1], w[2] + w[3]);
    p.x = sample_linear(u.x, lerpf(p.y, w[0], w[2]), lerpf(p.y, w[1], w[3]));
    return p;
}
SGD float bilinear_pdf(float2 p, const float w[4]) {             // :395-408
    if (p.x < 0.0f || p.x > 1.0f || p.y < 0.0f || p.y > 1.0f) return 0.0f;
    if (w[0] + w[1] + w[2] + w[3] == 0.0f) return 1.0f;
    return 4.0f * ((1.0f - p.x) * (1.0f - p.y) * w[0] + p.x * (1.0f - p.y) * w[1] + (1.0f - p.x) * p.y * w[2] + p.x * p.y * w[3])
           / (w[0] + w[1] + w[2] + w[3]);
}
SGD float2 sample_disk_concentric(float2 u) {                    // :324-339
    float ox = 2.0f * u.x - 1.0f, oy = 2.0f * u.y - 1.0f;
    if (ox == 0.0f && oy == 0.0f) return make_float2(0.0f, 0.0f);
    float r, theta;
    if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kPiOver4 * (oy / ox); }
    else { r = oy; theta = kPiOver2 - kPiOver4 * (ox / oy); }
    return make_float2(r * cosf(theta), r * sinf(theta));
}
SGD float2 sample_disk_polar(float2 u) {                         // :341-345
    float r = sqrtf(u.x), theta = 2.0f * kPi * u.y;
    return make_float2(r * cosf(theta), r * sinf(theta));
}
SGD float3 sample_cosine_hemisphere(float2 u) {                  // :310-318
    float2 d = sample_disk_concentric(u);
    float z = safe_sqrt(1.0f - sqr(d.x) - sqr(d.y));
    return f3(d.x, d.y, z);
}
// :412-499 -- reproduces `divisor = e1.dot(e1)` and the `(b1 / b1 + b2, ...)` renormalisation
SGD void sample_spherical_triangle(float3 v0, float3 v1, float3 v2, float3 p, float2 u, float& o0, float& o1, float& o2, float& pdf_out) {
    float3 a = normalize3(v0 - p), b = normalize3(v1 - p), c = normalize3(v2 - p);
    float3 n_ab = cross3(a, b), n_bc = cross3(b, c), n_ca = cross3(c, a);
    if (len2(n_ab) == 0.0f || len2(n_bc) == 0.0f || len2(n_ca) == 0.0f) { o0 = o1 = o2 = 0.0f; pdf_out = 0.0f; return; }
    n_ab = normalize3(n_ab); n_bc = normalize3(n_bc); n_ca = normalize3(n_ca);
    float alpha = angle_between3(n_ab, -n_ca), beta = angle_between3(n_bc, -n_ab), gam = angle_between3(n_ca, -n_bc);
    float a_pi = alpha + beta + gam;
    float ap_pi = lerpf(u.x, kPi, a_pi);
    float area = a_pi - kPi;
    float pdf = area <= 0.0f ? 0.0f : 1.0f / area;
    float cos_alpha = cosf(alpha), sin_alpha = sinf(alpha);
    float s_ap = sinf(ap_pi), c_ap = cosf(ap_pi);
    float sin_phi = s_ap * cos_alpha - c_ap * sin_alpha;
    float cos_phi = c_ap * cos_alpha + s_ap * sin_alpha;
    float k1 = cos_phi + cos_alpha;
    float k2 = sin_phi - sin_alpha * dot3(a, b);
    float cos_bp = (k2 + (dop(k2, cos_phi, k1, sin_phi)) * cos_alpha) / (sop(k2, sin_phi, k1, cos_phi) * sin_alpha);
    cos_bp = clampf(cos_bp, -1.0f, 1.0f);
    float sin_bp = safe_sqrt(1.0f - cos_bp * cos_bp);
    float3 cp = cos_bp * a + sin_bp * normalize3(gram_schmidt3(c, a));
    float cos_theta = 1.0f - u.y * (1.0f - dot3(cp, b));
    float sin_theta = safe_sqrt(1.0f - cos_theta * cos_theta);
    float3 w = cos_theta * b + sin_theta * normalize3(gram_schmidt3(cp, b));
    float3 e1 = v1 - v0, e2 = v2 - v0;
    float3 s1 = cross3(w, e2);
    float divisor = dot3(e1, e1);
    if (divisor == 0.0f) { o0 = o1 = o2 = 1.0f / 3.0f; pdf_out = pdf; return; }
    float inv_divisor = 1.0f / divisor;
    float3 s = p - v0;
    float b1 = dot3(s, s1) * inv_divisor;
    float b2 = dot3(w, cross3(s, e1)) * inv_divisor;
    b1 = clampf(b1, 0.0f, 1.0f); b2 = clampf(b2, 0.0f, 1.0f);
    if (b1 + b2 > 1.0f) { float nb1 = b1 / b1 + b2, nb2 = b2 / b1 + b2; b1 = nb1; b2 = nb2; }
    o0 = 1.0f - b1 - b2; o1 = b1; o2 = b2; pdf_out = pdf;
}
// :581-641
SGD float2 invert_spherical_triangle_sample(float3 v0, float3 v1, float3 v2, float3 p, float3 w) {
    float3 a = normalize3(v0 - p), b = normalize3(v1 - p), c = normalize3(v2 - p);
    float3 n_ab = cross3(a, b), n_bc = cross3(b, c), n_ca = cross3(c, a);
    if (len2(n_ab) == 0.0f || len2(n_bc) == 0.0f || len2(n_ca) == 0.0f) return make_float2(0.0f, 0.0f);
    n_ab = normalize3(n_ab); n_bc = normalize3(n_bc); n_ca = normalize3(n_ca);
    float alpha = angle_between3(n_ab, -n_ca), beta = angle_between3(n_bc, -n_ab), gam = angle_between3(n_ca, -n_bc);
    float3 cp = normalize3(cross3(cross3(b, w), cross3(c, a)));
    if (dot3(cp, a + c) < 0.0f) cp = -cp;
    float u0;
    if (dot3(a, cp) > 0.99999847691f) u0 = 0.0f;
    else {
        float3 n_cpb = cross3(cp, b), n_acp = cross3(a, cp);
        if (len2(n_cpb) == 0.0f || len2(n_acp) == 0.0f) return make_float2(0.5f, 0.5f);
        n_cpb = normalize3(n_cpb); n_acp = normalize3(n_acp);
        float ap = alpha + angle_between3(n_ab, n_cpb) + angle_between3(n_acp, -n_cpb) - kPi;
        float area = alpha + beta + gam - kPi;
        u0 = ap / area;
    }
    float u1 = (1.0f - dot3(w, b)) / (1.0f - dot3(cp, b));
    return make_float2(clampf(u0, 0.0f, 1.0f), clampf(u1, 0.0f, 1.0f));
}

// ---------------- scattering.rs / vecmath/spherical.rs ----------------
SGD float cos2_theta(float3 w) { return w.z * w.z; }
SGD float sin2_theta(float3 w) { return fmaxf(0.0f, 1.0f - cos2_theta(w)); }
SGD float sin_theta(float3 w) { return sqrtf(sin2_theta(w)); }
SGD float tan2_theta(float3 w) { return sin2_theta(w) / cos2_theta(w); }
SGD float cos_phi(float3 w) { float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }   // spherical.rs:60-67
SGD float sin_phi(float3 w) { float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.y / s, -1.0f, 1.0f); }   // spherical.rs:69-76 (1.0: reference quirk)
SGD bool same_hemisphere(float3 w, float3 wp) { return w.z * wp.z > 0.0f; }
SGD float3 reflect3(float3 wo, float3 n) { return -wo + 2.0f * dot3(wo, n) * n; }                                  // scattering.rs:12-14
SGD bool refract3(float3 wi, float3 n, float eta, float3& wt, float& etap) {                                      // :21-43
    float ci = dot3(n, wi);
    if (ci < 0.0f) { eta = 1.0f / eta; ci = -ci; n = -n; }
    float s2i = fmaxf(0.0f, 1.0f - sqr(ci));
    float s2t = s2i / sqr(eta);
    if (s2t >= 1.0f) return false;
    float ct = sqrtf(1.0f - s2t);
    wt = -wi / eta + (ci / eta - ct) * n;
    etap = eta;
    return true;
}
SGD float fresnel_dielectric(float ci, float eta) {                                                               // :49-70
    ci = clampf(ci, -1.0f, 1.0f);
    if (ci < 0.0f) { eta = 1.0f / eta; ci = -ci; }
    float s2i = 1.0f - ci * ci;
    float s2t = s2i / (eta * eta);
    if (s2t >= 1.0f) return 1.0f;
    float ct = safe_sqrt(1.0f - s2t);
    float r_parl = (eta * ci - ct) / (eta * ci + ct);
    float r_perp = (ci - eta * ct) / (ci + eta * ct);
    return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// num-complex 0.4.4 Complex<f32> arithmetic (third-party, restated from its published source)
struct Cx { float re, im; };
SGD Cx cxm(float r, float i) { Cx c; c.re = r; c.im = i; return c; }
SGD Cx cx_mul(Cx a, Cx b) { return cxm(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
SGD Cx cx_div(Cx a, Cx b) { float ns = b.re * b.re + b.im * b.im; return cxm((a.re * b.re + a.im * b.im) / ns, (a.im * b.re - a.re * b.im) / ns); }
SGD Cx cx_sqrt(Cx z) {
    if (z.im == 0.0f) {
        if (!signbit(z.re)) return cxm(sqrtf(z.re), z.im);
        float im = sqrtf(-z.re);
        return cxm(0.0f, signbit(z.im) ? -im : im);
    } else if (z.re == 0.0f) {
        float x = sqrtf(fabsf(z.im) / 2.0f);
        return cxm(x, signbit(z.im) ? -x : x);
    }
    float r = hypotf(z.re, z.im), theta = atan2f(z.im, z.re);
    float sr = sqrtf(r), ht = theta / 2.0f;
    return cxm(sr * cosf(ht), sr * sinf(ht));
}
SGD float fresnel_complex(float ci, Cx eta) {                                                                     // :78-89
    ci = clampf(ci, 0.0f, 1.0f);
    float s2i = 1.0f - sqr(ci);
    Cx s2t = cx_div(cxm(s2i, 0.0f), cx_mul(eta, eta));
    Cx ct = cx_sqrt(cxm(1.0f - s2t.re, 0.0f - s2t.im));
    Cx eci = cxm(eta.re * ci, eta.im * ci);
    Cx r_parl = cx_div(cxm(eci.re - ct.re, eci.im - ct.im), cxm(eci.re + ct.re, eci.im + ct.im));
    Cx ect = cx_mul(eta, ct);
    Cx r_perp = cx_div(cxm(ci - ect.re, 0.0f - ect.im), cxm(ci + ect.re, 0.0f + ect.im));
    return ((r_parl.re * r_parl.re + r_parl.im * r_parl.im) + (r_perp.re * r_perp.re + r_perp.im * r_perp.im)) / 2.0f;
}
SGD Spec fresnel_complex_spectral(float ci, Spec eta, Spec k) {                                                   // :94-105
    return make_float4(fresnel_complex(ci, cxm(eta.x, k.x)), fresnel_complex(ci, cxm(eta.y, k.y)),
                       fresnel_complex(ci, cxm(eta.z, k.z)), fresnel_complex(ci, cxm(eta.w, k.w)));
}
// TrowbridgeReitzDistribution, scattering.rs:107-220
struct TR {
    float ax, ay;
    SGD static TR make(float ax, float ay) {
        TR d; d.ax = ax; d.ay = ay;
        if (!d.smooth()) { d.ax = fmaxf(d.ax, 1e-4f); d.ay = fmaxf(d.ay, 1e-4f); }
        return d;
    }
    SGD bool smooth() const { return ax < 1e-3f && ay < 1e-3f; }
    SGD float d(float3 wm) const {
        float t2 = tan2_theta(wm);
        if (isinf(t2)) return 0.0f;
        float c4 = sqr(cos2_theta(wm));
        if (c4 < 1e-16f) return 0.0f;
        float e = t2 * (sqr(cos_phi(wm) / ax) + sqr(sin_phi(wm) / ay));
        return 1.0f / (kPi * ax * ay * c4 * sqr(1.0f + e));
    }
    SGD float lambda(float3 w) const {
        float t2 = tan2_theta(w);
        if (isinf(t2)) return 0.0f;
        float a2 = sqr(cos_phi(w) * ax) + sqr(sin_phi(w) * ay);
        return (-1.0f + sqrtf(1.0f + a2 * t2)) / 2.0f;
    }
    SGD float g1(float3 w) const { return 1.0f / (1.0f + lambda(w)); }
    SGD float g(float3 wo, float3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    SGD float pdf(float3 w, float3 wm) const { return g1(w) / fabsf(w.z) * d(wm) * absdot3(w, wm); }
    SGD float3 sample_wm(float3 w, float2 u) const {
        float3 wh = normalize3(f3(ax * w.x, ay * w.y, w.z));
        if (wh.z < 0.0f) wh = -wh;
        float3 t1 = wh.z < 0.99999f ? normalize3(cross3(f3(0.0f, 0.0f, 1.0f), wh)) : f3(1.0f, 0.0f, 0.0f);
        float3 t2 = cross3(wh, t1);
        float2 p = sample_disk_polar(u);
        float h = sqrtf(1.0f - sqr(p.x));
        p.y = lerpf((1.0f + wh.z) / 2.0f, h, p.y);
        float pz = sqrtf(fmaxf(0.0f, 1.0f - (p.x * p.x + p.y * p.y)));
        float3 nh = p.x * t1 + p.y * t2 + pz * wh;
        return normalize3(f3(ax * nh.x, ay * nh.y, fmaxf(1e-6f, nh.z)));
    }
    SGD void regularize() {
        if (ax < 0.3f) ax = clampf(2.0f * ax, 0.1f, 0.3f);
        if (ay < 0.3f) ay = clampf(2.0f * ay, 0.1f, 0.3f);
    }
};

// ---------------- BxDFs (bxdf.rs) behind the BSDF frame wrapper (bsdf.rs) ----------------
enum { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16 };   // bxdf.rs:1773-1789
struct BSDFSample { Spec f; float3 wi; float pdf; int flags; float eta; };

}  // namespace sg
#include "sg_layered.cuh"     // CoatedDiffuse (needs TR, BSDFSample and the dielectric helpers above)
namespace sg {

template <int KIND> struct BSDF {
    Spec r, k;            // diffuse: r ; conductor: eta (in r), k
    float eta;            // dielectric
    TR mf;
    static constexpr bool LAYERED = KIND == SG_MATERIAL_COATED_DIFFUSE || KIND == SG_MATERIAL_COATED_CONDUCTOR;
    LayeredT<KIND == SG_MATERIAL_COATED_CONDUCTOR> lay;   // CoatedDiffuse / CoatedConductor only
    uint64_t layer_seed;  // seeds the LayeredBxDF's private generator for the next f / sample_f / pdf call
    bool proportional;    // BSDFSample::pdf_is_proportional of the last sample_f
    float3 fx, fy, fz;    // Frame::from_xz(normalize(dpdus), ns), bsdf.rs:22-28
    SGD Rng layer_rng() const { Rng r; r.seed_from_u64(layer_seed); return r; }

    SGD float3 to_local(float3 v) const { return f3(dot3(v, fx), dot3(v, fy), dot3(v, fz)); }      // frame.rs:39-41
    SGD float3 from_local(float3 v) const { return v.x * fx + v.y * fy + v.z * fz; }               // frame.rs:51-53
    SGD int flags() const {
        if (LAYERED) return lay.flags();
        if (KIND == SG_MATERIAL_THIN_DIELECTRIC) return BX_REFLECTION | BX_TRANSMISSION | BX_SPECULAR;   // bxdf.rs:873-875
        if (KIND == SG_MATERIAL_DIFFUSE) return spec_zero(r) ? 0 : (BX_DIFFUSE | BX_REFLECTION);     // bxdf.rs:256-262
        if (KIND == SG_MATERIAL_CONDUCTOR) return (mf.smooth() ? BX_SPECULAR : BX_GLOSSY) | BX_REFLECTION;   // :447-453
        int f = (eta == 1.0f) ? BX_TRANSMISSION : (BX_REFLECTION | BX_TRANSMISSION);               // :778-790
        return f | (mf.smooth() ? BX_SPECULAR : BX_GLOSSY);
    }
    SGD Spec f_local(float3 wo, float3 wi) const {
        if (LAYERED) return lay.f(wo, wi, layer_rng());
        if (KIND == SG_MATERIAL_THIN_DIELECTRIC) return spec1(0.0f);                                 // bxdf.rs:808-810
        if (KIND == SG_MATERIAL_DIFFUSE) {                                                          // :196-202
            if (!same_hemisphere(wo, wi)) return spec1(0.0f);
            return r * kInvPi;
        } else if (KIND == SG_MATERIAL_CONDUCTOR) {                                                 // :349-376
            if (!same_hemisphere(wo, wi)) return spec1(0.0f);
            if (mf.smooth()) return spec1(0.0f);
            float cto = fabsf(wo.z), cti = fabsf(wi.z);
            if (cti == 0.0f || cto == 0.0f) return spec1(0.0f);
            float3 wm = wi + wo;
            if (len2(wm) == 0.0f) return spec1(0.0f);
            wm = normalize3(wm);
            Spec F = fresnel_complex_spectral(absdot3(wo, wm), r, k);
            return mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
        } else {                                                                                    // :533-584
            if (eta == 1.0f || mf.smooth()) return spec1(0.0f);
            float cto = wo.z, cti = wi.z;
            bool refl = cti * cto > 0.0f;
            float etap = 1.0f;
            if (!refl) etap = cto > 0.0f ? eta : (1.0f / eta);
            float3 wm = wi * etap + wo;
            if (cti == 0.0f || cto == 0.0f || len2(wm) == 0.0f) return spec1(0.0f);
            wm = faceforward3(normalize3(wm), f3(0.0f, 0.0f, 1.0f));
            if (dot3(wm, wi) * cti < 0.0f || dot3(wm, wo) * cto < 0.0f) return spec1(0.0f);
            float F = fresnel_dielectric(dot3(wo, wm), eta);
            if (refl) return spec1(mf.d(wm) * mf.g(wo, wi) * F / fabsf(4.0f * cti * cto));
            float denom = sqr(dot3(wi, wm) + dot3(wo, wm) / etap) * cti * cto;
            float ft = mf.d(wm) * (1.0f - F) * mf.g(wo, wi) * fabsf(dot3(wi, wm) * dot3(wo, wm) / denom);
            ft /= sqr(etap);                                        // TransportMode::Radiance
            return spec1(ft);
        }
    }
    SGD float pdf_local(float3 wo, float3 wi) const {
        if (LAYERED) return lay.pdf(wo, wi, layer_rng());
        if (KIND == SG_MATERIAL_THIN_DIELECTRIC) return 0.0f;                                        // bxdf.rs:863-871
        if (KIND == SG_MATERIAL_DIFFUSE) {                                                          // :240-254
            if (!same_hemisphere(wo, wi)) return 0.0f;
            return fabsf(wi.z) * kInvPi;
        } else if (KIND == SG_MATERIAL_CONDUCTOR) {                                                 // :424-445
            if (!same_hemisphere(wo, wi) || mf.smooth()) return 0.0f;
            float3 wm = wo + wi;
            if (len2(wm) == 0.0f) return 0.0f;
            wm = faceforward3(normalize3(wm), f3(0.0f, 0.0f, 1.0f));
            return mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm));
        } else {                                                                                    // :715-776
            if (eta == 1.0f || mf.smooth()) return 0.0f;
            float cto = wo.z, cti = wi.z;
            bool refl = cti * cto > 0.0f;
            float etap = 1.0f;
            if (!refl) etap = cto > 0.0f ? eta : (1.0f / eta);
            float3 wm = wi * etap + wo;
            if (cti == 0.0f || cto == 0.0f || len2(wm) == 0.0f) return 0.0f;
            wm = faceforward3(normalize3(wm), f3(0.0f, 0.0f, 1.0f));
            if (dot3(wm, wi) * cti < 0.0f || dot3(wm, wo) * cto < 0.0f) return 0.0f;
            float R = fresnel_dielectric(dot3(wo, wm), eta), T = 1.0f - R;
            float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return 0.0f;
            if (refl) return mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm)) * pr / (pr + pt);
            float denom = sqr(dot3(wi, wm) + dot3(wo, wm) / etap);
            float dwm_dwi = absdot3(wi, wm) / denom;
            return mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
        }
    }
    SGD bool sample_local(float3 wo, float uc, float2 u, BSDFSample& bs, bool& prop) const {
        bs.eta = 1.0f; prop = false;
        if (LAYERED) return lay.sample_f(wo, uc, u, layer_rng(), bs, prop);
        if (KIND == SG_MATERIAL_THIN_DIELECTRIC) {                                                   // ThinDielectricBxDF::sample_f bxdf.rs:812-861
            float R = fresnel_dielectric(fabsf(wo.z), eta), T = 1.0f - R;
            if (R < 1.0f) { R += sqr(T) * R / (1.0f - sqr(R)); T = 1.0f - R; }
            const float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return false;
            if (uc < pr / (pr + pt)) {
                const float3 wi = f3(-wo.x, -wo.y, wo.z);
                bs.f = spec1(R / fabsf(wi.z)); bs.wi = wi; bs.pdf = pr / (pr + pt); bs.flags = BX_SPECULAR | BX_REFLECTION;
            } else {
                const float3 wi = -wo;
                bs.f = spec1(T / fabsf(wi.z)); bs.wi = wi; bs.pdf = pt / (pr + pt); bs.flags = BX_SPECULAR | BX_TRANSMISSION;
            }
            return true;
        }
        if (KIND == SG_MATERIAL_DIFFUSE) {                                                          // :204-238
            float3 wi = sample_cosine_hemisphere(u);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            bs.f = r * kInvPi; bs.wi = wi; bs.pdf = fabsf(wi.z) * kInvPi; bs.flags = BX_DIFFUSE | BX_REFLECTION;
            return true;
        } else if (KIND == SG_MATERIAL_CONDUCTOR) {                                                 // :378-422
            if (mf.smooth()) {
                float3 wi = f3(-wo.x, -wo.y, wo.z);
                bs.f = fresnel_complex_spectral(fabsf(wi.z), r, k) / fabsf(wi.z);
                bs.wi = wi; bs.pdf = 1.0f; bs.flags = BX_SPECULAR | BX_REFLECTION;
                return true;
            }
            if (wo.z == 0.0f) return false;
            float3 wm = mf.sample_wm(wo, u);
            float3 wi = reflect3(wo, wm);
            if (!same_hemisphere(wo, wi)) return false;
            float pdf = mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm));
            float cto = fabsf(wo.z), cti = fabsf(wi.z);
            if (cti == 0.0f || cto == 0.0f) return false;
            Spec F = fresnel_complex_spectral(absdot3(wo, wm), r, k);
            bs.f = mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
            bs.wi = wi; bs.pdf = pdf; bs.flags = BX_GLOSSY | BX_REFLECTION;
            return true;
        } else {                                                                                    // :586-713
            if (eta == 1.0f || mf.smooth()) {
                float R = fresnel_dielectric(wo.z, eta), T = 1.0f - R;
                float pr = R, pt = T;
                if (pr == 0.0f && pt == 0.0f) return false;
                if (uc < pr / (pr + pt)) {
                    float3 wi = f3(-wo.x, -wo.y, wo.z);
                    bs.f = spec1(R / fabsf(wi.z)); bs.wi = wi; bs.pdf = pr / (pr + pt); bs.flags = BX_SPECULAR | BX_REFLECTION;
                    return true;
                }
                float3 wi; float etap;
                if (!refract3(wo, f3(0.0f, 0.0f, 1.0f), eta, wi, etap)) return false;
                float ft = T / fabsf(wi.z);
                ft /= sqr(etap);
                bs.f = spec1(ft); bs.wi = wi; bs.pdf = pt / (pr + pt); bs.flags = BX_SPECULAR | BX_TRANSMISSION; bs.eta = etap;
                return true;
            }
            float3 wm = mf.sample_wm(wo, u);
            float R = fresnel_dielectric(dot3(wo, wm), eta), T = 1.0f - R;
            float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return false;
            if (uc < pr / (pr + pt)) {
                float3 wi = reflect3(wo, wm);
                if (!same_hemisphere(wo, wi)) return false;
                float pdf = mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm)) * pr / (pr + pt);
                bs.f = spec1(mf.d(wm) * mf.g(wo, wi) * R / (4.0f * wi.z * wo.z));
                bs.wi = wi; bs.pdf = pdf; bs.flags = BX_GLOSSY | BX_REFLECTION;
                return true;
            }
            float3 wi; float etap;
            if (!refract3(wo, wm, eta, wi, etap)) return false;
            if (same_hemisphere(wo, wi) || wi.z == 0.0f) return false;
            float denom = sqr(dot3(wi, wm) + dot3(wo, wm) / etap);
            float dwm_dwi = absdot3(wi, wm) / denom;
            float pdf = mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
            float ft = T * mf.d(wm) * mf.g(wo, wi) * fabsf(dot3(wi, wm) * dot3(wo, wm) / (wi.z * wo.z * denom));
            ft /= sqr(etap);
            bs.f = spec1(ft); bs.wi = wi; bs.pdf = pdf; bs.flags = BX_GLOSSY | BX_TRANSMISSION; bs.eta = etap;
            return true;
        }
    }
    SGD Spec f(float3 wo_r, float3 wi_r) const {                     // bsdf.rs:44-58
        float3 wi = to_local(wi_r), wo = to_local(wo_r);
        if (wo.z == 0.0f) return spec1(0.0f);
        return f_local(wo, wi);
    }
    SGD bool sample_f(float3 wo_r, float uc, float2 u, BSDFSample& bs, bool& prop) const {     // bsdf.rs:60-82
        float3 wo = to_local(wo_r);
        if (wo.z == 0.0f || !(flags() & (BX_REFLECTION | BX_TRANSMISSION))) return false;
        if (!sample_local(wo, uc, u, bs, prop)) return false;
        if (spec_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) return false;
        bs.wi = from_local(bs.wi);
        return true;
    }
    SGD float pdf(float3 wo_r, float3 wi_r) const {                  // bsdf.rs:84-97
        float3 wo = to_local(wo_r), wi = to_local(wi_r);
        if (wo.z == 0.0f) return 0.0f;
        return pdf_local(wo, wi);
    }
};

// ---------------- surface interaction (triangle.rs:305-504, interaction.rs:111-148,379-405) -------------
struct Surf {
    P3fi pi; float3 n; float3 sn; float3 sdpdu, sdpdv;
};
SGD float3 ldv3(const float* a, size_t i) { return f3(__ldg(a + 3 * i), __ldg(a + 3 * i + 1), __ldg(a + 3 * i + 2)); }

// Triangle geometry handle.  Vertices come from the pre-gathered tri_verts / light_verts records
// (one cache line, already touched by the traversal kernel); the index/attribute arrays are only
// read for meshes that carry normals, uvs or tangents.
struct TriGeo { float3 p0, p1, p2; uint32_t flags; uint32_t mesh; uint32_t tri; uint32_t prim; uint32_t kind; };   // kind: SgMaterialKind of the primitive's material (hits only)
static constexpr uint32_t kTriUnknown = 0xffffffffu;
SGD TriGeo geo_from_prim(const DScene& sc, uint32_t prim_id, uint32_t& material, int& light) {
    const float4 v0 = __ldg(sc.tri_verts + 3 * (size_t)prim_id), v1 = __ldg(sc.tri_verts + 3 * (size_t)prim_id + 1),
                 v2 = __ldg(sc.tri_verts + 3 * (size_t)prim_id + 2);
    const uint32_t w0 = __float_as_uint(v0.w);
    TriGeo g; g.p0 = f3(v0.x, v0.y, v0.z); g.p1 = f3(v1.x, v1.y, v1.z); g.p2 = f3(v2.x, v2.y, v2.z);
    g.flags = (w0 >> 23) & 31u; g.mesh = __float_as_uint(v2.w) & 0x7fffffffu; g.tri = kTriUnknown; g.prim = prim_id; g.kind = (w0 >> 28) & 7u;
    material = w0 & 0x7fffffu; light = (int)__float_as_uint(v1.w);
    return g;
}
SGD TriGeo geo_from_light(const DScene& sc, uint32_t light_id, const SgLight& lt) {
    const float4 v0 = __ldg(sc.light_verts + 3 * (size_t)light_id), v1 = __ldg(sc.light_verts + 3 * (size_t)light_id + 1),
                 v2 = __ldg(sc.light_verts + 3 * (size_t)light_id + 2);
    TriGeo g; g.p0 = f3(v0.x, v0.y, v0.z); g.p1 = f3(v1.x, v1.y, v1.z); g.p2 = f3(v2.x, v2.y, v2.z);
    g.flags = __float_as_uint(v0.w); g.mesh = lt.mesh; g.tri = lt.tri; g.prim = 0; g.kind = 0;
    return g;
}
SGD void geo_indices(const DScene& sc, const TriGeo& g, uint32_t& i0, uint32_t& i1, uint32_t& i2, size_t& fv) {
    const uint32_t tri = g.tri != kTriUnknown ? g.tri : sc.prims[g.prim].tri;
    const SgMesh m = sc.meshes[g.mesh];
    const uint32_t* ix = sc.indices + m.first_index + 3 * (size_t)tri;
    i0 = __ldg(ix); i1 = __ldg(ix + 1); i2 = __ldg(ix + 2); fv = m.first_vertex;
}

// Builds geometric + shading frame for the FINAL hit only (the reference does it for every
// accepted candidate along the ray, triangle.rs:529-535; only the last survives).
// The uv-derived dpdu/dpdv follow :314-372.  Scenes with image textures (or a non-zero displacement) also need the
// hit uv, the geometric dpdu/dpdv and dndu/dndv (:451-498): `x` (sg_texture.cuh SurfTex) receives them when TEX.
struct SurfTex;
template <bool TEX> SGD void surf_tex_store(SurfTex* x, float2 uv, float3 dpdu, float3 dpdv, float3 dndu, float3 dndv);
template <bool TEX = false>
SGD Surf make_surface(const DScene& sc, const TriGeo& g, float b0, float b1, float b2, SurfTex* x = nullptr) {
    struct { uint32_t flags; } m; m.flags = g.flags;
    const float3 p0 = g.p0, p1 = g.p1, p2 = g.p2;
    uint32_t i0 = 0, i1 = 0, i2 = 0; size_t fv = 0;
    if (m.flags & (SG_MESH_HAS_UV | SG_MESH_HAS_N | SG_MESH_HAS_S)) geo_indices(sc, g, i0, i1, i2, fv);
    float2 uv0 = make_float2(0.0f, 0.0f), uv1 = make_float2(1.0f, 0.0f), uv2 = make_float2(1.0f, 1.0f);
    if (m.flags & SG_MESH_HAS_UV) {
        uv0 = make_float2(__ldg(sc.uv + 2 * (fv + i0)), __ldg(sc.uv + 2 * (fv + i0) + 1));
        uv1 = make_float2(__ldg(sc.uv + 2 * (fv + i1)), __ldg(sc.uv + 2 * (fv + i1) + 1));
        uv2 = make_float2(__ldg(sc.uv + 2 * (fv + i2)), __ldg(sc.uv + 2 * (fv + i2) + 1));
    }
    float2 duv02 = make_float2(uv0.x - uv2.x, uv0.y - uv2.y), duv12 = make_float2(uv1.x - uv2.x, uv1.y - uv2.y);
    float3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = dop(duv02.x, duv12.y, duv02.y, duv12.x);
    bool degenerate_uv = fabsf(determinant) < 1e-9f;
    float3 dpdu = f3(0.0f, 0.0f, 0.0f), dpdv = f3(0.0f, 0.0f, 0.0f);
    if (!degenerate_uv) {
        float inv_det = 1.0f / determinant;
        // difference_of_products_float_vec, math.rs:214-219: (a*b - c*d) + ((-c)*d + c*d), unfused
        float3 cd = duv02.y * dp12; float3 df = duv12.y * dp02 - cd; float3 er = (-duv02.y) * dp12 + cd;
        dpdu = (df + er) * inv_det;
        cd = duv12.x * dp02; df = duv02.x * dp12 - cd; er = (-duv12.x) * dp02 + cd;
        dpdv = (df + er) * inv_det;
    }
    if (degenerate_uv || len2(cross3(dpdu, dpdv)) == 0.0f) {
        float3 ng = cross3(p2 - p0, p1 - p0);
        if (len2(ng) == 0.0f) {
            float3 v1 = p2 - p0, v2 = p1 - p0;
            ng = f3((float)dop_d(v1.y, v2.z, v1.z, v2.y), (float)dop_d(v1.z, v2.x, v1.x, v2.z), (float)dop_d(v1.x, v2.y, v1.y, v2.x));
        }
        coord_system(normalize3(ng), dpdu, dpdv);
    }
    float3 p_hit = b0 * p0 + b1 * p1 + b2 * p2;
    float3 p_abs_sum = abs3(b0 * p0) + abs3(b1 * p1) + abs3(b2 * p2);
    float3 p_error = gamma_n(7) * p_abs_sum;
    Surf s;
    s.pi = p3fi_make(p_hit, p_error);
    s.n = normalize3(cross3(dp02, dp12));                                   // :407-412
    const bool flip = ((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0);
    if (flip) s.n = -s.n;
    s.sn = s.n; s.sdpdu = dpdu; s.sdpdv = dpdv;
    float3 dndu = f3(0.0f, 0.0f, 0.0f), dndv = f3(0.0f, 0.0f, 0.0f);
    if (m.flags & (SG_MESH_HAS_N | SG_MESH_HAS_S)) {                        // :414-501
        float3 ns = s.n;
        if (m.flags & SG_MESH_HAS_N) {
            const float3 n0 = ldv3(sc.n, fv + i0), n1 = ldv3(sc.n, fv + i1), n2 = ldv3(sc.n, fv + i2);
            float3 nn = b0 * n0 + b1 * n1 + b2 * n2;
            if (len2(nn) > 0.0f) ns = normalize3(nn);
            if (TEX) {                                                      // dndu, dndv :451-498
                if (degenerate_uv) {
                    const float3 dn = cross3(n2 - n0, n1 - n0);
                    if (len2(dn) != 0.0f) coord_system(dn, dndu, dndv);
                } else {
                    const float inv_det = 1.0f / determinant;
                    const float3 dn1 = n0 - n2, dn2 = n1 - n2;
                    float3 cd = duv02.y * dn2; float3 df = duv12.y * dn1 - cd; float3 er = (-duv02.y) * dn2 + cd;
                    dndu = (df + er) * inv_det;
                    cd = duv12.x * dn1; df = duv02.x * dn2 - cd; er = (-duv12.x) * dn1 + cd;
                    dndv = (df + er) * inv_det;
                }
            }
        }
        float3 ss = dpdu;
        if (m.flags & SG_MESH_HAS_S) {
            float3 sv = b0 * ldv3(sc.s, fv + i0) + b1 * ldv3(sc.s, fv + i1) + b2 * ldv3(sc.s, fv + i2);
            if (len2(sv) != 0.0f) ss = sv;
        }
        float3 ts = cross3(ns, ss);
        if (len2(ts) > 0.0f) ss = cross3(ts, ns); else coord_system(ns, ss, ts);
        s.sn = ns;
        s.n = faceforward3(s.n, s.sn);                                      // orientation_is_authoritative = true
        s.sdpdu = ss; s.sdpdv = ts;
        while (len2(s.sdpdu) > 1e16f || len2(s.sdpdv) > 1e16f) { s.sdpdu = s.sdpdu / 1e8f; s.sdpdv = s.sdpdv / 1e8f; }
    }
    if (TEX) surf_tex_store<TEX>(x, make_float2(b0 * uv0.x + b1 * uv1.x + b2 * uv2.x, b0 * uv0.y + b1 * uv1.y + b2 * uv2.y), dpdu, dpdv, dndu, dndv);
    return s;
}
// bump_map with the constant displacement texture of this path (material.rs:1477-1509) followed by
// set_shading_geometry(.., false) (interaction.rs:229-250,379-405)
SGD void apply_constant_bump(Surf& s) {
    float3 ns = normalize3(cross3(s.sdpdu, s.sdpdv));
    s.sn = faceforward3(ns, s.n);
    while (len2(s.sdpdu) > 1e16f || len2(s.sdpdv) > 1e16f) { s.sdpdu = s.sdpdu / 1e8f; s.sdpdv = s.sdpdv / 1e8f; }
}

// ---------------- lights (light.rs) + Triangle sampling (triangle.rs:548-745) ----------------
struct LightCtx { P3fi pi; float3 n, ns; };
struct LightSample { Spec l; float3 wi; float pdf; P3fi p_light; float3 n_light; };

SGD Spec light_l(const DScene& sc, const SgLight& lt, float3 n, float3 w, const Wavelengths& lam) {     // light.rs:670-684
    if (!lt.two_sided && dot3(n, w) < 0.0f) return spec1(0.0f);
    return lt.scale * spectrum_sample(sc, lt.spectrum, lam);
}
SGD float tri_area(const TriGeo& g) { return 0.5f * len3(cross3(g.p1 - g.p0, g.p2 - g.p0)); }           // triangle.rs:543-546
SGD float tri_solid_angle(const TriGeo& g, float3 p) {                                                    // :162-169
    return spherical_tri_area(normalize3(g.p0 - p), normalize3(g.p1 - p), normalize3(g.p2 - p));
}
SGD bool tri_sample_with_context(const DScene& sc, const TriGeo& g, const LightCtx& ctx, float2 u, P3fi& out_pi, float3& out_n, float& out_pdf) {
    const float3 cp = p3fi_mid(ctx.pi);
    const float sa = tri_solid_angle(g, cp);
    if (sa < 3e-4f || sa > 6.22f) {
        // Triangle::sample (:548-589) then area -> solid-angle pdf (:604-620)
        float bb0, bb1;
        if (u.x < u.y) { bb0 = u.x / 2.0f; bb1 = u.y - bb0; } else { bb1 = u.y / 2.0f; bb0 = u.x - bb1; }   // sample_uniform_triangle sampling.rs:373-384
        float bb2 = 1.0f - bb1 - bb0;
        float3 p = bb0 * g.p0 + bb1 * g.p1 + bb2 * g.p2;
        float3 n = normalize3(cross3(g.p1 - g.p0, g.p2 - g.p0));
        if (!(g.flags & SG_MESH_HAS_N)) n = n * -1.0f;                 // :558-560 always negated (reference quirk)
        else { uint32_t i0, i1, i2; size_t fv; geo_indices(sc, g, i0, i1, i2, fv);
               float3 ns = bb0 * ldv3(sc.n, fv + i0) + bb1 * ldv3(sc.n, fv + i1) + bb2 * ldv3(sc.n, fv + i2); n = faceforward3(n, ns); }
        float3 p_abs_sum = abs3(bb0 * g.p0) + abs3(bb1 * g.p1) + abs3(bb2 * g.p2);
        out_pi = p3fi_make(p, gamma_n(6) * p_abs_sum);
        out_n = n;
        float pdf = 1.0f / tri_area(g);
        float3 sp = p3fi_mid(out_pi);
        float3 wi = sp - cp;
        if (len2(wi) == 0.0f) return false;
        wi = normalize3(wi);
        pdf /= absdot3(n, -wi) / dist2(cp, sp);
        if (isinf(pdf)) return false;
        out_pdf = pdf;
        return true;
    }
    float pdf = 1.0f;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        float3 w0 = normalize3(g.p0 - cp), w1 = normalize3(g.p1 - cp), w2 = normalize3(g.p2 - cp);
        float w[4] = {fmaxf(0.01f, absdot3(ctx.ns, w1)), fmaxf(0.01f, absdot3(ctx.ns, w1)),
                      fmaxf(0.01f, absdot3(ctx.ns, w0)), fmaxf(0.01f, absdot3(ctx.ns, w2))};
        float2 uw = sample_bilinear(u, w);      // warped u only feeds the pdf (:642-644): reference quirk kept
        pdf = bilinear_pdf(uw, w);
    }
    float b0, b1, b2, tri_pdf;
    sample_spherical_triangle(g.p0, g.p1, g.p2, cp, u, b0, b1, b2, tri_pdf);
    if (tri_pdf == 0.0f) return false;
    pdf = pdf * tri_pdf;
    float3 p_abs_sum = abs3(b0 * g.p0) + abs3(b1 * g.p1) + abs3((1.0f - b0 - b1) * g.p2);
    float3 p = b0 * g.p0 + b1 * g.p1 + b2 * g.p2;
    float3 n = normalize3(cross3(g.p1 - g.p0, g.p2 - g.p0));
    if (g.flags & SG_MESH_HAS_N) { uint32_t i0, i1, i2; size_t fv; geo_indices(sc, g, i0, i1, i2, fv);
                                   float3 ns = b0 * ldv3(sc.n, fv + i0) + b1 * ldv3(sc.n, fv + i1) + b2 * ldv3(sc.n, fv + i2); n = faceforward3(n, ns); }
    else if (((g.flags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((g.flags & SG_MESH_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    out_pi = p3fi_make(p, gamma_n(6) * p_abs_sum); out_n = n; out_pdf = pdf;
    return true;
}
// Triangle::pdf_with_context :696-745 (mesh_id/tri needed for the rare area-sampling branch)
SGD float tri_pdf_with_context(const DScene& sc, const TriGeo& g, const LightCtx& ctx, float3 wi) {
    const float3 cp = p3fi_mid(ctx.pi);
    const float sa = tri_solid_angle(g, cp);
    if (sa < 3e-4f || sa > 6.22f) {
        float3 o = offset_ray_origin(ctx.pi, ctx.n, wi);               // ShapeSampleContext::spawn_ray shape.rs:276-283
        RayPre rp = ray_precompute(wi);
        float b0, b1, b2, t;
        if (!intersect_triangle(o, rp, INFINITY, g.p0, g.p1, g.p2, b0, b1, b2, t)) return 0.0f;
        Surf s = make_surface(sc, g, b0, b1, b2);
        float pdf = (1.0f / tri_area(g)) / (absdot3(s.n, -wi) / dist2(cp, p3fi_mid(s.pi)));
        if (isinf(pdf)) return 0.0f;
        return pdf;
    }
    float pdf = 1.0f / sa;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        float2 u = invert_spherical_triangle_sample(g.p0, g.p1, g.p2, cp, wi);
        float3 w0 = normalize3(g.p0 - cp), w1 = normalize3(g.p1 - cp), w2 = normalize3(g.p2 - cp);
        float w[4] = {fmaxf(0.01f, absdot3(ctx.ns, w1)), fmaxf(0.01f, absdot3(ctx.ns, w1)),
                      fmaxf(0.01f, absdot3(ctx.ns, w0)), fmaxf(0.01f, absdot3(ctx.ns, w2))};
        pdf *= bilinear_pdf(u, w);
    }
    return pdf;
}
// Sphere emitters (sphere.rs:299-457): defined in sg_sphere_surface.cuh (they need the sphere's SurfaceInteraction)
static __device__ bool sphere_sample_with_context(const DSphere& S, const LightCtx& ctx, float2 u, P3fi& out_pi, float3& out_n, float& out_pdf);
static __device__ float sphere_pdf_with_context(const DScene& sc, const DSphere& S, const LightCtx& ctx, float3 wi);
// Everything that is not a triangle emitter goes through ONE out-of-line call per routine (defined in sg_patch_light.cuh, after
// the sphere / patch / environment-map code it dispatches to): the shade kernels' register allocation is sensitive to the number
// of call sites on the hot path, and triangle area lights are the common case.
// (they take the light's index, not the SgLight copy the caller holds: a by-reference struct argument would pin all 64 bytes in local memory)
static __device__ bool light_sample_li_other(const DScene& sc, uint32_t light_id, const LightCtx& ctx, float2 u, const Wavelengths& lam,
                                      bool allow_incomplete, LightSample& ls);
static __device__ float light_pdf_li_other(const DScene& sc, uint32_t light_id, uint32_t hit_mesh_word, const LightCtx& ctx, float3 wi);
static __device__ float infinite_pdf_li(const DScene& sc, const SgLight& lt, float3 wi, bool allow_incomplete);
static __device__ Spec infinite_le(const DScene& sc, const SgLight& lt, float3 ray_d, const Wavelengths& lam);
// Light::sample_li; allow_incomplete_pdf = true from PathIntegrator::sample_ld (integrator.rs:933), false from SimplePath (:652-656).
// GENERAL = false: the lean shade kernels, which only ever run on scenes whose lights are triangle emitters (plus uniform infinite
// lights, which sample_li never samples with incomplete pdfs, light.rs:748-750) -- no call site at all: even a never-taken
// out-of-line call costs those kernels ~3 % (measured on C2: register allocation around the call).
template <bool GENERAL = true>
SGD bool light_sample_li(const DScene& sc, uint32_t light_id, const SgLight& lt, const LightCtx& ctx, float2 u, const Wavelengths& lam, LightSample& ls,
                         bool allow_incomplete = true) {
    if (lt.kind != SG_LIGHT_DIFFUSE_AREA) {
        if constexpr (GENERAL) return light_sample_li_other(sc, light_id, ctx, u, lam, allow_incomplete, ls);
        else return false;
    }
    P3fi pi; float3 n; float pdf;                                            // DiffuseAreaLight::sample_li light.rs:632-661 over a Triangle
    const TriGeo g = geo_from_light(sc, light_id, lt);
    if (!tri_sample_with_context(sc, g, ctx, u, pi, n, pdf)) return false;
    float3 sp = p3fi_mid(pi), cp = p3fi_mid(ctx.pi);
    if (pdf == 0.0f || len2(sp - cp) == 0.0f) return false;
    float3 wi = normalize3(sp - cp);
    Spec le = light_l(sc, lt, n, -wi, lam);
    if (spec_zero(le)) return false;
    ls.l = le; ls.wi = wi; ls.pdf = pdf; ls.p_light = pi; ls.n_light = n;
    return true;
}
template <bool GENERAL = true>
SGD float light_pdf_li(const DScene& sc, uint32_t light_id, const SgLight& lt, const TriGeo& g, const LightCtx& ctx, float3 wi) {
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA) return tri_pdf_with_context(sc, g, ctx, wi);                   // light.rs:663-666
    if constexpr (GENERAL) return light_pdf_li_other(sc, light_id, g.mesh, ctx, wi);
    else return 0.0f;
}

// ---------------- camera (camera.rs:1003-1079, transform.rs:385-457,515-532,753-776) ----------------
SGD float3 xform_point(const float* m, float3 p) {
    float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1.0f) return f3(xp, yp, zp);
    return f3(xp, yp, zp) / wp;
}
SGD float3 xform_vector(const float* m, float3 v) {
    return f3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
SGD void xform_ray(const float* m, float3& o, float3& d) {
    // Transform::apply_ray (transform.rs:515-532): `self.apply(val.o)` is the Point3f overload (apply_point_helper :753-767, summed
    // left to right) and `.into()` makes a ZERO-width Point3fi, so the error-bound shift is dt = 0; what remains of it is the
    // interval addition o + (d * dt) (lo rounded down, hi rounded up, interval.rs:353-356) and the midpoint taken by `o.into()`.
    // Only apply_ray_inverse (:701-723) carries a real error term.
    const float x = o.x, y = o.y, z = o.z;
    const float xp = ((m[0] * x + m[1] * y) + m[2] * z) + m[3];
    const float yp = ((m[4] * x + m[5] * y) + m[6] * z) + m[7];
    const float zp = ((m[8] * x + m[9] * y) + m[10] * z) + m[11];
    P3fi oi = p3fi_exact(f3(xp, yp, zp));
    float3 dd = xform_vector(m, d);
    float ls = len2(dd);
    if (ls > 0.0f) {
        float dt = dot3(abs3(dd), p3fi_err(oi)) / ls;
        float3 off = dd * dt;
        oi.lo = f3(next_down(oi.lo.x + off.x), next_down(oi.lo.y + off.y), next_down(oi.lo.z + off.z));   // interval.rs:353-356
        oi.hi = f3(next_up(oi.hi.x + off.x), next_up(oi.hi.y + off.y), next_up(oi.hi.z + off.z));
    }
    o = p3fi_mid(oi); d = dd;
}
// evaluate_pixel_sample's camera stage (integrator.rs:339-362) + get_camera_sample (sampling.rs:347-371)
// + BoxFilter::sample (filter.rs:99-105) + PerspectiveCamera::generate_ray_differential (main ray)
struct AuxRays;
SGD void camera_aux(const DScene& sc, float3 p_camera, float2 p_lens, float3 o_cam, AuxRays* aux);      // sg_texture.cuh
SGD void camera_stage(const DScene& sc, uint32_t option_flags, int px, int py, Rng& rng, Wavelengths& lam, float3& o, float3& d, float& weight,
                      AuxRays* aux = nullptr) {
    float lu = (option_flags & SG_OPT_DISABLE_WAVELENGTH_JITTER) ? 0.5f : rng.get_1d();
    lam = sample_visible(lu);
    float2 pu; pu.x = rng.get_1d(); pu.y = rng.get_1d();
    float2 p_film, p_lens;
    if (option_flags & SG_OPT_DISABLE_PIXEL_JITTER) {
        p_film = make_float2((float)px + 0.5f, (float)py + 0.5f);
        p_lens = make_float2(0.5f, 0.5f);
    } else {
        float rx = sc.film.filter_radius[0], ry = sc.film.filter_radius[1];
        p_film = make_float2((float)px + lerpf(pu.x, -rx, rx) + 0.5f, (float)py + lerpf(pu.y, -ry, ry) + 0.5f);
        p_lens.x = rng.get_1d(); p_lens.y = rng.get_1d();
        (void)rng.get_1d();                                            // time sample (static scenes)
    }
    weight = 1.0f;
    float3 p_camera = xform_point(sc.camera.camera_from_raster, f3(p_film.x, p_film.y, 0.0f));
    if (sc.camera.kind == SG_CAMERA_ORTHOGRAPHIC) {                    // OrthographicCamera::generate_ray_differential camera.rs:760-784
        o = p_camera; d = f3(0.0f, 0.0f, 1.0f);                        // sic: stays in CAMERA space (see SgCameraKind)
        if (aux) camera_aux(sc, p_camera, p_lens, o, aux);
        return;
    }
    o = f3(0.0f, 0.0f, 0.0f);
    d = normalize3(p_camera);
    if (sc.camera.lens_radius > 0.0f) {
        float2 pl = sample_disk_concentric(p_lens);
        pl.x = sc.camera.lens_radius * pl.x; pl.y = sc.camera.lens_radius * pl.y;
        float ft = sc.camera.focal_distance / d.z;
        float3 p_focus = o + d * ft;
        o = f3(pl.x, pl.y, 0.0f);
        d = normalize3(p_focus - o);
    }
    if (aux) camera_aux(sc, p_camera, p_lens, o, aux);
    xform_ray(sc.camera.render_from_camera, o, d);
}

// ---------------- film: PixelSensor::to_sensor_rgb (film.rs:907-914) + RgbFilm::add_sample (:548-574) --------
// PixelSensor::to_sensor_rgb film.rs:907-914 + the clamp of RgbFilm::add_sample film.rs:548-567: the sensor RGB one
// sample contributes (before the filter weight).
SGD void film_sample_rgb(const DScene& sc, Spec L, const Wavelengths& lam, float rgb[3]) {
    Spec l = make_float4(lam.pdf.x != 0.0f ? L.x / lam.pdf.x : 0.0f, lam.pdf.y != 0.0f ? L.y / lam.pdf.y : 0.0f,
                         lam.pdf.z != 0.0f ? L.z / lam.pdf.z : 0.0f, lam.pdf.w != 0.0f ? L.w / lam.pdf.w : 0.0f);   // safe_div
    const int ids[3] = {sc.film.r_bar, sc.film.g_bar, sc.film.b_bar};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Spec s = spectrum_sample(sc, ids[c], lam) * l;
        float sum = 0.0f; sum = sum + s.x; sum = sum + s.y; sum = sum + s.z; sum = sum + s.w;
        rgb[c] = (sum / 4.0f) * sc.film.imaging_ratio;
    }
    float m = fmaxf(fmaxf(rgb[0], rgb[1]), rgb[2]);
    if (m > sc.film.max_component_value) { for (int c = 0; c < 3; ++c) rgb[c] = rgb[c] * sc.film.max_component_value / m; }
}

}  // namespace sg
