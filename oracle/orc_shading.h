// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).  parity status: see orc_math.h.
// Spectra, sampling routines, BxDFs, lights, camera and film of the CPU restatement.
#pragma once
#include "orc_scene.h"

namespace orc {

// ---- SampledSpectrum / SampledWavelengths (spectra/mod.rs:17: 4 samples) ----
struct Spec { Float v[4]; };
inline Spec spec_const(Float c) { Spec s = {{c, c, c, c}}; return s; }
inline Spec operator+(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] + b.v[i]; return r; }
inline Spec operator*(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * b.v[i]; return r; }
inline Spec operator*(Spec a, Float s) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] * s; return r; }
inline Spec operator*(Float s, Spec a) { return a * s; }
inline Spec operator/(Spec a, Float s) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = a.v[i] / s; return r; }
inline bool spec_is_zero(Spec a) { return a.v[0] == 0.0f && a.v[1] == 0.0f && a.v[2] == 0.0f && a.v[3] == 0.0f; }   // sampled_spectrum.rs:35-37
inline Float spec_max(Spec a) { Float m = std::nanf(""); for (int i = 0; i < 4; ++i) m = fmax_(m, a.v[i]); return m; }   // :113-118
inline Spec spec_clamp(Spec a, Float lo, Float hi) { Spec r; for (int i = 0; i < 4; ++i) r.v[i] = clampf(a.v[i], lo, hi); return r; }

struct Wavelengths { Float lambda[4]; Float pdf[4]; };

// sampling.rs:268-278
inline Float sample_visible_wavelengths(Float u) { return 538.0f - 138.888889f * std::atanh(0.85691062f - 1.82750197f * u); }
inline Float visible_wavelengths_pdf(Float lambda) {
    if (lambda < 360.0f || lambda > 830.0f) return 0.0f;
    Float x = std::cosh(0.0072f * (lambda - 538.0f));
    return 0.0039398042f / (x * x);
}
// sampled_wavelengths.rs:57-71
inline Wavelengths sample_visible(Float u) {
    Wavelengths w;
    for (int i = 0; i < 4; ++i) {
        Float up = u + (Float)i / 4.0f;
        if (up > 1.0f) up -= 1.0f;
        w.lambda[i] = sample_visible_wavelengths(up);
        w.pdf[i] = visible_wavelengths_pdf(w.lambda[i]);
    }
    return w;
}
// sampled_wavelengths.rs:79-96
inline bool secondary_terminated(const Wavelengths& w) { for (int i = 1; i < 4; ++i) if (w.pdf[i] != 0.0f) return false; return true; }
inline void terminate_secondary(Wavelengths& w) {
    if (secondary_terminated(w)) return;
    for (int i = 1; i < 4; ++i) w.pdf[i] = 0.0f;
    w.pdf[0] /= 4.0f;
}

// Rust `as i32` saturates and maps NaN to 0.
inline int32_t f2i_sat(Float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
// spectrum.rs:462-476
inline Float blackbody(Float lambda, Float temperature) {
    if (temperature < 0.0f) return 0.0f;
    const Float c = 299792458.0f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
    Float l = lambda * 1e-9f;
    Float l2 = l * l; Float l5 = l2 * l2 * l;      // powi(5): llvm.powi expands to repeated multiplication
    return (2.0f * h * c * c) / (l5 * (std::exp((h * c) / (l * kb * temperature)) - 1.0f));
}
// Spectrum::get: spectrum.rs:52-62 dispatch over :157, :264-270, :408-423, :480-482
inline Float spectrum_get(const SgSceneDesc* D, int id, Float lambda) {
    const SgSpectrum& s = D->spectra[id];
    const float* pool = D->spectrum_pool;
    switch (s.kind) {
    case SG_SPECTRUM_CONSTANT: return s.c;
    case SG_SPECTRUM_DENSE: {
        int32_t offset = f2i_sat(lambda) - s.lambda_min;          // `lambda as i32` truncates
        if (offset < 0 || offset >= s.n) return 0.0f;
        return pool[s.off_a + offset];
    }
    case SG_SPECTRUM_PIECEWISE_LINEAR: {
        const float* L = pool + s.off_a; const float* V = pool + s.off_b;
        if (s.n == 0 || lambda < L[0] || lambda > L[s.n - 1]) return 0.0f;
        int o = find_interval(s.n, [&](int i) { return L[i] <= lambda; });
        Float t = (lambda - L[o]) / (L[o + 1] - L[o]);
        return lerp(t, V[o], V[o + 1]);
    }
    case SG_SPECTRUM_BLACKBODY: return blackbody(lambda, s.c) * s.scale;
    }
    return 0.0f;
}
// Spectrum::sample: :165-167, :280-291 (ROUND to nearest nm, half away from zero), :433-439, :488-494
inline Spec spectrum_sample(const SgSceneDesc* D, int id, const Wavelengths& w) {
    const SgSpectrum& s = D->spectra[id];
    Spec r;
    if (s.kind == SG_SPECTRUM_DENSE) {
        for (int i = 0; i < 4; ++i) {
            int32_t offset = f2i_sat(std::round(w.lambda[i])) - s.lambda_min;
            r.v[i] = (offset < 0 || offset >= s.n) ? 0.0f : D->spectrum_pool[s.off_a + offset];
        }
        return r;
    }
    for (int i = 0; i < 4; ++i) r.v[i] = spectrum_get(D, id, w.lambda[i]);
    return r;
}

// ---- sampling.rs ---------------------------------------------------------------
inline Float power_heuristic(Float f_pdf, Float g_pdf) {     // :187-194 (nf = ng = 1)
    Float f = 1.0f * f_pdf, g = 1.0f * g_pdf;
    if (std::isinf(sqr(f))) return 1.0f;
    return (f * f) / (f * f + g * g);
}
inline Float sample_linear(Float u, Float a, Float b) {      // :250-257
    if (u == 0.0f && a == 0.0f) return 0.0f;
    Float x = u * (a + b) / (a + std::sqrt(lerp(u, a * a, b * b)));
    return fmin_(x, 1.0f - 1.1920929e-07f);
}
inline V2 sample_bilinear(V2 u, const Float w[4]) {          // :386-393
    V2 p;
    p.y = sample_linear(u.y, w[0] + w[1], w[2] + w[3]);
    p.x = sample_linear(u.x, lerp(p.y, w[0], w[2]), lerp(p.y, w[1], w[3]));
    return p;
}
inline Float bilinear_pdf(V2 p, const Float w[4]) {          // :395-408
    if (p.x < 0.0f || p.x > 1.0f || p.y < 0.0f || p.y > 1.0f) return 0.0f;
    if (w[0] + w[1] + w[2] + w[3] == 0.0f) return 1.0f;
    return 4.0f * ((1.0f - p.x) * (1.0f - p.y) * w[0] + p.x * (1.0f - p.y) * w[1] + (1.0f - p.x) * p.y * w[2] + p.x * p.y * w[3])
           / (w[0] + w[1] + w[2] + w[3]);
}
inline V2 sample_uniform_disk_concentric(V2 u) {             // :324-339
    V2 o = {2.0f * u.x - 1.0f, 2.0f * u.y - 1.0f};
    if (o.x == 0.0f && o.y == 0.0f) { V2 z = {0.0f, 0.0f}; return z; }
    Float r, theta;
    if (std::fabs(o.x) > std::fabs(o.y)) { r = o.x; theta = PI_OVER_4 * (o.y / o.x); }
    else { r = o.y; theta = PI_OVER_2 - PI_OVER_4 * (o.x / o.y); }
    V2 p = {r * std::cos(theta), r * std::sin(theta)};
    return p;
}
inline V2 sample_uniform_disk_polar(V2 u) {                  // :341-345
    Float r = std::sqrt(u.x), theta = 2.0f * PI_F * u.y;
    V2 p = {r * std::cos(theta), r * std::sin(theta)};
    return p;
}
inline V3 sample_cosine_hemisphere(V2 u) {                   // :310-318
    V2 d = sample_uniform_disk_concentric(u);
    Float z = safe_sqrt(1.0f - sqr(d.x) - sqr(d.y));
    return v3(d.x, d.y, z);
}
inline void sample_uniform_triangle(V2 u, Float b[3]) {      // :373-384
    Float b0, b1;
    if (u.x < u.y) { b0 = u.x / 2.0f; b1 = u.y - b0; }
    else { b1 = u.y / 2.0f; b0 = u.x - b1; }
    b[0] = b0; b[1] = b1; b[2] = 1.0f - b1 - b0;
}
// :412-499, including the reference's `divisor = e1.dot(e1)` (pbrt: dot(s1,e1)) and
// `(b1 / b1 + b2, b2 / b1 + b2)` renormalisation -- both are reproduced on purpose.
inline void sample_spherical_triangle(const V3 v[3], V3 p, V2 u, Float bary[3], Float* pdf_out) {
    V3 a = normalize(v[0] - p), b = normalize(v[1] - p), c = normalize(v[2] - p);
    V3 n_ab = cross(a, b), n_bc = cross(b, c), n_ca = cross(c, a);
    if (length_squared(n_ab) == 0.0f || length_squared(n_bc) == 0.0f || length_squared(n_ca) == 0.0f) {
        bary[0] = bary[1] = bary[2] = 0.0f; *pdf_out = 0.0f; return;
    }
    n_ab = normalize(n_ab); n_bc = normalize(n_bc); n_ca = normalize(n_ca);
    Float alpha = angle_between(n_ab, -n_ca), beta = angle_between(n_bc, -n_ab), gam = angle_between(n_ca, -n_bc);
    Float a_pi = alpha + beta + gam;
    Float ap_pi = lerp(u.x, PI_F, a_pi);
    Float area = a_pi - PI_F;
    Float pdf = area <= 0.0f ? 0.0f : 1.0f / area;
    Float cos_alpha = std::cos(alpha), sin_alpha = std::sin(alpha);
    Float sin_phi = std::sin(ap_pi) * cos_alpha - std::cos(ap_pi) * sin_alpha;
    Float cos_phi = std::cos(ap_pi) * cos_alpha + std::sin(ap_pi) * sin_alpha;
    Float k1 = cos_phi + cos_alpha;
    Float k2 = sin_phi - sin_alpha * dot(a, b);
    Float cos_bp = (k2 + (difference_of_products(k2, cos_phi, k1, sin_phi)) * cos_alpha) / (sum_of_products(k2, sin_phi, k1, cos_phi) * sin_alpha);
    cos_bp = clampf(cos_bp, -1.0f, 1.0f);
    Float sin_bp = safe_sqrt(1.0f - cos_bp * cos_bp);
    V3 cp = cos_bp * a + sin_bp * normalize(gram_schmidt(c, a));
    Float cos_theta = 1.0f - u.y * (1.0f - dot(cp, b));
    Float sin_theta = safe_sqrt(1.0f - cos_theta * cos_theta);
    V3 w = cos_theta * b + sin_theta * normalize(gram_schmidt(cp, b));
    V3 e1 = v[1] - v[0], e2 = v[2] - v[0];
    V3 s1 = cross(w, e2);
    Float divisor = dot(e1, e1);
    if (divisor == 0.0f) { bary[0] = bary[1] = bary[2] = 1.0f / 3.0f; *pdf_out = pdf; return; }
    Float inv_divisor = 1.0f / divisor;
    V3 s = p - v[0];
    Float b1 = dot(s, s1) * inv_divisor;
    Float b2 = dot(w, cross(s, e1)) * inv_divisor;
    b1 = clampf(b1, 0.0f, 1.0f); b2 = clampf(b2, 0.0f, 1.0f);
    if (b1 + b2 > 1.0f) { Float nb1 = b1 / b1 + b2, nb2 = b2 / b1 + b2; b1 = nb1; b2 = nb2; }
    bary[0] = 1.0f - b1 - b2; bary[1] = b1; bary[2] = b2; *pdf_out = pdf;
}
// :581-641
inline V2 invert_spherical_triangle_sample(const V3 v[3], V3 p, V3 w) {
    V3 a = normalize(v[0] - p), b = normalize(v[1] - p), c = normalize(v[2] - p);
    V3 n_ab = cross(a, b), n_bc = cross(b, c), n_ca = cross(c, a);
    V2 zero = {0.0f, 0.0f};
    if (length_squared(n_ab) == 0.0f || length_squared(n_bc) == 0.0f || length_squared(n_ca) == 0.0f) return zero;
    n_ab = normalize(n_ab); n_bc = normalize(n_bc); n_ca = normalize(n_ca);
    Float alpha = angle_between(n_ab, -n_ca), beta = angle_between(n_bc, -n_ab), gam = angle_between(n_ca, -n_bc);
    V3 cp = normalize(cross(cross(b, w), cross(c, a)));
    if (dot(cp, a + c) < 0.0f) cp = -cp;
    Float u0;
    if (dot(a, cp) > 0.99999847691f) u0 = 0.0f;
    else {
        V3 n_cpb = cross(cp, b), n_acp = cross(a, cp);
        if (length_squared(n_cpb) == 0.0f || length_squared(n_acp) == 0.0f) { V2 h = {0.5f, 0.5f}; return h; }
        n_cpb = normalize(n_cpb); n_acp = normalize(n_acp);
        Float ap = alpha + angle_between(n_ab, n_cpb) + angle_between(n_acp, -n_cpb) - PI_F;
        Float area = alpha + beta + gam - PI_F;
        u0 = ap / area;
    }
    Float u1 = (1.0f - dot(w, b)) / (1.0f - dot(cp, b));
    V2 r = {clampf(u0, 0.0f, 1.0f), clampf(u1, 0.0f, 1.0f)};
    return r;
}

// ---- scattering.rs ---------------------------------------------------------------
inline Float cos_theta(V3 w) { return w.z; }
inline Float cos2_theta(V3 w) { return w.z * w.z; }
inline Float abs_cos_theta(V3 w) { return std::fabs(w.z); }
inline Float sin2_theta(V3 w) { return fmax_(0.0f, 1.0f - cos2_theta(w)); }
inline Float sin_theta(V3 w) { return std::sqrt(sin2_theta(w)); }
inline Float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
inline Float cos_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.x / s, -1.0f, 1.0f); }   // spherical.rs:60-67
inline Float sin_phi(V3 w) { Float s = sin_theta(w); return s == 0.0f ? 1.0f : clampf(w.y / s, -1.0f, 1.0f); }   // spherical.rs:69-76 (1.0, not 0.0: reference quirk)
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
inline V3 reflect(V3 wo, V3 n) { return -wo + 2.0f * dot(wo, n) * n; }                      // scattering.rs:12-14
inline bool refract(V3 wi, V3 n, Float eta, V3* wt, Float* etap) {                          // :21-43
    Float cos_theta_i = dot(n, wi);
    if (cos_theta_i < 0.0f) { eta = 1.0f / eta; cos_theta_i = -cos_theta_i; n = -n; }
    Float sin2_theta_i = fmax_(0.0f, 1.0f - sqr(cos_theta_i));
    Float sin2_theta_t = sin2_theta_i / sqr(eta);
    if (sin2_theta_t >= 1.0f) return false;
    Float cos_theta_t = std::sqrt(1.0f - sin2_theta_t);
    *wt = -wi / eta + (cos_theta_i / eta - cos_theta_t) * n;
    *etap = eta;
    return true;
}
inline Float fresnel_dielectric(Float cos_theta_i, Float eta) {                             // :49-70
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    if (cos_theta_i < 0.0f) { eta = 1.0f / eta; cos_theta_i = -cos_theta_i; }
    Float sin2_theta_i = 1.0f - cos_theta_i * cos_theta_i;
    Float sin2_theta_t = sin2_theta_i / (eta * eta);
    if (sin2_theta_t >= 1.0f) return 1.0f;
    Float cos_theta_t = safe_sqrt(1.0f - sin2_theta_t);
    Float r_parl = (eta * cos_theta_i - cos_theta_t) / (eta * cos_theta_i + cos_theta_t);
    Float r_perp = (cos_theta_i - eta * cos_theta_t) / (cos_theta_i + eta * cos_theta_t);
    return 0.5f * (r_parl * r_parl + r_perp * r_perp);
}
// num-complex 0.4.4 Complex<f32> (third-party; restated from its published source; parity unpinned)
struct Cx { Float re, im; };
inline Cx cx(Float r, Float i) { Cx c = {r, i}; return c; }
inline Cx cx_mul(Cx a, Cx b) { return cx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
inline Cx cx_div(Cx a, Cx b) { Float ns = b.re * b.re + b.im * b.im; return cx((a.re * b.re + a.im * b.im) / ns, (a.im * b.re - a.re * b.im) / ns); }
inline Cx cx_sqrt(Cx z) {
    if (z.im == 0.0f) {
        if (!std::signbit(z.re)) return cx(std::sqrt(z.re), z.im);
        Float im = std::sqrt(-z.re);
        return cx(0.0f, std::signbit(z.im) ? -im : im);
    } else if (z.re == 0.0f) {
        Float x = std::sqrt(std::fabs(z.im) / 2.0f);
        return cx(x, std::signbit(z.im) ? -x : x);
    }
    Float r = std::hypot(z.re, z.im), theta = std::atan2(z.im, z.re);
    Float sr = std::sqrt(r), ht = theta / 2.0f;
    return cx(sr * std::cos(ht), sr * std::sin(ht));
}
inline Float fresnel_complex(Float cos_theta_i, Cx eta) {                                   // :78-89
    cos_theta_i = clampf(cos_theta_i, 0.0f, 1.0f);
    Float sin2_theta_i = 1.0f - sqr(cos_theta_i);
    Cx sin2_theta_t = cx_div(cx(sin2_theta_i, 0.0f), cx_mul(eta, eta));
    Cx cos_theta_t = cx_sqrt(cx(1.0f - sin2_theta_t.re, 0.0f - sin2_theta_t.im));
    Cx eci = cx(eta.re * cos_theta_i, eta.im * cos_theta_i);
    Cx r_parl = cx_div(cx(eci.re - cos_theta_t.re, eci.im - cos_theta_t.im), cx(eci.re + cos_theta_t.re, eci.im + cos_theta_t.im));
    Cx ect = cx_mul(eta, cos_theta_t);
    Cx r_perp = cx_div(cx(cos_theta_i - ect.re, 0.0f - ect.im), cx(cos_theta_i + ect.re, 0.0f + ect.im));
    return ((r_parl.re * r_parl.re + r_parl.im * r_parl.im) + (r_perp.re * r_perp.re + r_perp.im * r_perp.im)) / 2.0f;
}
// TrowbridgeReitzDistribution :107-220
struct TR {
    Float ax, ay;
    static TR make(Float ax, Float ay) {
        TR d = {ax, ay};
        if (!d.effectively_smooth()) { d.ax = fmax_(d.ax, 1e-4f); d.ay = fmax_(d.ay, 1e-4f); }
        return d;
    }
    bool effectively_smooth() const { return ax < 1e-3f && ay < 1e-3f; }
    Float d(V3 wm) const {
        Float t2 = tan2_theta(wm);
        if (std::isinf(t2)) return 0.0f;
        Float cos4 = sqr(cos2_theta(wm));
        if (cos4 < 1e-16f) return 0.0f;
        Float e = t2 * (sqr(cos_phi(wm) / ax) + sqr(sin_phi(wm) / ay));
        return 1.0f / (PI_F * ax * ay * cos4 * sqr(1.0f + e));
    }
    Float lambda(V3 w) const {
        Float t2 = tan2_theta(w);
        if (std::isinf(t2)) return 0.0f;
        Float alpha2 = sqr(cos_phi(w) * ax) + sqr(sin_phi(w) * ay);
        return (-1.0f + std::sqrt(1.0f + alpha2 * t2)) / 2.0f;
    }
    Float g1(V3 w) const { return 1.0f / (1.0f + lambda(w)); }
    Float g(V3 wo, V3 wi) const { return 1.0f / (1.0f + lambda(wo) + lambda(wi)); }
    Float d_w(V3 w, V3 wm) const { return g1(w) / abs_cos_theta(w) * d(wm) * abs_dot(w, wm); }
    Float pdf(V3 w, V3 wm) const { return d_w(w, wm); }
    V3 sample_wm(V3 w, V2 u) const {
        V3 wh = normalize(v3(ax * w.x, ay * w.y, w.z));
        if (wh.z < 0.0f) wh = -wh;
        V3 t1 = wh.z < 0.99999f ? normalize(cross(v3(0, 0, 1), wh)) : v3(1, 0, 0);
        V3 t2 = cross(wh, t1);
        V2 p = sample_uniform_disk_polar(u);
        Float h = std::sqrt(1.0f - sqr(p.x));
        p.y = lerp((1.0f + wh.z) / 2.0f, h, p.y);
        Float pz = std::sqrt(fmax_(0.0f, 1.0f - (p.x * p.x + p.y * p.y)));
        V3 nh = p.x * t1 + p.y * t2 + pz * wh;
        return normalize(v3(ax * nh.x, ay * nh.y, fmax_(1e-6f, nh.z)));
    }
    void regularize() {
        if (ax < 0.3f) ax = clampf(2.0f * ax, 0.1f, 0.3f);
        if (ay < 0.3f) ay = clampf(2.0f * ay, 0.1f, 0.3f);
    }
};

// ---- BxDFs (bxdf.rs) + BSDF frame wrapper (bsdf.rs) ---------------------------
enum { BX_UNSET = 0, BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16 };   // bxdf.rs:1773-1789
struct BSDFSample { Spec f; V3 wi; Float pdf; int flags; Float eta; bool proportional = false; };

}  // namespace orc
#include "orc_layered.h"     // CoatedDiffuse = LayeredBxDF<Dielectric, Diffuse> (needs TR / BSDFSample above)
namespace orc {

struct BSDF {
    int kind;            // SgMaterialKind
    Spec r;              // diffuse reflectance / conductor eta
    Spec k;              // conductor k
    Float eta;           // dielectric
    TR mf;
    Layered lay;         // CoatedDiffuse
    uint64_t layer_seed = 0;   // seeds the LayeredBxDF's private generator for the NEXT f/sample_f/pdf call
    V3 fx, fy, fz;       // Frame::from_xz(normalize(dpdus), ns)  bsdf.rs:22-28, frame.rs:14-17
    Rng layer_rng() const { Rng r; r.seed_from_u64(layer_seed); return r; }

    V3 to_local(V3 v) const { return v3(dot(v, fx), dot(v, fy), dot(v, fz)); }          // frame.rs:39-41
    V3 from_local(V3 v) const { return v.x * fx + v.y * fy + v.z * fz; }               // frame.rs:51-53

    int flags() const {
        switch (kind) {
        case SG_MATERIAL_DIFFUSE: return spec_is_zero(r) ? BX_UNSET : (BX_DIFFUSE | BX_REFLECTION);                 // bxdf.rs:256-262
        case SG_MATERIAL_CONDUCTOR: return mf.effectively_smooth() ? (BX_SPECULAR | BX_REFLECTION) : (BX_GLOSSY | BX_REFLECTION);   // :447-453
        case SG_MATERIAL_COATED_DIFFUSE: case SG_MATERIAL_COATED_CONDUCTOR: return lay.flags();
        case SG_MATERIAL_THIN_DIELECTRIC: return BX_REFLECTION | BX_TRANSMISSION | BX_SPECULAR;                    // bxdf.rs:873-875
        default: {                                                                                                 // :778-790
            int f = (eta == 1.0f) ? BX_TRANSMISSION : (BX_REFLECTION | BX_TRANSMISSION);
            return f | (mf.effectively_smooth() ? BX_SPECULAR : BX_GLOSSY);
        }
        }
    }
    // local-space f
    Spec f_local(V3 wo, V3 wi) const {
        if (kind == SG_MATERIAL_COATED_DIFFUSE || kind == SG_MATERIAL_COATED_CONDUCTOR) { Rng r = layer_rng(); return lay.f(wo, wi, r); }
        if (kind == SG_MATERIAL_THIN_DIELECTRIC) return spec_const(0.0f);       // bxdf.rs:808-810
        switch (kind) {
        case SG_MATERIAL_DIFFUSE:                                               // bxdf.rs:196-202
            if (!same_hemisphere(wo, wi)) return spec_const(0.0f);
            return r * INV_PI;
        case SG_MATERIAL_CONDUCTOR: {                                           // :349-376
            if (!same_hemisphere(wo, wi)) return spec_const(0.0f);
            if (mf.effectively_smooth()) return spec_const(0.0f);
            Float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
            if (cti == 0.0f || cto == 0.0f) return spec_const(0.0f);
            V3 wm = wi + wo;
            if (length_squared(wm) == 0.0f) return spec_const(0.0f);
            wm = normalize(wm);
            Spec F; Float c = abs_dot(wo, wm);
            for (int i = 0; i < 4; ++i) F.v[i] = fresnel_complex(c, cx(r.v[i], k.v[i]));
            return mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
        }
        default: {                                                              // :533-584
            if (eta == 1.0f || mf.effectively_smooth()) return spec_const(0.0f);
            Float cto = cos_theta(wo), cti = cos_theta(wi);
            bool refl = cti * cto > 0.0f;
            Float etap = 1.0f;
            if (!refl) etap = cto > 0.0f ? eta : (1.0f / eta);
            V3 wm = wi * etap + wo;
            if (cti == 0.0f || cto == 0.0f || length_squared(wm) == 0.0f) return spec_const(0.0f);
            wm = face_forward(normalize(wm), v3(0, 0, 1));
            if (dot(wm, wi) * cti < 0.0f || dot(wm, wo) * cto < 0.0f) return spec_const(0.0f);
            Float F = fresnel_dielectric(dot(wo, wm), eta);
            if (refl) return spec_const(mf.d(wm) * mf.g(wo, wi) * F / std::fabs(4.0f * cti * cto));
            Float denom = sqr(dot(wi, wm) + dot(wo, wm) / etap) * cti * cto;
            Float ft = mf.d(wm) * (1.0f - F) * mf.g(wo, wi) * std::fabs(dot(wi, wm) * dot(wo, wm) / denom);
            ft /= sqr(etap);                                                    // TransportMode::Radiance
            return spec_const(ft);
        }
        }
    }
    Float pdf_local(V3 wo, V3 wi) const {
        if (kind == SG_MATERIAL_COATED_DIFFUSE || kind == SG_MATERIAL_COATED_CONDUCTOR) { Rng r = layer_rng(); return lay.pdf(wo, wi, r); }
        if (kind == SG_MATERIAL_THIN_DIELECTRIC) return 0.0f;                   // bxdf.rs:863-871
        switch (kind) {
        case SG_MATERIAL_DIFFUSE:                                               // :240-254
            if (!same_hemisphere(wo, wi)) return 0.0f;
            return abs_cos_theta(wi) * INV_PI;
        case SG_MATERIAL_CONDUCTOR: {                                           // :424-445
            if (!same_hemisphere(wo, wi) || mf.effectively_smooth()) return 0.0f;
            V3 wm = wo + wi;
            if (length_squared(wm) == 0.0f) return 0.0f;
            wm = face_forward(normalize(wm), v3(0, 0, 1));
            return mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm));
        }
        default: {                                                              // :715-776
            if (eta == 1.0f || mf.effectively_smooth()) return 0.0f;
            Float cto = cos_theta(wo), cti = cos_theta(wi);
            bool refl = cti * cto > 0.0f;
            Float etap = 1.0f;
            if (!refl) etap = cto > 0.0f ? eta : (1.0f / eta);
            V3 wm = wi * etap + wo;
            if (cti == 0.0f || cto == 0.0f || length_squared(wm) == 0.0f) return 0.0f;
            wm = face_forward(normalize(wm), v3(0, 0, 1));
            if (dot(wm, wi) * cti < 0.0f || dot(wm, wo) * cto < 0.0f) return 0.0f;
            Float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
            Float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return 0.0f;
            if (refl) return mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm)) * pr / (pr + pt);
            Float denom = sqr(dot(wi, wm) + dot(wo, wm) / etap);
            Float dwm_dwi = abs_dot(wi, wm) / denom;
            return mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
        }
        }
    }
    bool sample_local(V3 wo, Float uc, V2 u, BSDFSample* bs) const {
        bs->eta = 1.0f; bs->proportional = false;
        if (kind == SG_MATERIAL_COATED_DIFFUSE || kind == SG_MATERIAL_COATED_CONDUCTOR) { Rng r = layer_rng(); return lay.sample_f(wo, uc, u, r, bs, &bs->proportional); }
        if (kind == SG_MATERIAL_THIN_DIELECTRIC) {                              // ThinDielectricBxDF::sample_f bxdf.rs:812-861
            Float R = fresnel_dielectric(abs_cos_theta(wo), eta), T = 1.0f - R;
            if (R < 1.0f) { R += sqr(T) * R / (1.0f - sqr(R)); T = 1.0f - R; }
            const Float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return false;
            if (uc < pr / (pr + pt)) {
                V3 wi = v3(-wo.x, -wo.y, wo.z);
                bs->f = spec_const(R / abs_cos_theta(wi)); bs->wi = wi; bs->pdf = pr / (pr + pt); bs->flags = BX_SPECULAR | BX_REFLECTION;
            } else {
                V3 wi = -wo;
                bs->f = spec_const(T / abs_cos_theta(wi)); bs->wi = wi; bs->pdf = pt / (pr + pt); bs->flags = BX_SPECULAR | BX_TRANSMISSION;
            }
            return true;
        }
        switch (kind) {
        case SG_MATERIAL_DIFFUSE: {                                             // :204-238
            V3 wi = sample_cosine_hemisphere(u);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            bs->f = r * INV_PI; bs->wi = wi; bs->pdf = abs_cos_theta(wi) * INV_PI; bs->flags = BX_DIFFUSE | BX_REFLECTION;
            return true;
        }
        case SG_MATERIAL_CONDUCTOR: {                                           // :378-422
            if (mf.effectively_smooth()) {
                V3 wi = v3(-wo.x, -wo.y, wo.z);
                Spec F; for (int i = 0; i < 4; ++i) F.v[i] = fresnel_complex(abs_cos_theta(wi), cx(r.v[i], k.v[i]));
                bs->f = F / abs_cos_theta(wi); bs->wi = wi; bs->pdf = 1.0f; bs->flags = BX_SPECULAR | BX_REFLECTION;
                return true;
            }
            if (wo.z == 0.0f) return false;
            V3 wm = mf.sample_wm(wo, u);
            V3 wi = reflect(wo, wm);
            if (!same_hemisphere(wo, wi)) return false;
            Float pdf = mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm));
            Float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
            if (cti == 0.0f || cto == 0.0f) return false;
            Spec F; Float c = abs_dot(wo, wm);
            for (int i = 0; i < 4; ++i) F.v[i] = fresnel_complex(c, cx(r.v[i], k.v[i]));
            bs->f = mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti); bs->wi = wi; bs->pdf = pdf; bs->flags = BX_GLOSSY | BX_REFLECTION;
            return true;
        }
        default: {                                                              // :586-713
            if (eta == 1.0f || mf.effectively_smooth()) {
                Float R = fresnel_dielectric(cos_theta(wo), eta), T = 1.0f - R;
                Float pr = R, pt = T;
                if (pr == 0.0f && pt == 0.0f) return false;
                if (uc < pr / (pr + pt)) {
                    V3 wi = v3(-wo.x, -wo.y, wo.z);
                    bs->f = spec_const(R / abs_cos_theta(wi)); bs->wi = wi; bs->pdf = pr / (pr + pt); bs->flags = BX_SPECULAR | BX_REFLECTION;
                    return true;
                }
                V3 wi; Float etap;
                if (!refract(wo, v3(0, 0, 1), eta, &wi, &etap)) return false;
                Float ft = T / abs_cos_theta(wi);
                ft /= sqr(etap);
                bs->f = spec_const(ft); bs->wi = wi; bs->pdf = pt / (pr + pt); bs->flags = BX_SPECULAR | BX_TRANSMISSION; bs->eta = etap;
                return true;
            }
            V3 wm = mf.sample_wm(wo, u);
            Float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
            Float pr = R, pt = T;
            if (pr == 0.0f && pt == 0.0f) return false;
            if (uc < pr / (pr + pt)) {
                V3 wi = reflect(wo, wm);
                if (!same_hemisphere(wo, wi)) return false;
                Float pdf = mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm)) * pr / (pr + pt);
                bs->f = spec_const(mf.d(wm) * mf.g(wo, wi) * R / (4.0f * cos_theta(wi) * cos_theta(wo)));
                bs->wi = wi; bs->pdf = pdf; bs->flags = BX_GLOSSY | BX_REFLECTION;
                return true;
            }
            V3 wi; Float etap;
            if (!refract(wo, wm, eta, &wi, &etap)) return false;
            if (same_hemisphere(wo, wi) || wi.z == 0.0f) return false;
            Float denom = sqr(dot(wi, wm) + dot(wo, wm) / etap);
            Float dwm_dwi = abs_dot(wi, wm) / denom;
            Float pdf = mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
            Float ft = T * mf.d(wm) * mf.g(wo, wi) * std::fabs(dot(wi, wm) * dot(wo, wm) / (cos_theta(wi) * cos_theta(wo) * denom));
            ft /= sqr(etap);
            bs->f = spec_const(ft); bs->wi = wi; bs->pdf = pdf; bs->flags = BX_GLOSSY | BX_TRANSMISSION; bs->eta = etap;
            return true;
        }
        }
    }
    // bsdf.rs:44-58
    Spec f(V3 wo_r, V3 wi_r) const {
        V3 wi = to_local(wi_r), wo = to_local(wo_r);
        if (wo.z == 0.0f) return spec_const(0.0f);
        return f_local(wo, wi);
    }
    // bsdf.rs:60-82
    bool sample_f(V3 wo_r, Float uc, V2 u, BSDFSample* bs) const {
        V3 wo = to_local(wo_r);
        if (wo.z == 0.0f || !(flags() & (BX_REFLECTION | BX_TRANSMISSION))) return false;
        if (!sample_local(wo, uc, u, bs)) return false;
        if (spec_is_zero(bs->f) || bs->pdf == 0.0f || bs->wi.z == 0.0f) return false;
        bs->wi = from_local(bs->wi);
        return true;
    }
    // bsdf.rs:84-97
    Float pdf(V3 wo_r, V3 wi_r) const {
        V3 wo = to_local(wo_r), wi = to_local(wi_r);
        if (wo.z == 0.0f) return 0.0f;
        return pdf_local(wo, wi);
    }
};

}  // namespace orc
#include "orc_texture.h"
namespace orc {

// SurfaceInteraction::get_bsdf (interaction.rs:187-278) + Material::get_bsdf
// (material.rs:301-322, 456-511, 603-648, 917-963).
// `mix_seed` seeds the generator MixMaterial::choose_material draws from (material.rs:1309-1330); the reference uses a
// per-thread SmallRng::from_entropy() (integrator.rs:255) -- here: layer_seed(path stream, site 5), see SG_MATERIAL_MIX.
inline BSDF get_bsdf(const SgSceneDesc* D, SurfaceInteraction& si, Wavelengths& lambda, const AuxRays& aux, const SgRenderParams* rp, uint64_t mix_seed = 0) {
    // compute_differentials feeds image-texture filtering, the bump-map step and specular ray differentials; scenes
    // without image textures never read its results (constant textures ignore the footprint)
    if (D->n_textures > 0) compute_differentials(D, si, aux, rp->samples_per_pixel, rp->option_flags);
    if (D->materials[si.material].kind == SG_MATERIAL_MIX) {                     // interaction.rs:206-221
        Rng mr; mr.seed_from_u64(mix_seed);
        const TexCoordCtx mc = {si.uv, si.dudx, si.dudy, si.dvdx, si.dvdy, si.p(), si.dpdx, si.dpdy, si.n};
        for (int guard = 0; guard < 64 && D->materials[si.material].kind == SG_MATERIAL_MIX; ++guard) {
            const SgMaterial& mm = D->materials[si.material];
            const Float amt = mm.tex_mix_amount >= 0 ? eval_float_texture(D, mm.tex_mix_amount, mc) : mm.mix_amount;
            if (amt <= 0.0f) si.material = mm.mix_materials[0];
            else if (amt >= 1.0f) si.material = mm.mix_materials[1];
            else { const Float u = mr.get_1d(); si.material = amt < u ? mm.mix_materials[0] : mm.mix_materials[1]; }
        }
    }
    const SgMaterial& m0 = D->materials[si.material];
    if ((m0.flags & SG_MAT_HAS_DISPLACEMENT) || m0.normal_map >= 0) {              // interaction.rs:225-250
        // bump_map (material.rs:1477-1509) or, only without a displacement, normal_map (:1453-1474); then
        // set_shading_geometry(ns, dpdu, dpdv, dndu, dndv, false)
        V3 dpdu, dpdv;
        if (m0.flags & SG_MAT_HAS_DISPLACEMENT) {
            if (m0.tex_displacement >= 0 || m0.displacement != 0.0f) bump_map(D, m0.tex_displacement, m0.displacement, si, &dpdu, &dpdv);
            else { dpdu = si.sdpdu; dpdv = si.sdpdv; }    // constant 0: dpdu + 0/du*n + 0*dndu
        } else normal_map(D, m0.normal_map, si, &dpdu, &dpdv);
        V3 ns = normalize(cross(dpdu, dpdv));                                   // interaction.rs:246
        si.sn = face_forward(ns, si.n);                                          // :379-405
        si.sdpdu = dpdu; si.sdpdv = dpdv;
        while (length_squared(si.sdpdu) > 1e16f || length_squared(si.sdpdv) > 1e16f) { si.sdpdu = si.sdpdu / 1e8f; si.sdpdv = si.sdpdv / 1e8f; }
    }
    TexCoordCtx tc = {si.uv, si.dudx, si.dudy, si.dvdx, si.dvdy, si.p(), si.dpdx, si.dpdy, si.n};
    // texture-valued parameters (SgMaterialTextures): tex_eval.evaluate_float / evaluate_spectrum of material.rs:456-499, 603-635, 917-963, 1188-1260
    SgMaterial m = D->materials[si.material];
    Spec tex_a, tex_b, tex_d; bool has_a = false, has_b = false, has_d = false;
    if (D->material_textures) {
        const SgMaterialTextures& mt = D->material_textures[si.material];
        if (mt.u_roughness >= 0) m.u_roughness = eval_float_texture(D, mt.u_roughness, tc);
        if (mt.v_roughness >= 0) m.v_roughness = eval_float_texture(D, mt.v_roughness, tc);
        if (mt.thickness >= 0) m.thickness = eval_float_texture(D, mt.thickness, tc);
        if (mt.g >= 0) m.g = eval_float_texture(D, mt.g, tc);
        if (mt.u_roughness2 >= 0) m.u_roughness2 = eval_float_texture(D, mt.u_roughness2, tc);
        if (mt.v_roughness2 >= 0) m.v_roughness2 = eval_float_texture(D, mt.v_roughness2, tc);
        if (mt.spec_a >= 0) { tex_a = eval_spectrum_texture(D, mt.spec_a, tc, lambda); has_a = true; }
        if (mt.spec_b >= 0) { tex_b = eval_spectrum_texture(D, mt.spec_b, tc, lambda); has_b = true; }
        if (mt.spec_d >= 0) { tex_d = eval_spectrum_texture(D, mt.spec_d, tc, lambda); has_d = true; }
    }
    BSDF b;
    b.kind = m.kind; b.r = spec_const(0.0f); b.k = spec_const(0.0f); b.eta = 1.0f; b.mf = TR::make(0.0f, 0.0f);
    b.lay.mf = TR::make(0.0f, 0.0f); b.lay.mfb = TR::make(0.0f, 0.0f);
    if (m.kind == SG_MATERIAL_DIFFUSE) {
        b.r = spec_clamp(m.tex_reflectance >= 0 ? eval_spectrum_texture(D, m.tex_reflectance, tc, lambda) : spectrum_sample(D, m.spec_a, lambda), 0.0f, 1.0f);
    } else if (m.kind == SG_MATERIAL_CONDUCTOR) {
        Float ur = m.u_roughness, vr = m.v_roughness;
        if (m.flags & SG_MAT_REMAP_ROUGHNESS) { ur = std::sqrt(ur); vr = std::sqrt(vr); }   // roughness_to_alpha scattering.rs:197-199
        b.r = has_a ? tex_a : spectrum_sample(D, m.spec_a, lambda);
        b.k = has_b ? tex_b : spectrum_sample(D, m.spec_b, lambda);
        b.mf = TR::make(ur, vr);
    } else if (m.kind == SG_MATERIAL_COATED_DIFFUSE) {                       // material.rs:917-963
        b.lay.r = spec_clamp(m.tex_reflectance >= 0 ? eval_spectrum_texture(D, m.tex_reflectance, tc, lambda) : spectrum_sample(D, m.spec_a, lambda), 0.0f, 1.0f);
        Float ur = m.u_roughness, vr = m.v_roughness;
        if (m.flags & SG_MAT_REMAP_ROUGHNESS) { ur = std::sqrt(ur); vr = std::sqrt(vr); }
        b.lay.mf = TR::make(ur, vr);
        b.lay.thickness = m.thickness;
        Float sampled_eta = spectrum_get(D, m.spec_c, lambda.lambda[0]);
        if (D->spectra[m.spec_c].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lambda);
        if (sampled_eta == 0.0f) sampled_eta = 1.0f;
        b.lay.eta = sampled_eta;
        b.lay.albedo = spec_clamp(has_b ? tex_b : spectrum_sample(D, m.spec_b, lambda), 0.0f, 1.0f);
        b.lay.g = clampf(m.g, -1.0f, 1.0f);
        b.lay.max_depth = m.max_depth; b.lay.n_samples = m.n_samples;
        b.mf = b.lay.mf;
    } else if (m.kind == SG_MATERIAL_COATED_CONDUCTOR) {                     // material.rs:1188-1260
        Float iur = m.u_roughness, ivr = m.v_roughness;
        if (m.flags & SG_MAT_REMAP_ROUGHNESS) { iur = std::sqrt(iur); ivr = std::sqrt(ivr); }
        b.lay.mf = TR::make(iur, ivr);
        b.lay.thickness = m.thickness;
        Float ieta = spectrum_get(D, m.spec_c, lambda.lambda[0]);
        if (D->spectra[m.spec_c].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lambda);
        if (ieta == 0.0f) ieta = 1.0f;
        b.lay.eta = ieta;
        Spec ce, ck;
        if (!(m.flags & SG_MAT_CONDUCTOR_REFLECTANCE)) { ce = has_a ? tex_a : spectrum_sample(D, m.spec_a, lambda); ck = has_d ? tex_d : spectrum_sample(D, m.spec_d, lambda); }
        else {                                                               // :1225-1233
            Spec r = spec_clamp(has_a ? tex_a : spectrum_sample(D, m.spec_a, lambda), 0.0f, 0.9999f);
            ce = spec_const(1.0f);
            for (int i = 0; i < 4; ++i) ck.v[i] = 2.0f * std::sqrt(r.v[i]) / std::sqrt(fmax_(0.0f, 1.0f - r.v[i]));
        }
        ce = ce / ieta; ck = ck / ieta;
        Float cur = m.u_roughness2, cvr = m.v_roughness2;
        if (m.flags & SG_MAT_REMAP_ROUGHNESS) { cur = std::sqrt(iur); cvr = std::sqrt(ivr); }   // sic: roughness_to_alpha(iurough), material.rs:1237-1241
        b.lay.cond = true; b.lay.ce = ce; b.lay.ck = ck; b.lay.mfb = TR::make(cur, cvr);
        b.lay.r = spec_const(0.0f);
        b.lay.albedo = spec_clamp(has_b ? tex_b : spectrum_sample(D, m.spec_b, lambda), 0.0f, 1.0f);
        b.lay.g = clampf(m.g, -1.0f, 1.0f);
        b.lay.max_depth = m.max_depth; b.lay.n_samples = m.n_samples;
        b.mf = b.lay.mf;
    } else {
        Float sampled_eta = spectrum_get(D, m.spec_a, lambda.lambda[0]);
        if (D->spectra[m.spec_a].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lambda);
        if (sampled_eta == 0.0f) sampled_eta = 1.0f;
        Float ur = m.u_roughness, vr = m.v_roughness;
        if (m.flags & SG_MAT_REMAP_ROUGHNESS) { ur = std::sqrt(ur); vr = std::sqrt(vr); }
        b.eta = sampled_eta; b.mf = TR::make(ur, vr);
    }
    b.fx = normalize(si.sdpdu); b.fz = si.sn; b.fy = cross(b.fz, b.fx);
    return b;
}

}  // namespace orc
#include "orc_envmap.h"
namespace orc {

// ---- lights ----------------------------------------------------------------------
struct LightSampleContext { P3fi pi; V3 n, ns; V3 p() const { return p3fi_mid(pi); } };
struct LightLiSample { Spec l; V3 wi; Float pdf; P3fi p_light; V3 n_light; };

inline Spec light_l(const SgSceneDesc* D, const SgLight& lt, V3 n, V3 w, const Wavelengths& lambda) {   // light.rs:670-684
    if (!lt.two_sided && dot(n, w) < 0.0f) return spec_const(0.0f);
    return lt.scale * spectrum_sample(D, lt.spectrum, lambda);
}
inline Float tri_area(V3 p0, V3 p1, V3 p2) { return 0.5f * length(cross(p1 - p0, p2 - p0)); }          // triangle.rs:543-546
inline Float tri_solid_angle(V3 p0, V3 p1, V3 p2, V3 p) {                                                // :162-169
    return spherical_triangle_area(normalize(p0 - p), normalize(p1 - p), normalize(p2 - p));
}
struct ShapeSample { P3fi pi; V3 n; Float pdf; };

// Triangle::sample triangle.rs:548-589
inline void tri_sample_area(const Scene& sc, uint32_t mesh_id, uint32_t tri, V2 u, ShapeSample* ss) {
    const SgMesh& m = sc.d->meshes[mesh_id];
    uint32_t v[3]; sc.tri_indices(mesh_id, tri, v);
    V3 p0 = sc.vertex(m, v[0]), p1 = sc.vertex(m, v[1]), p2 = sc.vertex(m, v[2]);
    Float b[3]; sample_uniform_triangle(u, b);
    V3 p = b[0] * p0 + b[1] * p1 + b[2] * p2;
    V3 n = normalize(cross(p1 - p0, p2 - p0));
    if (!(m.flags & SG_MESH_HAS_N)) n = n * -1.0f;                       // :558-560 (always negated: reference quirk)
    else { V3 ns = b[0] * sc.normal(m, v[0]) + b[1] * sc.normal(m, v[1]) + b[2] * sc.normal(m, v[2]); n = face_forward(n, ns); }
    V3 p_abs_sum = vabs(b[0] * p0) + vabs(b[1] * p1) + vabs(b[2] * p2);
    V3 p_error = gamma_n(6) * p_abs_sum;
    ss->pi = p3fi_from_value_and_error(p, p_error); ss->n = n; ss->pdf = 1.0f / tri_area(p0, p1, p2);
}
// Triangle::sample_with_context triangle.rs:595-694
inline bool tri_sample_with_context(const Scene& sc, uint32_t mesh_id, uint32_t tri, const LightSampleContext& ctx, V2 u, ShapeSample* ss) {
    const SgMesh& m = sc.d->meshes[mesh_id];
    uint32_t v[3]; sc.tri_indices(mesh_id, tri, v);
    V3 p0 = sc.vertex(m, v[0]), p1 = sc.vertex(m, v[1]), p2 = sc.vertex(m, v[2]);
    V3 cp = ctx.p();
    Float solid_angle = tri_solid_angle(p0, p1, p2, cp);
    if (solid_angle < 3e-4f || solid_angle > 6.22f) {
        tri_sample_area(sc, mesh_id, tri, u, ss);
        V3 wi = p3fi_mid(ss->pi) - cp;
        if (length_squared(wi) == 0.0f) return false;
        wi = normalize(wi);
        ss->pdf /= abs_dot(ss->n, -wi) / distance_squared(cp, p3fi_mid(ss->pi));
        if (std::isinf(ss->pdf)) return false;
        return true;
    }
    Float pdf = 1.0f;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        V3 wi[3] = {normalize(p0 - cp), normalize(p1 - cp), normalize(p2 - cp)};
        Float w[4] = {fmax_(0.01f, abs_dot(ctx.ns, wi[1])), fmax_(0.01f, abs_dot(ctx.ns, wi[1])),
                      fmax_(0.01f, abs_dot(ctx.ns, wi[0])), fmax_(0.01f, abs_dot(ctx.ns, wi[2]))};
        V2 uw = sample_bilinear(u, w);      // the warped u is only used for the pdf (:642-644): reference quirk
        pdf = bilinear_pdf(uw, w);
    }
    V3 tv[3] = {p0, p1, p2};
    Float b[3], tri_pdf;
    sample_spherical_triangle(tv, cp, u, b, &tri_pdf);
    if (tri_pdf == 0.0f) return false;
    pdf = pdf * tri_pdf;
    V3 p_abs_sum = vabs(b[0] * p0) + vabs(b[1] * p1) + vabs((1.0f - b[0] - b[1]) * p2);
    V3 p_error = gamma_n(6) * p_abs_sum;
    V3 p = b[0] * p0 + b[1] * p1 + b[2] * p2;
    V3 n = normalize(cross(p1 - p0, p2 - p0));
    if (m.flags & SG_MESH_HAS_N) { V3 ns = b[0] * sc.normal(m, v[0]) + b[1] * sc.normal(m, v[1]) + b[2] * sc.normal(m, v[2]); n = face_forward(n, ns); }
    else if (((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0)) n = n * -1.0f;
    ss->pi = p3fi_from_value_and_error(p, p_error); ss->n = n; ss->pdf = pdf;
    return true;
}
// Triangle::pdf_with_context triangle.rs:696-745
inline Float tri_pdf_with_context(const Scene& sc, uint32_t mesh_id, uint32_t tri, const LightSampleContext& ctx, V3 wi) {
    V3 p0, p1, p2; sc.tri_points(mesh_id, tri, &p0, &p1, &p2);
    V3 cp = ctx.p();
    Float solid_angle = tri_solid_angle(p0, p1, p2, cp);
    if (solid_angle < 3e-4f || solid_angle > 6.22f) {
        Ray ray; ray.o = offset_ray_origin(ctx.pi, ctx.n, wi); ray.d = wi;       // ShapeSampleContext::spawn_ray shape.rs:276-283
        TriHit th;
        if (!intersect_triangle(ray, F_INF, p0, p1, p2, &th)) return 0.0f;
        SurfaceInteraction isect = interaction_from_intersection(sc, mesh_id, tri, th, -wi);
        Float pdf = (1.0f / tri_area(p0, p1, p2)) / (abs_dot(isect.n, -wi) / distance_squared(cp, isect.p()));
        if (std::isinf(pdf)) return 0.0f;
        return pdf;
    }
    Float pdf = 1.0f / solid_angle;
    if (!(ctx.ns.x == 0.0f && ctx.ns.y == 0.0f && ctx.ns.z == 0.0f)) {
        V3 tv[3] = {p0, p1, p2};
        V2 u = invert_spherical_triangle_sample(tv, cp, wi);
        V3 wv[3] = {normalize(p0 - cp), normalize(p1 - cp), normalize(p2 - cp)};
        Float w[4] = {fmax_(0.01f, abs_dot(ctx.ns, wv[1])), fmax_(0.01f, abs_dot(ctx.ns, wv[1])),
                      fmax_(0.01f, abs_dot(ctx.ns, wv[0])), fmax_(0.01f, abs_dot(ctx.ns, wv[2]))};
        pdf *= bilinear_pdf(u, w);
    }
    return pdf;
}
// ---- Sphere as an emitter: sphere.rs:299-457 ----
inline Float sphere_area(const SgSphere& S) { return S.phi_max * S.radius * (S.z_max - S.z_min); }            // :295-297
inline V3 sphere_center(const SgSphere& S) { return v3(S.render_from_object[3], S.render_from_object[7], S.render_from_object[11]); }   // apply(Point3f::ZERO)
// Sphere::sample :299-333 (area sampling, used when the reference point is inside the sphere)
inline void sphere_sample_area(const SgSphere& S, V2 u, ShapeSample* ss) {
    const Float z = 1.0f - 2.0f * u.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * PI_F * u.y;             // sample_uniform_sphere sampling.rs:280-289
    V3 p_obj = v3(0, 0, 0) + v3(r * std::cos(phi), r * std::sin(phi), z) * S.radius;
    const Float sc_ = S.radius / length(p_obj);
    p_obj = v3(p_obj.x * sc_, p_obj.y * sc_, p_obj.z * sc_);
    const V3 p_err = gamma_n(5) * vabs(p_obj);
    const float* Mi = S.object_from_render; const float* M = S.render_from_object;
    V3 n = normalize(v3(Mi[0] * p_obj.x + Mi[4] * p_obj.y + Mi[8] * p_obj.z, Mi[1] * p_obj.x + Mi[5] * p_obj.y + Mi[9] * p_obj.z,
                        Mi[2] * p_obj.x + Mi[6] * p_obj.y + Mi[10] * p_obj.z));                          // apply(Normal3f) transform.rs:377-383,779-786
    if (S.flags & SG_MESH_REVERSE_ORIENTATION) n = n * -1.0f;
    // render_from_object.apply(Point3fi::from_value_and_error(p_obj, p_err)) : transform.rs:385-457
    const P3fi pin = p3fi_from_value_and_error(p_obj, p_err);
    const V3 pm = p3fi_mid(pin), pe = p3fi_error(pin);
    const bool exact = p3fi_is_exact(pin);
    Float qa[3], ea[3];
    for (int r_ = 0; r_ < 3; ++r_) {
        qa[r_] = (M[4 * r_] * pm.x + M[4 * r_ + 1] * pm.y) + (M[4 * r_ + 2] * pm.z + M[4 * r_ + 3]);
        const Float a = gamma_n(3) * (std::fabs(M[4 * r_] * pm.x) + std::fabs(M[4 * r_ + 1] * pm.y) + std::fabs(M[4 * r_ + 2] * pm.z) + std::fabs(M[4 * r_ + 3]));
        ea[r_] = exact ? a : (gamma_n(3) + 1.0f) * (std::fabs(M[4 * r_]) * pe.x + std::fabs(M[4 * r_ + 1]) * pe.y + std::fabs(M[4 * r_ + 2]) * pe.z) + a;
    }
    ss->pi = p3fi_from_value_and_error(v3(qa[0], qa[1], qa[2]), v3(ea[0], ea[1], ea[2])); ss->n = n; ss->pdf = 1.0f / sphere_area(S);
}
// Sphere::sample_with_context :339-420
inline bool sphere_sample_with_context(const SgSphere& S, const LightSampleContext& ctx, V2 u, ShapeSample* ss) {
    const V3 pc = sphere_center(S), cp = ctx.p();
    const V3 p_origin = offset_ray_origin(ctx.pi, ctx.n, pc - cp);                                        // offset_ray_origin_pt shape.rs:275-277
    if (distance_squared(p_origin, pc) <= sqr(S.radius)) {
        sphere_sample_area(S, u, ss);
        V3 wi = p3fi_mid(ss->pi) - cp;
        if (length_squared(wi) == 0.0f) return false;
        wi = normalize(wi);
        ss->pdf /= abs_dot(ss->n, -wi) / distance_squared(cp, p3fi_mid(ss->pi));
        if (std::isinf(ss->pdf)) return false;
        return true;
    }
    const Float sin_theta_max = S.radius / std::sqrt(distance_squared(cp, pc));
    const Float sin2_theta_max = sqr(sin_theta_max);
    const Float cos_theta_max = safe_sqrt(1.0f - sin2_theta_max);
    Float one_minus_cos_theta_max = 1.0f - cos_theta_max;
    Float cos_theta = (cos_theta_max - 1.0f) * u.x + 1.0f;
    Float sin2_theta = 1.0f - sqr(cos_theta);
    if (sin2_theta_max < 0.00068523f) {
        sin2_theta = sin2_theta_max * u.x;
        cos_theta = std::sqrt(1.0f - sin2_theta);
        one_minus_cos_theta_max = sin2_theta_max / 2.0f;
    }
    const Float cos_alpha = sin2_theta / sin_theta_max + cos_theta * safe_sqrt(1.0f - sin2_theta / sqr(sin_theta_max));
    const Float sin_alpha = safe_sqrt(1.0f - sqr(cos_alpha));
    const Float phi = u.y * 2.0f * PI_F;
    const V3 w = v3(clampf(sin_alpha, -1.0f, 1.0f) * std::cos(phi), clampf(sin_alpha, -1.0f, 1.0f) * std::sin(phi), clampf(cos_alpha, -1.0f, 1.0f));   // vector.rs:1024-1032
    const V3 fz = normalize(pc - cp); V3 fx, fy; coordinate_system(fz, &fx, &fy);                        // Frame::from_z frame.rs:24-27
    const V3 mw = -w;
    V3 n = mw.x * fx + mw.y * fy + mw.z * fz;                                                            // from_local_v frame.rs:51-53
    if (S.flags & SG_MESH_REVERSE_ORIENTATION) n = n * -1.0f;                                            // sic: flipped BEFORE the point is placed (:388-389)
    const V3 p = pc + v3(n.x, n.y, n.z) * S.radius;
    ss->pi = p3fi_from_value_and_error(p, gamma_n(5) * vabs(p)); ss->n = n;
    ss->pdf = 1.0f / (2.0f * PI_F * one_minus_cos_theta_max);
    return true;
}
// Sphere::pdf_with_context :422-456
inline Float sphere_pdf_with_context(const SgSceneDesc* D, const SgSphere& S, const LightSampleContext& ctx, V3 wi) {
    const V3 pc = sphere_center(S), cp = ctx.p();
    const V3 p_origin = offset_ray_origin(ctx.pi, ctx.n, pc - cp);
    if (distance_squared(p_origin, pc) <= S.radius * S.radius) {
        Ray ray; ray.o = offset_ray_origin(ctx.pi, ctx.n, wi); ray.d = wi;
        QuadricHit q;
        if (!sphere_basic_intersect(S, ray, F_INF, &q)) return 0.0f;
        const SurfaceInteraction isect = sphere_interaction(D, S, q.p_obj, q.phi, -wi);
        const Float pdf = (1.0f / sphere_area(S)) / abs_dot(isect.n, -wi) / distance_squared(cp, isect.p());   // sic: (a / b) / c, :436-438
        if (std::isinf(pdf)) return 0.0f;
        return pdf;
    }
    const Float sin2_theta_max = S.radius * S.radius / distance_squared(cp, pc);
    const Float cos_theta_max = safe_sqrt(1.0f - sin2_theta_max);
    Float one_minus_cos_theta_max = 1.0f - cos_theta_max;
    if (sin2_theta_max < 0.00068523f) one_minus_cos_theta_max = sin2_theta_max / 2.0f;
    return 1.0f / (2.90f * PI_F * one_minus_cos_theta_max);                                              // sic: 2.90, :455
}

}  // namespace orc
#include "orc_patch_light.h"
namespace orc {

// sample_uniform_sphere sampling.rs:280-289
inline V3 sample_uniform_sphere(V2 u) {
    const Float z = 1.0f - 2.0f * u.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * PI_F * u.y;
    return v3(r * std::cos(phi), r * std::sin(phi), z);
}
// Light::sample_li: light.rs:632-661 (area), :461-484 (point), :740-766 (uniform infinite), :847-880 (image infinite).
// `allow_incomplete` is true from PathIntegrator::sample_ld (integrator.rs:933) and false from SimplePathIntegrator (:652-656).
inline bool light_sample_li(const Scene& sc, const SgLight& lt, const LightSampleContext& ctx, V2 u, const Wavelengths& lambda, LightLiSample* ls,
                            bool allow_incomplete = true) {
    const SgSceneDesc* D = sc.d;
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA || lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE || lt.kind == SG_LIGHT_DIFFUSE_AREA_PATCH) {
        ShapeSample ss;
        if (lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE) { if (!sphere_sample_with_context(D->spheres[lt.tri], ctx, u, &ss)) return false; }
        else if (lt.kind == SG_LIGHT_DIFFUSE_AREA_PATCH) { if (!patch_sample_with_context(sc, lt.mesh, lt.tri, ctx, u, &ss)) return false; }
        else if (!tri_sample_with_context(sc, lt.mesh, lt.tri, ctx, u, &ss)) return false;
        V3 sp = p3fi_mid(ss.pi);
        if (ss.pdf == 0.0f || length_squared(sp - ctx.p()) == 0.0f) return false;
        V3 wi = normalize(sp - ctx.p());
        Spec le = light_l(D, lt, ss.n, -wi, lambda);
        if (spec_is_zero(le)) return false;
        ls->l = le; ls->wi = wi; ls->pdf = ss.pdf; ls->p_light = ss.pi; ls->n_light = ss.n;
        return true;
    } else if (lt.kind == SG_LIGHT_POINT) {
        V3 p = v3(lt.pos[0], lt.pos[1], lt.pos[2]);
        V3 wi = normalize(p - ctx.p());
        ls->l = lt.scale * spectrum_sample(D, lt.spectrum, lambda) / distance_squared(p, ctx.p());
        ls->wi = wi; ls->pdf = 1.0f; ls->p_light = p3fi_exact(p); ls->n_light = v3(0, 0, 0);
        return true;
    } else if (lt.kind == SG_LIGHT_IMAGE_INFINITE) {                                                  // light.rs:847-880
        const SgEnvMap& E = D->env_maps[lt.tri];
        Float map_pdf;
        const V2 uv = pc2d_sample(D, allow_incomplete ? E.compensated : E.distribution, u, &map_pdf);
        if (map_pdf == 0.0f) return false;
        const V3 wi = xform_vector3(E.render_from_light, equal_area_square_to_sphere(uv));
        ls->l = env_image_le(D, lt, uv, lambda); ls->wi = wi; ls->pdf = map_pdf / (4.0f * PI_F);
        ls->p_light = p3fi_exact(ctx.p() + wi * (2.0f * lt.scene_radius)); ls->n_light = v3(0, 0, 0);
        return true;
    } else if (lt.kind == SG_LIGHT_UNIFORM_INFINITE && !allow_incomplete) {                            // light.rs:750-765
        const V3 wi = sample_uniform_sphere(u);
        ls->l = lt.scale * spectrum_sample(D, lt.spectrum, lambda); ls->wi = wi; ls->pdf = INV_4PI;   // uniform_hemisphere_pdf() == 1/(4 pi), sampling.rs:306-308
        ls->p_light = p3fi_exact(ctx.p() + wi * (2.0f * lt.scene_radius)); ls->n_light = v3(0, 0, 0);
        return true;
    }
    return false;   // UniformInfiniteLight::sample_li returns None when allow_incomplete_pdf (light.rs:748-750)
}
inline Float light_pdf_li(const Scene& sc, const SgLight& lt, const LightSampleContext& ctx, V3 wi, bool allow_incomplete = true) {
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA) return tri_pdf_with_context(sc, lt.mesh, lt.tri, ctx, wi);     // light.rs:663-666
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA_SPHERE) return sphere_pdf_with_context(sc.d, sc.d->spheres[lt.tri], ctx, wi);
    if (lt.kind == SG_LIGHT_DIFFUSE_AREA_PATCH) return patch_pdf_with_context(sc, lt.mesh, lt.tri, ctx, wi);
    if (lt.kind == SG_LIGHT_IMAGE_INFINITE) return env_pdf_li(sc.d, lt, wi, allow_incomplete);           // :882-892
    if (lt.kind == SG_LIGHT_UNIFORM_INFINITE && !allow_incomplete) return INV_4PI;                       // :768-780
    return 0.0f;                                                                                       // :486-494, :770-781
}
// Light::le for the infinite lights: light.rs:792-794, :907-911
inline bool light_is_infinite(const SgLight& lt) { return lt.kind == SG_LIGHT_UNIFORM_INFINITE || lt.kind == SG_LIGHT_IMAGE_INFINITE; }
inline Spec light_le(const SgSceneDesc* D, const SgLight& lt, V3 ray_d, const Wavelengths& lambda) {
    if (lt.kind == SG_LIGHT_IMAGE_INFINITE) return env_le(D, lt, ray_d, lambda);
    return lt.scale * spectrum_sample(D, lt.spectrum, lambda);
}

// ---- camera ------------------------------------------------------------------------
inline V3 xform_point(const float m[16], V3 p) {           // apply_point_helper transform.rs:753-767
    Float xp = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    Float yp = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    Float zp = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    Float wp = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (wp == 1.0f) return v3(xp, yp, zp);
    return v3(xp, yp, zp) / wp;
}
inline V3 xform_vector(const float m[16], V3 v) {          // apply_vector_helper transform.rs:770-776
    return v3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z);
}
// Transform::apply(Point3fi) for an EXACT input point, transform.rs:385-457 (exact branch + wp != 1 division
// is never hit for render_from_camera, whose last row is 0 0 0 1).
inline P3fi xform_point_exact_fi(const float m[16], V3 p) {
    Float x = p.x, y = p.y, z = p.z;
    Float xp = (m[0] * x + m[1] * y) + (m[2] * z + m[3]);
    Float yp = (m[4] * x + m[5] * y) + (m[6] * z + m[7]);
    Float zp = (m[8] * x + m[9] * y) + (m[10] * z + m[11]);
    V3 err = v3(gamma_n(3) * (std::fabs(m[0] * x) + std::fabs(m[1] * y) + std::fabs(m[2] * z) + std::fabs(m[3])),
                gamma_n(3) * (std::fabs(m[4] * x) + std::fabs(m[5] * y) + std::fabs(m[6] * z) + std::fabs(m[7])),
                gamma_n(3) * (std::fabs(m[8] * x) + std::fabs(m[9] * y) + std::fabs(m[10] * z) + std::fabs(m[11])));
    return p3fi_from_value_and_error(v3(xp, yp, zp), err);
}
// Transform::apply_ray transform.rs:515-532 (t_max = None): `self.apply(val.o)` is the Point3f overload (apply_point_helper
// :753-767, summed left to right) and `.into()` makes a zero-width Point3fi -> o.error() == 0, dt == 0; the interval addition
// o + (d * dt) and the midpoint of `o.into()` remain.
inline Ray xform_ray(const float m[16], Ray r) {
    const Float x = r.o.x, y = r.o.y, z = r.o.z;
    V3 pp = v3(((m[0] * x + m[1] * y) + m[2] * z) + m[3], ((m[4] * x + m[5] * y) + m[6] * z) + m[7], ((m[8] * x + m[9] * y) + m[10] * z) + m[11]);
    P3fi o = p3fi_from_value_and_error(pp, v3(0.0f, 0.0f, 0.0f));
    V3 d = xform_vector(m, r.d);
    Float ls = length_squared(d);
    if (ls > 0.0f) {
        Float dt = dot(vabs(d), p3fi_error(o)) / ls;
        V3 off = d * dt;
        // Point3fi + Vector3fi(exact): Interval + Interval = (add_round_down, add_round_up) interval.rs:353-356
        o.lo = v3(next_float_down(o.lo.x + off.x), next_float_down(o.lo.y + off.y), next_float_down(o.lo.z + off.z));
        o.hi = v3(next_float_up(o.hi.x + off.x), next_float_up(o.hi.y + off.y), next_float_up(o.hi.z + off.z));
    }
    Ray out; out.o = p3fi_mid(o); out.d = d; return out;
}
struct CameraSample { V2 p_film, p_lens; Float time; Float filter_weight; };
// PerspectiveCamera::generate_ray_differential camera.rs:1003-1079; `aux` (may be null) receives the auxiliary rays in
// render space (Transform::apply_ray(RayDifferential) transform.rs:534-556: plain point / vector transforms)
inline Ray camera_generate_ray(const SgCamera& cam, const CameraSample& cs, AuxRays* aux = nullptr) {
    V3 p_film = v3(cs.p_film.x, cs.p_film.y, 0.0f);
    V3 p_camera = xform_point(cam.camera_from_raster, p_film);
    if (cam.kind == SG_CAMERA_ORTHOGRAPHIC) {                // OrthographicCamera::generate_ray_differential camera.rs:760-784
        Ray ro; ro.o = p_camera; ro.d = v3(0, 0, 1);         // sic: returned in CAMERA space (no render_from_camera)
        if (aux) {
            aux->has = true;
            aux->rxo = ro.o + v3(cam.dx_camera[0], cam.dx_camera[1], cam.dx_camera[2]); aux->rxd = ro.d;
            aux->ryo = ro.o + v3(cam.dy_camera[0], cam.dy_camera[1], cam.dy_camera[2]); aux->ryd = ro.d;
        }
        return ro;
    }
    Ray r; r.o = v3(0, 0, 0); r.d = normalize(p_camera);
    if (cam.lens_radius > 0.0f) {
        V2 pl = sample_uniform_disk_concentric(cs.p_lens);
        pl.x = cam.lens_radius * pl.x; pl.y = cam.lens_radius * pl.y;
        Float ft = cam.focal_distance / r.d.z;
        V3 p_focus = r.o + r.d * ft;
        r.o = v3(pl.x, pl.y, 0.0f);
        r.d = normalize(p_focus - r.o);
    }
    if (aux) {
        V3 dxc = v3(cam.dx_camera[0], cam.dx_camera[1], cam.dx_camera[2]), dyc = v3(cam.dy_camera[0], cam.dy_camera[1], cam.dy_camera[2]);
        V3 rxo, rxd, ryo, ryd;
        if (cam.lens_radius > 0.0f) {
            V2 pl = sample_uniform_disk_concentric(cs.p_lens);
            pl.x = cam.lens_radius * pl.x; pl.y = cam.lens_radius * pl.y;
            V3 dx = normalize(p_camera + dxc);
            Float ft = cam.focal_distance / dx.z;
            V3 pf = v3(0, 0, 0) + ft * dx;
            rxo = v3(pl.x, pl.y, 0.0f); rxd = normalize(pf - rxo);
            V3 dy = normalize(p_camera + dyc);
            ft = cam.focal_distance / dy.z;
            pf = v3(0, 0, 0) + ft * dy;
            ryo = v3(pl.x, pl.y, 0.0f); ryd = normalize(pf - ryo);
        } else {
            rxo = r.o; ryo = r.o;
            rxd = normalize(p_camera + dxc); ryd = normalize(p_camera + dyc);
        }
        aux->has = true;
        aux->rxo = xform_point(cam.render_from_camera, rxo); aux->rxd = xform_vector(cam.render_from_camera, rxd);
        aux->ryo = xform_point(cam.render_from_camera, ryo); aux->ryd = xform_vector(cam.render_from_camera, ryd);
    }
    return xform_ray(cam.render_from_camera, r);
}

// ---- film --------------------------------------------------------------------------
// PixelSensor::to_sensor_rgb film.rs:907-914 + RgbFilm::add_sample :548-574
inline void film_add_sample(const SgSceneDesc* D, SgFilmPixel* px, Spec L, const Wavelengths& lambda, Float weight) {
    Spec l;
    for (int i = 0; i < 4; ++i) l.v[i] = lambda.pdf[i] != 0.0f ? L.v[i] / lambda.pdf[i] : 0.0f;     // safe_div
    const SgFilm& F = D->film;
    int ids[3] = {F.r_bar, F.g_bar, F.b_bar};
    Float rgb[3];
    for (int c = 0; c < 3; ++c) {
        Spec s = spectrum_sample(D, ids[c], lambda) * l;
        Float sum = 0.0f;                                     // Iterator::sum::<f32>() folds from 0.0
        for (int i = 0; i < 4; ++i) sum = sum + s.v[i];
        rgb[c] = (sum / 4.0f) * F.imaging_ratio;
    }
    Float m = fmax_(fmax_(rgb[0], rgb[1]), rgb[2]);
    if (m > F.max_component_value) for (int c = 0; c < 3; ++c) rgb[c] = rgb[c] * F.max_component_value / m;
    for (int c = 0; c < 3; ++c) px->rgb_sum[c] += (double)(weight * rgb[c]);
    px->weight_sum += (double)weight;
}

}  // namespace orc
