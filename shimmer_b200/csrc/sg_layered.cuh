// CoatedDiffuseBxDF = LayeredBxDF<DielectricBxDF, DiffuseBxDF, TWO_SIDED = true> on the device
// (bxdf.rs:269-326, 883-1620; HGPhaseFunction media.rs:8-33; henyey_greenstein /
// sample_henyey_greenstein scattering.rs:231-260; sample_exponential sampling.rs:789-792).
//
// The reference draws the random walk's numbers from SmallRng::from_entropy() inside every
// f / sample_f / pdf call (bxdf.rs:1011,1270,1423): non-deterministic by construction.  Here that
// private generator is seeded from the path stream's current state and the call site
// (layer_seed in sg_wavefront.cuh), identically to the CPU oracle, so coated paths stay
// sample-for-sample comparable.
// Included from the middle of sg_shading.cuh (needs TR, BSDFSample, refract3, fresnel_dielectric).
#pragma once

namespace sg {

enum { SF_REFLECTION = 1, SF_TRANSMISSION = 2, SF_ALL = 3 };     // BxDFReflTransFlags bxdf.rs:1763-1771

// ---- DielectricBxDF with explicit TransportMode / sample flags (bxdf.rs:532-795) ----
SGD int dielectric_flags(float eta, const TR& mf) {
    int f = (eta == 1.0f) ? BX_TRANSMISSION : (BX_REFLECTION | BX_TRANSMISSION);
    return f | (mf.smooth() ? BX_SPECULAR : BX_GLOSSY);
}
SGD Spec dielectric_f(float eta, const TR& mf, float3 wo, float3 wi, bool radiance) {
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
    if (radiance) ft /= sqr(etap);
    return spec1(ft);
}
SGD float dielectric_pdf(float eta, const TR& mf, float3 wo, float3 wi, int sflags) {
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
    if (!(sflags & SF_REFLECTION)) pr = 0.0f;
    if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
    if (pr == 0.0f && pt == 0.0f) return 0.0f;
    if (refl) return mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm)) * pr / (pr + pt);
    float denom = sqr(dot3(wi, wm) + dot3(wo, wm) / etap);
    float dwm_dwi = absdot3(wi, wm) / denom;
    return mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
}
SGD bool dielectric_sample(float eta, const TR& mf, float3 wo, float uc, float2 u, bool radiance, int sflags, BSDFSample& bs) {
    bs.eta = 1.0f;
    if (eta == 1.0f || mf.smooth()) {
        float R = fresnel_dielectric(wo.z, eta), T = 1.0f - R;
        float pr = R, pt = T;
        if (!(sflags & SF_REFLECTION)) pr = 0.0f;
        if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
        if (pr == 0.0f && pt == 0.0f) return false;
        if (uc < pr / (pr + pt)) {
            float3 wi = f3(-wo.x, -wo.y, wo.z);
            bs.f = spec1(R / fabsf(wi.z)); bs.wi = wi; bs.pdf = pr / (pr + pt); bs.flags = BX_SPECULAR | BX_REFLECTION;
            return true;
        }
        float3 wi; float etap;
        if (!refract3(wo, f3(0.0f, 0.0f, 1.0f), eta, wi, etap)) return false;
        float ft = T / fabsf(wi.z);
        if (radiance) ft /= sqr(etap);
        bs.f = spec1(ft); bs.wi = wi; bs.pdf = pt / (pr + pt); bs.flags = BX_SPECULAR | BX_TRANSMISSION; bs.eta = etap;
        return true;
    }
    float3 wm = mf.sample_wm(wo, u);
    float R = fresnel_dielectric(dot3(wo, wm), eta), T = 1.0f - R;
    float pr = R, pt = T;
    if (!(sflags & SF_REFLECTION)) pr = 0.0f;
    if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
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
    if (radiance) ft /= sqr(etap);
    bs.f = spec1(ft); bs.wi = wi; bs.pdf = pdf; bs.flags = BX_GLOSSY | BX_TRANSMISSION; bs.eta = etap;
    return true;
}
// ---- DiffuseBxDF with sample flags (bxdf.rs:195-267) ----
SGD int diffuse_flags(Spec r) { return spec_zero(r) ? 0 : (BX_DIFFUSE | BX_REFLECTION); }
SGD Spec diffuse_f(Spec r, float3 wo, float3 wi) { return same_hemisphere(wo, wi) ? r * kInvPi : spec1(0.0f); }
SGD float diffuse_pdf(float3 wo, float3 wi, int sflags) {
    if (!(sflags & SF_REFLECTION) || !same_hemisphere(wo, wi)) return 0.0f;
    return fabsf(wi.z) * kInvPi;
}
SGD bool diffuse_sample(Spec r, float3 wo, float2 u, int sflags, BSDFSample& bs) {
    if (!(sflags & SF_REFLECTION)) return false;
    float3 wi = sample_cosine_hemisphere(u);
    if (wo.z < 0.0f) wi.z *= -1.0f;
    bs.f = r * kInvPi; bs.wi = wi; bs.pdf = fabsf(wi.z) * kInvPi; bs.flags = BX_DIFFUSE | BX_REFLECTION; bs.eta = 1.0f;
    return true;
}

// ---- ConductorBxDF with sample flags (bxdf.rs:328-458): the bottom interface of CoatedConductorBxDF ----
SGD int conductor_flags(const TR& mf) { return (mf.smooth() ? BX_SPECULAR : BX_GLOSSY) | BX_REFLECTION; }
SGD Spec conductor_f(const TR& mf, Spec eta, Spec k, float3 wo, float3 wi) {                    // :349-376
    if (!same_hemisphere(wo, wi)) return spec1(0.0f);
    if (mf.smooth()) return spec1(0.0f);
    const float cto = fabsf(wo.z), cti = fabsf(wi.z);
    if (cti == 0.0f || cto == 0.0f) return spec1(0.0f);
    float3 wm = wi + wo;
    if (len2(wm) == 0.0f) return spec1(0.0f);
    wm = normalize3(wm);
    const Spec F = fresnel_complex_spectral(absdot3(wo, wm), eta, k);
    return mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
}
SGD float conductor_pdf(const TR& mf, float3 wo, float3 wi, int sflags) {                       // :424-445
    if (!(sflags & SF_REFLECTION) || !same_hemisphere(wo, wi) || mf.smooth()) return 0.0f;
    float3 wm = wo + wi;
    if (len2(wm) == 0.0f) return 0.0f;
    wm = faceforward3(normalize3(wm), f3(0.0f, 0.0f, 1.0f));
    return mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm));
}
SGD bool conductor_sample(const TR& mf, Spec eta, Spec k, float3 wo, float2 u, int sflags, BSDFSample& bs) {   // :378-422
    bs.eta = 1.0f;
    if (!(sflags & SF_REFLECTION)) return false;
    if (mf.smooth()) {
        const float3 wi = f3(-wo.x, -wo.y, wo.z);
        bs.f = fresnel_complex_spectral(fabsf(wi.z), eta, k) / fabsf(wi.z);
        bs.wi = wi; bs.pdf = 1.0f; bs.flags = BX_SPECULAR | BX_REFLECTION;
        return true;
    }
    if (wo.z == 0.0f) return false;
    const float3 wm = mf.sample_wm(wo, u);
    const float3 wi = reflect3(wo, wm);
    if (!same_hemisphere(wo, wi)) return false;
    const float pdf = mf.pdf(wo, wm) / (4.0f * absdot3(wo, wm));
    const float cto = fabsf(wo.z), cti = fabsf(wi.z);
    if (cti == 0.0f || cto == 0.0f) return false;
    const Spec F = fresnel_complex_spectral(absdot3(wo, wm), eta, k);
    bs.f = mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
    bs.wi = wi; bs.pdf = pdf; bs.flags = BX_GLOSSY | BX_REFLECTION;
    return true;
}

static constexpr float kInv4Pi = 0.07957747154594766788f;
SGD float henyey_greenstein(float cos_t, float g) {                       // scattering.rs:231-236
    g = clampf(g, -0.99f, 0.99f);
    float denom = 1.0f + sqr(g) + 2.0f * g * cos_t;
    return kInv4Pi * (1.0f - sqr(g)) / (denom * safe_sqrt(denom));
}
SGD float sample_henyey_greenstein(float3 wo, float g, float2 u, float3& wi) {   // scattering.rs:239-260
    g = clampf(g, -0.99f, 0.99f);
    float cos_t;
    if (fabsf(g) < 1e-3f) cos_t = 1.0f - 2.0f * u.x;
    else cos_t = -1.0f / (2.0f * g) * (1.0f + sqr(g) - sqr((1.0f - sqr(g)) / (1.0f + g - 2.0f * g * u.x)));
    float sin_t = safe_sqrt(1.0f - sqr(cos_t));
    float phi = 2.0f * kPi * u.y;
    float3 fx, fy; coord_system(wo, fx, fy);                              // Frame::from_z frame.rs:24-27
    float3 l = f3(clampf(sin_t, -1.0f, 1.0f) * cosf(phi), clampf(sin_t, -1.0f, 1.0f) * sinf(phi), clampf(cos_t, -1.0f, 1.0f));
    wi = l.x * fx + l.y * fy + l.z * wo;
    return henyey_greenstein(cos_t, g);
}
// sampling.rs:789-792: the reference evaluates the exponential PDF, not its inverse CDF (kept)
SGD float sample_exponential(float x, float a) { return a * expf(-a * x); }

// COND = false: CoatedDiffuseBxDF (bottom = DiffuseBxDF r); COND = true: CoatedConductorBxDF (bottom = ConductorBxDF ce, ck, mfb; bxdf.rs:460-463)
template <bool COND>
struct LayeredT {
    float eta; TR mf;            // top: DielectricBxDF
    Spec r;                      // bottom: DiffuseBxDF
    Spec ce, ck; TR mfb;         // bottom: ConductorBxDF
    Spec albedo; float thickness, g; int max_depth, n_samples;

    SGD int i_flags(bool top) const {
        if (top) return dielectric_flags(eta, mf);
        if constexpr (COND) return conductor_flags(mfb); else return diffuse_flags(r);
    }
    SGD Spec i_f(bool top, float3 wo, float3 wi, bool radiance) const {
        if (top) return dielectric_f(eta, mf, wo, wi, radiance);
        if constexpr (COND) return conductor_f(mfb, ce, ck, wo, wi); else return diffuse_f(r, wo, wi);
    }
    SGD float i_pdf(bool top, float3 wo, float3 wi, int sf) const {
        if (top) return dielectric_pdf(eta, mf, wo, wi, sf);
        if constexpr (COND) return conductor_pdf(mfb, wo, wi, sf); else return diffuse_pdf(wo, wi, sf);
    }
    SGD bool i_sample(bool top, float3 wo, float uc, float2 u, bool radiance, int sf, BSDFSample& bs) const {
        if (top) return dielectric_sample(eta, mf, wo, uc, u, radiance, sf, bs);
        if constexpr (COND) return conductor_sample(mfb, ce, ck, wo, u, sf, bs); else return diffuse_sample(r, wo, u, sf, bs);
    }
    SGD static float tr(float dz, float3 w) {                             // bxdf.rs:923-931 (`<= Float::MIN` never holds)
        if (fabsf(dz) <= -3.40282347e+38f) return 1.0f;
        return expf(-fabsf(dz / w.z));
    }
    SGD static float r1(Rng& rng) { return fminf(rng.get_1d(), next_down(1.0f)); }
    SGD int flags() const {                                               // bxdf.rs:1586-1614
        int tf = i_flags(true), bf = i_flags(false);
        int fl = BX_REFLECTION;
        if (tf & BX_SPECULAR) fl |= BX_SPECULAR;
        if ((tf & BX_DIFFUSE) || (bf & BX_DIFFUSE) || !spec_zero(albedo)) fl |= BX_DIFFUSE;
        else if ((tf & BX_GLOSSY) || (bf & BX_GLOSSY)) fl |= BX_GLOSSY;
        if ((tf & BX_TRANSMISSION) && (bf & BX_TRANSMISSION)) fl |= BX_TRANSMISSION;
        return fl;
    }

    // bxdf.rs:940-1247, mode = Radiance
    __device__ __noinline__ Spec f(float3 wo, float3 wi, Rng rng) const {
        const bool radiance = true;
        Spec f = spec1(0.0f);
        if (wo.z < 0.0f) { wo = -wo; wi = -wi; }                          // TWO_SIDED
        const bool entered_top = true, enter_top = true;
        const bool exit_is_bottom = same_hemisphere(wo, wi) != entered_top;
        const bool exit_top = !exit_is_bottom, non_exit_top = exit_is_bottom;
        const float exit_z = exit_is_bottom ? 0.0f : thickness;
        if (same_hemisphere(wo, wi)) f = i_f(enter_top, wo, wi, radiance) * (float)n_samples;
        for (int s = 0; s < n_samples; ++s) {
            float uc = r1(rng); float2 uu; uu.x = r1(rng); uu.y = r1(rng);
            BSDFSample wos;
            if (!i_sample(enter_top, wo, uc, uu, radiance, SF_TRANSMISSION, wos)) continue;
            if (spec_zero(wos.f) || wos.pdf == 0.0f || wos.wi.z == 0.0f) continue;
            uc = r1(rng); uu.x = r1(rng); uu.y = r1(rng);
            BSDFSample wis;
            if (!i_sample(exit_top, wi, uc, uu, !radiance, SF_TRANSMISSION, wis)) continue;
            if (spec_zero(wis.f) || wis.pdf == 0.0f || wis.wi.z == 0.0f) continue;
            Spec beta = wos.f * fabsf(wos.wi.z) / wos.pdf;
            float z = entered_top ? thickness : 0.0f;
            float3 w = wos.wi;
            for (int depth = 0; depth < max_depth; ++depth) {
                if (depth > 3 && spec_max(beta) < 0.25f) {
                    float q = fmaxf(0.0f, 1.0f - spec_max(beta));
                    if (r1(rng) < q) break;
                    beta = beta / (1.0f - q);
                }
                if (spec_zero(albedo)) {
                    z = (z == thickness) ? 0.0f : thickness;
                    beta = beta * tr(thickness, w);
                } else {
                    float sigma_t = 1.0f;
                    float dz = sample_exponential(r1(rng), sigma_t / fabsf(w.z));
                    float zp = w.z > 0.0f ? (z + dz) : (z - dz);
                    if (z == zp) continue;
                    if (0.0f < zp && zp < thickness) {
                        float wt = 1.0f;
                        if (!(i_flags(exit_top) & BX_SPECULAR)) wt = power_heuristic(wis.pdf, henyey_greenstein(dot3(-w, -wis.wi), g));
                        f = f + beta * albedo * henyey_greenstein(dot3(-w, -wis.wi), g) * wt * tr(zp - exit_z, wis.wi) * wis.f / wis.pdf;
                        float2 u2; u2.x = r1(rng); u2.y = r1(rng);
                        float3 pwi; float pp = sample_henyey_greenstein(-w, g, u2, pwi);
                        if (pp == 0.0f || pwi.z == 0.0f) continue;
                        beta = beta * (albedo * pp / pp);
                        w = pwi; z = zp;
                        if (((z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f)) && !(i_flags(exit_top) & BX_SPECULAR)) {
                            Spec f_exit = i_f(exit_top, -w, wi, radiance);
                            if (!spec_zero(f_exit)) {
                                float exit_pdf = i_pdf(exit_top, -w, wi, SF_TRANSMISSION);
                                float wt2 = power_heuristic(pp, exit_pdf);
                                f = f + beta * tr(zp - exit_z, pwi) * f_exit * wt2;
                            }
                        }
                        continue;
                    }
                    z = clampf(zp, 0.0f, thickness);
                }
                if (z == exit_z) {
                    float uc2 = r1(rng); float2 u2; u2.x = r1(rng); u2.y = r1(rng);
                    BSDFSample bs;
                    if (!i_sample(exit_top, -w, uc2, u2, radiance, SF_REFLECTION, bs)) break;
                    if (spec_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                    beta = beta * (bs.f * fabsf(bs.wi.z) / bs.pdf);
                    w = bs.wi;
                } else {
                    if (!(i_flags(non_exit_top) & BX_SPECULAR)) {
                        float wt = 1.0f;
                        if (!(i_flags(exit_top) & BX_SPECULAR)) wt = power_heuristic(wis.pdf, i_pdf(non_exit_top, -w, -wis.wi, SF_ALL));
                        f = f + beta * i_f(non_exit_top, -w, -wis.wi, radiance) * fabsf(wis.wi.z) * wt * tr(thickness, wis.wi) * wis.f / wis.pdf;
                    }
                    float uc2 = r1(rng); float2 u2; u2.x = r1(rng); u2.y = r1(rng);
                    BSDFSample bs;
                    if (!i_sample(non_exit_top, -w, uc2, u2, radiance, SF_REFLECTION, bs)) break;
                    if (spec_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                    beta = beta * (bs.f * fabsf(bs.wi.z) / bs.pdf);
                    w = bs.wi;
                    if (!(i_flags(exit_top) & BX_SPECULAR)) {
                        Spec f_exit = i_f(exit_top, -w, wi, radiance);
                        if (!spec_zero(f_exit)) {
                            float wt = 1.0f;
                            if (!(i_flags(non_exit_top) & BX_SPECULAR)) {
                                float exit_pdf = i_pdf(exit_top, -w, wi, SF_TRANSMISSION);
                                wt = power_heuristic(bs.pdf, exit_pdf);
                            }
                            f = f + beta * tr(thickness, bs.wi) * f_exit * wt;
                        }
                    }
                }
            }
        }
        return f / (float)n_samples;
    }

    // bxdf.rs:1249-1402; false = None; proportional = pdf_is_proportional
    __device__ __noinline__ bool sample_f(float3 wo, float uc, float2 u, Rng rng, BSDFSample& out, bool& proportional) const {
        const bool radiance = true;
        bool flip_wi = false;
        if (wo.z < 0.0f) { wo = -wo; flip_wi = true; }
        const bool entered_top = true;
        BSDFSample bs;
        if (!i_sample(entered_top, wo, uc, u, radiance, SF_ALL, bs)) return false;
        if (spec_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) return false;
        if (bs.flags & BX_REFLECTION) {
            if (flip_wi) bs.wi = -bs.wi;
            out = bs; proportional = true;
            return true;
        }
        float3 w = bs.wi;
        bool specular_path = (bs.flags & BX_SPECULAR) != 0;
        Spec f = bs.f * fabsf(bs.wi.z);
        float pdf = bs.pdf;
        float z = entered_top ? thickness : 0.0f;
        for (int depth = 0; depth < max_depth; ++depth) {
            float rr_beta = spec_max(f) / pdf;
            if (depth > 3 && rr_beta < 0.25f) {
                float q = fmaxf(0.0f, 1.0f - rr_beta);
                if (r1(rng) < q) return false;
                pdf *= 1.0f - q;
            }
            if (w.z == 0.0f) return false;
            if (!spec_zero(albedo)) {
                float sigma_t = 1.0f;
                float dz = sample_exponential(r1(rng), sigma_t / fabsf(w.z));
                float zp = w.z > 0.0f ? (z + dz) : (z - dz);
                if (zp == z) return false;
                if (0.0f < zp && zp < thickness) {
                    float2 u2; u2.x = r1(rng); u2.y = r1(rng);
                    float3 pwi; float pp = sample_henyey_greenstein(-w, g, u2, pwi);
                    if (pp == 0.0f || pwi.z == 0.0f) return false;
                    f = f * (albedo * pp);
                    pdf *= pp;
                    specular_path = false;
                    w = pwi; z = zp;
                    continue;
                }
                z = clampf(zp, 0.0f, thickness);
            } else {
                z = (z == thickness) ? 0.0f : thickness;
                f = f * tr(thickness, w);
            }
            const bool iface_top = !(z == 0.0f);
            float uc2 = r1(rng); float2 u2; u2.x = r1(rng); u2.y = r1(rng);
            BSDFSample b2;
            if (!i_sample(iface_top, -w, uc2, u2, radiance, SF_ALL, b2)) return false;
            if (spec_zero(b2.f) || b2.pdf == 0.0f || b2.wi.z == 0.0f) return false;
            f = f * b2.f;
            pdf *= b2.pdf;
            specular_path = specular_path && ((b2.flags & BX_SPECULAR) != 0);
            w = b2.wi;
            if (b2.flags & BX_TRANSMISSION) {
                int fl = same_hemisphere(wo, w) ? BX_REFLECTION : BX_TRANSMISSION;
                fl |= specular_path ? BX_SPECULAR : BX_GLOSSY;
                if (flip_wi) w = -w;
                out.f = f; out.wi = w; out.pdf = pdf; out.flags = fl; out.eta = 1.0f; proportional = true;
                return true;
            }
            f = f * fabsf(b2.wi.z);
        }
        return false;
    }

    // bxdf.rs:1404-1584
    __device__ __noinline__ float pdf(float3 wo, float3 wi, Rng rng) const {
        const bool radiance = true;
        if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
        const bool entered_top = true;
        float pdf_sum = 0.0f;
        if (same_hemisphere(wo, wi)) pdf_sum += (float)n_samples * i_pdf(entered_top, wo, wi, SF_REFLECTION);
        for (int s = 0; s < n_samples; ++s) {
            if (same_hemisphere(wo, wi)) {
                const bool r_top = !entered_top, t_top = entered_top;
                float uc = r1(rng); float2 u; u.x = r1(rng); u.y = r1(rng);
                BSDFSample wos; bool has_wos = i_sample(t_top, wo, uc, u, radiance, SF_TRANSMISSION, wos);
                uc = r1(rng); u.x = r1(rng); u.y = r1(rng);
                BSDFSample wis; bool has_wis = i_sample(t_top, wi, uc, u, !radiance, SF_TRANSMISSION, wis);
                if (has_wos && has_wis && !spec_zero(wos.f) && wos.pdf > 0.0f && !spec_zero(wis.f) && wis.pdf > 0.0f) {
                    if (!(i_flags(t_top) & (BX_DIFFUSE | BX_GLOSSY))) pdf_sum += i_pdf(r_top, -wos.wi, -wis.wi, SF_ALL);
                    else {
                        uc = r1(rng); u.x = r1(rng); u.y = r1(rng);
                        BSDFSample rs;
                        if (i_sample(r_top, -wos.wi, uc, u, radiance, SF_ALL, rs)) {
                            if (!(i_flags(r_top) & (BX_DIFFUSE | BX_GLOSSY))) pdf_sum += i_pdf(t_top, -rs.wi, wi, SF_ALL);
                            else {
                                float r_pdf = i_pdf(r_top, -wos.wi, -wis.wi, SF_ALL);
                                float wt = power_heuristic(wis.pdf, r_pdf);
                                pdf_sum += wt * r_pdf;
                                float t_pdf = i_pdf(t_top, -rs.wi, wi, SF_ALL);
                                wt = power_heuristic(rs.pdf, t_pdf);
                                pdf_sum += wt * t_pdf;
                            }
                        }
                    }
                }
            } else {
                const bool to_top = entered_top, ti_top = !entered_top;
                float uc = r1(rng); float2 u; u.x = r1(rng); u.y = r1(rng);
                BSDFSample wos;
                if (!i_sample(to_top, wo, uc, u, radiance, SF_ALL, wos)) continue;
                if (spec_zero(wos.f) || wos.pdf == 0.0f || wos.wi.z == 0.0f || (wos.flags & BX_REFLECTION)) continue;
                uc = r1(rng); u.x = r1(rng); u.y = r1(rng);
                BSDFSample wis;
                if (!i_sample(ti_top, wi, uc, u, !radiance, SF_ALL, wis)) continue;
                if (spec_zero(wis.f) || wis.pdf == 0.0f || wis.wi.z == 0.0f || (wis.flags & BX_REFLECTION)) continue;
                if (i_flags(to_top) & BX_SPECULAR) pdf_sum += i_pdf(ti_top, -wos.wi, wi, SF_ALL);
                else if (i_flags(ti_top) & BX_SPECULAR) pdf_sum += i_pdf(to_top, wo, -wis.wi, SF_ALL);
                else pdf_sum += (i_pdf(to_top, wo, -wis.wi, SF_ALL) + i_pdf(ti_top, -wos.wi, wi, SF_ALL)) / 2.0f;
            }
        }
        return lerpf(0.9f, 1.0f / (4.0f * kPi), pdf_sum / (float)n_samples);
    }
};

}  // namespace sg
