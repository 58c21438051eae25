// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.h header).  parity status: see orc_math.h.
// LayeredBxDF<DielectricBxDF, DiffuseBxDF, TWO_SIDED = true> == CoatedDiffuseBxDF (bxdf.rs:269-326,
// 883-1620), HGPhaseFunction (media.rs:8-33), henyey_greenstein / sample_henyey_greenstein
// (scattering.rs:231-260), sample_exponential (sampling.rs:789-792).
// Included from the middle of orc_shading.h (needs TR, BSDFSample, the dielectric helpers).
//
// RNG: the reference seeds a fresh SmallRng::from_entropy() inside every f / sample_f / pdf call
// (bxdf.rs:1011,1270,1423) -- non-deterministic by construction.  We seed that generator from the
// path's own stream state and a call-site id (see BSDF::layer_rng), identically on CPU and GPU.
#pragma once

namespace orc {

enum { SF_REFLECTION = 1, SF_TRANSMISSION = 2, SF_ALL = 3 };   // BxDFReflTransFlags bxdf.rs:1763-1771

// ---- DielectricBxDF with explicit TransportMode / sample flags (bxdf.rs:532-795) ----
inline int dielectric_flags(Float eta, const TR& mf) {
    int f = (eta == 1.0f) ? BX_TRANSMISSION : (BX_REFLECTION | BX_TRANSMISSION);
    return f | (mf.effectively_smooth() ? BX_SPECULAR : BX_GLOSSY);
}
inline Spec dielectric_f(Float eta, const TR& mf, V3 wo, V3 wi, bool radiance) {
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
    if (radiance) ft /= sqr(etap);
    return spec_const(ft);
}
inline Float dielectric_pdf(Float eta, const TR& mf, V3 wo, V3 wi, int sflags) {
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
    if (!(sflags & SF_REFLECTION)) pr = 0.0f;
    if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
    if (pr == 0.0f && pt == 0.0f) return 0.0f;
    if (refl) return mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm)) * pr / (pr + pt);
    Float denom = sqr(dot(wi, wm) + dot(wo, wm) / etap);
    Float dwm_dwi = abs_dot(wi, wm) / denom;
    return mf.pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
}
inline bool dielectric_sample(Float eta, const TR& mf, V3 wo, Float uc, V2 u, bool radiance, int sflags, BSDFSample* bs) {
    bs->eta = 1.0f;
    if (eta == 1.0f || mf.effectively_smooth()) {
        Float R = fresnel_dielectric(cos_theta(wo), eta), T = 1.0f - R;
        Float pr = R, pt = T;
        if (!(sflags & SF_REFLECTION)) pr = 0.0f;
        if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
        if (pr == 0.0f && pt == 0.0f) return false;
        if (uc < pr / (pr + pt)) {
            V3 wi = v3(-wo.x, -wo.y, wo.z);
            bs->f = spec_const(R / abs_cos_theta(wi)); bs->wi = wi; bs->pdf = pr / (pr + pt); bs->flags = BX_SPECULAR | BX_REFLECTION;
            return true;
        }
        V3 wi; Float etap;
        if (!refract(wo, v3(0, 0, 1), eta, &wi, &etap)) return false;
        Float ft = T / abs_cos_theta(wi);
        if (radiance) ft /= sqr(etap);
        bs->f = spec_const(ft); bs->wi = wi; bs->pdf = pt / (pr + pt); bs->flags = BX_SPECULAR | BX_TRANSMISSION; bs->eta = etap;
        return true;
    }
    V3 wm = mf.sample_wm(wo, u);
    Float R = fresnel_dielectric(dot(wo, wm), eta), T = 1.0f - R;
    Float pr = R, pt = T;
    if (!(sflags & SF_REFLECTION)) pr = 0.0f;
    if (!(sflags & SF_TRANSMISSION)) pt = 0.0f;
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
    if (radiance) ft /= sqr(etap);
    bs->f = spec_const(ft); bs->wi = wi; bs->pdf = pdf; bs->flags = BX_GLOSSY | BX_TRANSMISSION; bs->eta = etap;
    return true;
}
// ---- DiffuseBxDF with sample flags (bxdf.rs:195-267) ----
inline int diffuse_flags(Spec r) { return spec_is_zero(r) ? BX_UNSET : (BX_DIFFUSE | BX_REFLECTION); }
inline Spec diffuse_f(Spec r, V3 wo, V3 wi) { return same_hemisphere(wo, wi) ? r * INV_PI : spec_const(0.0f); }
inline Float diffuse_pdf(V3 wo, V3 wi, int sflags) {
    if (!(sflags & SF_REFLECTION) || !same_hemisphere(wo, wi)) return 0.0f;
    return abs_cos_theta(wi) * INV_PI;
}
inline bool diffuse_sample(Spec r, V3 wo, V2 u, int sflags, BSDFSample* bs) {
    if (!(sflags & SF_REFLECTION)) return false;
    V3 wi = sample_cosine_hemisphere(u);
    if (wo.z < 0.0f) wi.z *= -1.0f;
    bs->f = r * INV_PI; bs->wi = wi; bs->pdf = abs_cos_theta(wi) * INV_PI; bs->flags = BX_DIFFUSE | BX_REFLECTION; bs->eta = 1.0f;
    return true;
}

// ---- ConductorBxDF with sample flags (bxdf.rs:328-458): the bottom interface of CoatedConductorBxDF ----
inline Spec fresnel_complex_spectral(Float c, Spec eta, Spec k) {                             // scattering.rs:94-105
    Spec F; for (int i = 0; i < 4; ++i) F.v[i] = fresnel_complex(c, cx(eta.v[i], k.v[i]));
    return F;
}
inline int conductor_flags(const TR& mf) { return mf.effectively_smooth() ? (BX_SPECULAR | BX_REFLECTION) : (BX_GLOSSY | BX_REFLECTION); }
inline Spec conductor_f(const TR& mf, Spec eta, Spec k, V3 wo, V3 wi) {                       // :349-376
    if (!same_hemisphere(wo, wi)) return spec_const(0.0f);
    if (mf.effectively_smooth()) return spec_const(0.0f);
    Float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
    if (cti == 0.0f || cto == 0.0f) return spec_const(0.0f);
    V3 wm = wi + wo;
    if (length_squared(wm) == 0.0f) return spec_const(0.0f);
    wm = normalize(wm);
    Spec F = fresnel_complex_spectral(abs_dot(wo, wm), eta, k);
    return mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
}
inline Float conductor_pdf(const TR& mf, V3 wo, V3 wi, int sflags) {                          // :424-445
    if (!(sflags & SF_REFLECTION) || !same_hemisphere(wo, wi) || mf.effectively_smooth()) return 0.0f;
    V3 wm = wo + wi;
    if (length_squared(wm) == 0.0f) return 0.0f;
    wm = face_forward(normalize(wm), v3(0, 0, 1));
    return mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm));
}
inline bool conductor_sample(const TR& mf, Spec eta, Spec k, V3 wo, V2 u, int sflags, BSDFSample* bs) {   // :378-422
    bs->eta = 1.0f;
    if (!(sflags & SF_REFLECTION)) return false;
    if (mf.effectively_smooth()) {
        V3 wi = v3(-wo.x, -wo.y, wo.z);
        bs->f = fresnel_complex_spectral(abs_cos_theta(wi), eta, k) / abs_cos_theta(wi);
        bs->wi = wi; bs->pdf = 1.0f; bs->flags = BX_SPECULAR | BX_REFLECTION;
        return true;
    }
    if (wo.z == 0.0f) return false;
    V3 wm = mf.sample_wm(wo, u);
    V3 wi = reflect(wo, wm);
    if (!same_hemisphere(wo, wi)) return false;
    Float pdf = mf.pdf(wo, wm) / (4.0f * abs_dot(wo, wm));
    Float cto = abs_cos_theta(wo), cti = abs_cos_theta(wi);
    if (cti == 0.0f || cto == 0.0f) return false;
    Spec F = fresnel_complex_spectral(abs_dot(wo, wm), eta, k);
    bs->f = mf.d(wm) * F * mf.g(wo, wi) / (4.0f * cto * cti);
    bs->wi = wi; bs->pdf = pdf; bs->flags = BX_GLOSSY | BX_REFLECTION;
    return true;
}

// scattering.rs:231-236
inline Float henyey_greenstein(Float cos_t, Float g) {
    g = clampf(g, -0.99f, 0.99f);
    Float denom = 1.0f + sqr(g) + 2.0f * g * cos_t;
    return INV_4PI * (1.0f - sqr(g)) / (denom * safe_sqrt(denom));
}
// scattering.rs:239-260 ; spherical_direction vector.rs:1024-1032 ; Frame::from_z frame.rs:24-27
inline Float sample_henyey_greenstein(V3 wo, Float g, V2 u, V3* wi) {
    g = clampf(g, -0.99f, 0.99f);
    Float cos_t;
    if (std::fabs(g) < 1e-3f) cos_t = 1.0f - 2.0f * u.x;
    else cos_t = -1.0f / (2.0f * g) * (1.0f + sqr(g) - sqr((1.0f - sqr(g)) / (1.0f + g - 2.0f * g * u.x)));
    Float sin_t = safe_sqrt(1.0f - sqr(cos_t));
    Float phi = 2.0f * PI_F * u.y;
    V3 fx, fy; coordinate_system(wo, &fx, &fy);
    V3 l = v3(clampf(sin_t, -1.0f, 1.0f) * std::cos(phi), clampf(sin_t, -1.0f, 1.0f) * std::sin(phi), clampf(cos_t, -1.0f, 1.0f));
    *wi = l.x * fx + l.y * fy + l.z * wo;
    return henyey_greenstein(cos_t, g);
}
// sampling.rs:789-792 (the reference evaluates the exponential PDF here, not its inverse CDF: kept)
inline Float sample_exponential(Float x, Float a) { return a * std::exp(-a * x); }

struct Layered {
    Float eta; TR mf;            // top: DielectricBxDF
    Spec r;                      // bottom: DiffuseBxDF (CoatedDiffuse)
    bool cond = false;           // bottom: ConductorBxDF (CoatedConductor, bxdf.rs:460-463) with ce, ck, mfb
    Spec ce, ck; TR mfb;
    Spec albedo; Float thickness, g; int max_depth, n_samples;

    // TopOrBottomBxDF dispatch (bxdf.rs:1622-1700)
    int i_flags(bool top) const { return top ? dielectric_flags(eta, mf) : (cond ? conductor_flags(mfb) : diffuse_flags(r)); }
    Spec i_f(bool top, V3 wo, V3 wi, bool radiance) const {
        return top ? dielectric_f(eta, mf, wo, wi, radiance) : (cond ? conductor_f(mfb, ce, ck, wo, wi) : diffuse_f(r, wo, wi));
    }
    Float i_pdf(bool top, V3 wo, V3 wi, int sf) const { return top ? dielectric_pdf(eta, mf, wo, wi, sf) : (cond ? conductor_pdf(mfb, wo, wi, sf) : diffuse_pdf(wo, wi, sf)); }
    bool i_sample(bool top, V3 wo, Float uc, V2 u, bool radiance, int sf, BSDFSample* bs) const {
        return top ? dielectric_sample(eta, mf, wo, uc, u, radiance, sf, bs) : (cond ? conductor_sample(mfb, ce, ck, wo, u, sf, bs) : diffuse_sample(r, wo, u, sf, bs));
    }
    static Float tr(Float dz, V3 w) {          // bxdf.rs:923-931: `abs(dz) <= Float::MIN` can never hold
        if (std::fabs(dz) <= -3.40282347e+38f) return 1.0f;
        return std::exp(-std::fabs(dz / w.z));
    }
    int flags() const {                        // bxdf.rs:1586-1614
        int tf = i_flags(true), bf = i_flags(false);
        int fl = BX_REFLECTION;
        if (tf & BX_SPECULAR) fl |= BX_SPECULAR;
        if ((tf & BX_DIFFUSE) || (bf & BX_DIFFUSE) || !spec_is_zero(albedo)) fl |= BX_DIFFUSE;
        else if ((tf & BX_GLOSSY) || (bf & BX_GLOSSY)) fl |= BX_GLOSSY;
        if ((tf & BX_TRANSMISSION) && (bf & BX_TRANSMISSION)) fl |= BX_TRANSMISSION;
        return fl;
    }

    // bxdf.rs:940-1247 (mode = Radiance at every call site of the path integrator)
    Spec f(V3 wo, V3 wi, Rng& rng) const {
        const bool radiance = true;
        Spec f = spec_const(0.0f);
        if (wo.z < 0.0f) { wo = -wo; wi = -wi; }                              // TWO_SIDED
        const bool entered_top = true;
        const bool enter_top = true;
        const bool exit_is_bottom = same_hemisphere(wo, wi) ^ entered_top;
        const bool exit_top = !exit_is_bottom, non_exit_top = exit_is_bottom;
        const Float exit_z = exit_is_bottom ? 0.0f : thickness;
        if (same_hemisphere(wo, wi)) f = i_f(enter_top, wo, wi, radiance) * (Float)n_samples;
        auto r1 = [&]() { return fmin_(rng.get_1d(), next_float_down(1.0f)); };
        for (int s = 0; s < n_samples; ++s) {
            Float uc = r1(); V2 uu; uu.x = r1(); uu.y = r1();
            BSDFSample wos;
            if (!i_sample(enter_top, wo, uc, uu, radiance, SF_TRANSMISSION, &wos)) continue;
            if (spec_is_zero(wos.f) || wos.pdf == 0.0f || wos.wi.z == 0.0f) continue;
            uc = r1(); uu.x = r1(); uu.y = r1();
            BSDFSample wis;
            if (!i_sample(exit_top, wi, uc, uu, !radiance, SF_TRANSMISSION, &wis)) continue;
            if (spec_is_zero(wis.f) || wis.pdf == 0.0f || wis.wi.z == 0.0f) continue;
            Spec beta = wos.f * abs_cos_theta(wos.wi) / wos.pdf;
            Float z = entered_top ? thickness : 0.0f;
            V3 w = wos.wi;
            for (int depth = 0; depth < max_depth; ++depth) {
                if (depth > 3 && spec_max(beta) < 0.25f) {
                    Float q = fmax_(0.0f, 1.0f - spec_max(beta));
                    if (r1() < q) break;
                    beta = beta / (1.0f - q);
                }
                if (spec_is_zero(albedo)) {
                    z = (z == thickness) ? 0.0f : thickness;
                    beta = beta * tr(thickness, w);
                } else {
                    Float sigma_t = 1.0f;
                    Float dz = sample_exponential(r1(), sigma_t / std::fabs(w.z));
                    Float zp = w.z > 0.0f ? (z + dz) : (z - dz);
                    if (z == zp) continue;
                    if (0.0f < zp && zp < thickness) {
                        Float wt = 1.0f;
                        if (!(i_flags(exit_top) & BX_SPECULAR)) wt = power_heuristic(wis.pdf, henyey_greenstein(dot(-w, -wis.wi), g));
                        f = f + beta * albedo * henyey_greenstein(dot(-w, -wis.wi), g) * wt * tr(zp - exit_z, wis.wi) * wis.f / wis.pdf;
                        V2 u2; u2.x = r1(); u2.y = r1();
                        V3 pwi; Float pp = sample_henyey_greenstein(-w, g, u2, &pwi);
                        if (pp == 0.0f || pwi.z == 0.0f) continue;
                        beta = beta * (albedo * pp / pp);
                        w = pwi; z = zp;
                        if (((z < exit_z && w.z > 0.0f) || (z > exit_z && w.z < 0.0f)) && !(i_flags(exit_top) & BX_SPECULAR)) {
                            Spec f_exit = i_f(exit_top, -w, wi, radiance);
                            if (!spec_is_zero(f_exit)) {
                                Float exit_pdf = i_pdf(exit_top, -w, wi, SF_TRANSMISSION);
                                Float wt2 = power_heuristic(pp, exit_pdf);
                                f = f + beta * tr(zp - exit_z, pwi) * f_exit * wt2;
                            }
                        }
                        continue;
                    }
                    z = clampf(zp, 0.0f, thickness);
                }
                if (z == exit_z) {
                    Float uc2 = r1(); V2 u2; u2.x = r1(); u2.y = r1();
                    BSDFSample bs;
                    if (!i_sample(exit_top, -w, uc2, u2, radiance, SF_REFLECTION, &bs)) break;
                    if (spec_is_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                    beta = beta * (bs.f * abs_cos_theta(bs.wi) / bs.pdf);
                    w = bs.wi;
                } else {
                    if (!(i_flags(non_exit_top) & BX_SPECULAR)) {
                        Float wt = 1.0f;
                        if (!(i_flags(exit_top) & BX_SPECULAR)) wt = power_heuristic(wis.pdf, i_pdf(non_exit_top, -w, -wis.wi, SF_ALL));
                        f = f + beta * i_f(non_exit_top, -w, -wis.wi, radiance) * abs_cos_theta(wis.wi) * wt * tr(thickness, wis.wi) * wis.f / wis.pdf;
                    }
                    Float uc2 = r1(); V2 u2; u2.x = r1(); u2.y = r1();
                    BSDFSample bs;
                    if (!i_sample(non_exit_top, -w, uc2, u2, radiance, SF_REFLECTION, &bs)) break;
                    if (spec_is_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) break;
                    beta = beta * (bs.f * abs_cos_theta(bs.wi) / bs.pdf);
                    w = bs.wi;
                    if (!(i_flags(exit_top) & BX_SPECULAR)) {
                        Spec f_exit = i_f(exit_top, -w, wi, radiance);
                        if (!spec_is_zero(f_exit)) {
                            Float wt = 1.0f;
                            if (!(i_flags(non_exit_top) & BX_SPECULAR)) {
                                Float exit_pdf = i_pdf(exit_top, -w, wi, SF_TRANSMISSION);
                                wt = power_heuristic(bs.pdf, exit_pdf);
                            }
                            f = f + beta * tr(thickness, bs.wi) * f_exit * wt;
                        }
                    }
                }
            }
        }
        return f / (Float)n_samples;
    }

    // bxdf.rs:1249-1402.  Returns false for None; *proportional = pdf_is_proportional.
    bool sample_f(V3 wo, Float uc, V2 u, Rng& rng, BSDFSample* out, bool* proportional) const {
        const bool radiance = true;
        bool flip_wi = false;
        if (wo.z < 0.0f) { wo = -wo; flip_wi = true; }
        const bool entered_top = true;
        BSDFSample bs;
        if (!i_sample(entered_top, wo, uc, u, radiance, SF_ALL, &bs)) return false;
        if (spec_is_zero(bs.f) || bs.pdf == 0.0f || bs.wi.z == 0.0f) return false;
        if (bs.flags & BX_REFLECTION) {
            if (flip_wi) bs.wi = -bs.wi;
            *out = bs; *proportional = true;
            return true;
        }
        V3 w = bs.wi;
        bool specular_path = (bs.flags & BX_SPECULAR) != 0;
        auto r1 = [&]() { return fmin_(rng.get_1d(), next_float_down(1.0f)); };
        Spec f = bs.f * abs_cos_theta(bs.wi);
        Float pdf = bs.pdf;
        Float z = entered_top ? thickness : 0.0f;
        for (int depth = 0; depth < max_depth; ++depth) {
            Float rr_beta = spec_max(f) / pdf;
            if (depth > 3 && rr_beta < 0.25f) {
                Float q = fmax_(0.0f, 1.0f - rr_beta);
                if (r1() < q) return false;
                pdf *= 1.0f - q;
            }
            if (w.z == 0.0f) return false;
            if (!spec_is_zero(albedo)) {
                Float sigma_t = 1.0f;
                Float dz = sample_exponential(r1(), sigma_t / abs_cos_theta(w));
                Float zp = w.z > 0.0f ? (z + dz) : (z - dz);
                if (zp == z) return false;
                if (0.0f < zp && zp < thickness) {
                    V2 u2; u2.x = r1(); u2.y = r1();
                    V3 pwi; Float pp = sample_henyey_greenstein(-w, g, u2, &pwi);
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
            Float uc2 = r1(); V2 u2; u2.x = r1(); u2.y = r1();
            BSDFSample b2;
            if (!i_sample(iface_top, -w, uc2, u2, radiance, SF_ALL, &b2)) return false;
            if (spec_is_zero(b2.f) || b2.pdf == 0.0f || b2.wi.z == 0.0f) return false;
            f = f * b2.f;
            pdf *= b2.pdf;
            specular_path = specular_path && ((b2.flags & BX_SPECULAR) != 0);
            w = b2.wi;
            if (b2.flags & BX_TRANSMISSION) {
                int fl = same_hemisphere(wo, w) ? BX_REFLECTION : BX_TRANSMISSION;
                fl |= specular_path ? BX_SPECULAR : BX_GLOSSY;
                if (flip_wi) w = -w;
                out->f = f; out->wi = w; out->pdf = pdf; out->flags = fl; out->eta = 1.0f; *proportional = true;
                return true;
            }
            f = f * abs_cos_theta(b2.wi);
        }
        return false;
    }

    // bxdf.rs:1404-1584
    Float pdf(V3 wo, V3 wi, Rng& rng) const {
        const bool radiance = true;
        if (wo.z < 0.0f) { wo = -wo; wi = -wi; }
        auto r1 = [&]() { return fmin_(rng.get_1d(), next_float_down(1.0f)); };
        const bool entered_top = true;
        Float pdf_sum = 0.0f;
        if (same_hemisphere(wo, wi)) pdf_sum += (Float)n_samples * i_pdf(entered_top, wo, wi, SF_REFLECTION);
        for (int s = 0; s < n_samples; ++s) {
            if (same_hemisphere(wo, wi)) {
                const bool r_top = !entered_top, t_top = entered_top;
                Float uc = r1(); V2 u; u.x = r1(); u.y = r1();
                BSDFSample wos; bool has_wos = i_sample(t_top, wo, uc, u, radiance, SF_TRANSMISSION, &wos);
                uc = r1(); u.x = r1(); u.y = r1();
                BSDFSample wis; bool has_wis = i_sample(t_top, wi, uc, u, !radiance, SF_TRANSMISSION, &wis);
                if (has_wos && has_wis && !spec_is_zero(wos.f) && wos.pdf > 0.0f && !spec_is_zero(wis.f) && wis.pdf > 0.0f) {
                    if (!(i_flags(t_top) & (BX_DIFFUSE | BX_GLOSSY))) pdf_sum += i_pdf(r_top, -wos.wi, -wis.wi, SF_ALL);
                    else {
                        uc = r1(); u.x = r1(); u.y = r1();
                        BSDFSample rs;
                        if (i_sample(r_top, -wos.wi, uc, u, radiance, SF_ALL, &rs)) {
                            if (!(i_flags(r_top) & (BX_DIFFUSE | BX_GLOSSY))) pdf_sum += i_pdf(t_top, -rs.wi, wi, SF_ALL);
                            else {
                                Float r_pdf = i_pdf(r_top, -wos.wi, -wis.wi, SF_ALL);
                                Float wt = power_heuristic(wis.pdf, r_pdf);
                                pdf_sum += wt * r_pdf;
                                Float t_pdf = i_pdf(t_top, -rs.wi, wi, SF_ALL);
                                wt = power_heuristic(rs.pdf, t_pdf);
                                pdf_sum += wt * t_pdf;
                            }
                        }
                    }
                }
            } else {
                const bool to_top = entered_top, ti_top = !entered_top;
                Float uc = r1(); V2 u; u.x = r1(); u.y = r1();
                BSDFSample wos;
                if (!i_sample(to_top, wo, uc, u, radiance, SF_ALL, &wos)) continue;
                if (spec_is_zero(wos.f) || wos.pdf == 0.0f || wos.wi.z == 0.0f || (wos.flags & BX_REFLECTION)) continue;
                uc = r1(); u.x = r1(); u.y = r1();
                BSDFSample wis;
                if (!i_sample(ti_top, wi, uc, u, !radiance, SF_ALL, &wis)) continue;
                if (spec_is_zero(wis.f) || wis.pdf == 0.0f || wis.wi.z == 0.0f || (wis.flags & BX_REFLECTION)) continue;
                if (i_flags(to_top) & BX_SPECULAR) pdf_sum += i_pdf(ti_top, -wos.wi, wi, SF_ALL);
                else if (i_flags(ti_top) & BX_SPECULAR) pdf_sum += i_pdf(to_top, wo, -wis.wi, SF_ALL);
                else pdf_sum += (i_pdf(to_top, wo, -wis.wi, SF_ALL) + i_pdf(ti_top, -wos.wi, wi, SF_ALL)) / 2.0f;
            }
        }
        return lerp(0.9f, 1.0f / (4.0f * PI_F), pdf_sum / (Float)n_samples);
    }
};

}  // namespace orc
