// ORACLE -- TEST INFRASTRUCTURE ONLY.
// CPU restatement of jalberse/shimmer's `path` integrator hot path
// (ImageTileIntegrator::render -> evaluate_pixel_sample -> PathIntegrator::li -> sample_ld
//  -> RgbFilm::add_sample; integrator.rs:227-396,748-963, film.rs:548-574).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
// legs may load liborc; the product library never does.
//
// PARITY STATUS: "pinned against the reference's own unit-test vectors" for the arithmetic
// that has any (tests/test_oracle_kat.py lists them with file:line); the reference binary
// itself cannot be built in this image, and rand::SmallRng / num-complex are
// restated from their published algorithms => those two are "parity unpinned".
#include <cstring>
#include "orc_shading.h"
#include <atomic>
#include <thread>
#include <chrono>
#include <vector>
#include <cstdio>
#include <algorithm>

using namespace orc;

namespace {

struct PathCounters { Counters c; };
float g_debug_perturb = 0.0f;

// Ray::spawn_ray_to_both_offset ray.rs:83-99 via Interaction::spawn_ray_to_interaction interaction.rs:81-85
inline Ray spawn_ray_to_both_offset(const P3fi& p_from, V3 n_from, const P3fi& p_to, V3 n_to) {
    V3 pf = offset_ray_origin(p_from, n_from, p3fi_mid(p_to) - p3fi_mid(p_from));
    V3 pt = offset_ray_origin(p_to, n_to, pf - p3fi_mid(p_to));
    Ray r; r.o = pf; r.d = pt - pf; return r;
}

struct PathCtx {
    const Scene* sc;
    const SgRenderParams* rp;
    Counters* ctr;
    std::vector<float>* ray_log = nullptr;     // debug (orc_path_rays): o, d, t_max, any_hit, hit prim, hit t per traced ray
};
static inline bool logged_intersect(const PathCtx& pc, const Ray& ray, Float t_max, bool any_hit, Hit* hit) {
    const bool found = bvh_intersect(*pc.sc, ray, t_max, any_hit, hit, pc.ctr);
    if (pc.ray_log) {
        const float rec[10] = {ray.o.x, ray.o.y, ray.o.z, ray.d.x, ray.d.y, ray.d.z, t_max, any_hit ? 1.0f : 0.0f, found ? (float)hit->prim : -1.0f, found ? hit->th.t : 0.0f};
        pc.ray_log->insert(pc.ray_log->end(), rec, rec + 10);
    }
    return found;
}

// PathIntegrator::sample_ld integrator.rs:897-963
// Seed for the LayeredBxDF's private generator (the reference uses from_entropy, bxdf.rs:1011): derived from the
// path stream's current state and the call site (1 = f in sample_ld, 2 = pdf in sample_ld, 3 = sample_f, 4 = pdf after
// sample_f) WITHOUT advancing the path stream.  Same definition in sg_wavefront.cuh.
static inline uint64_t layer_seed(const Rng& rng, uint64_t site) { return mix64(rng.s[0] ^ (site * 0x9e3779b97f4a7c15ULL)); }
// Options::force_diffuse, interaction.rs:258-273: the material's BSDF is replaced by DiffuseBxDF(rho_hd(wo, [get_1d], [get_2d])) on the same
// shading frame.  BxDFI::rho_hd bxdf.rs:49-71 with one sample: 0 when wo.z == 0, else f |cos wi| / pdf of one BxDF-level sample_f
// (no BSDF-level rejection tests), kept only when pdf > 0.  A layered BxDF's private generator is seeded at site 6.
static inline void force_diffuse(BSDF& b, V3 wo_render, Rng& rng) {
    b.layer_seed = layer_seed(rng, 6);
    const Float uc = rng.get_1d();
    V2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
    Spec r = spec_const(0.0f);
    const V3 wo = b.to_local(wo_render);
    if (wo.z != 0.0f) {
        BSDFSample bs;
        if (b.sample_local(wo, uc, u2, &bs) && bs.pdf > 0.0f) r = r + bs.f * abs_cos_theta(bs.wi) / bs.pdf;
        r = r / 1.0f;
    }
    b.kind = SG_MATERIAL_DIFFUSE; b.r = r; b.k = spec_const(0.0f); b.eta = 1.0f; b.mf = TR::make(0.0f, 0.0f);
}

static SurfaceInteraction hit_interaction(const Scene& sc, const Hit& hit, const Ray& ray);

static Spec sample_ld(const PathCtx& pc, const SurfaceInteraction& intr, BSDF& bsdf, const Wavelengths& lambda, Rng& rng) {
    const SgSceneDesc* D = pc.sc->d;
    LightSampleContext ctx; ctx.pi = intr.pi; ctx.n = intr.n; ctx.ns = intr.sn;
    int flags = bsdf.flags();
    bool refl = flags & BX_REFLECTION, trans = flags & BX_TRANSMISSION;
    if (refl && !trans) ctx.pi = p3fi_exact(offset_ray_origin(intr.pi, intr.n, intr.wo));
    else if (trans && !refl) ctx.pi = p3fi_exact(offset_ray_origin(intr.pi, intr.n, -intr.wo));
    Float u = rng.get_1d();
    V2 u_light; u_light.x = rng.get_1d(); u_light.y = rng.get_1d();
    if (D->n_lights == 0) return spec_const(0.0f);
    // UniformLightSampler::sample_light light_sampler.rs:91-103 (`as usize` saturates)
    Float fl = u * (Float)D->n_lights;
    uint32_t li = fl != fl ? 0u : (fl <= 0.0f ? 0u : (fl >= 4294967296.0f ? 0xffffffffu : (uint32_t)fl));
    if (li > D->n_lights - 1) li = D->n_lights - 1;
    Float p_choose = 1.0f / (Float)D->n_lights;
    const SgLight& lt = D->lights[li];
    LightLiSample ls;
    if (!light_sample_li(*pc.sc, lt, ctx, u_light, lambda, &ls)) return spec_const(0.0f);
    if (spec_is_zero(ls.l) || ls.pdf == 0.0f) return spec_const(0.0f);
    V3 wo = intr.wo, wi = ls.wi;
    bsdf.layer_seed = layer_seed(rng, 1);
    Spec f = bsdf.f(wo, wi) * abs_dot(wi, intr.sn);
    if (spec_is_zero(f)) return spec_const(0.0f);
    Ray sray = spawn_ray_to_both_offset(intr.pi, intr.n, ls.p_light, ls.n_light);   // IntegratorBase::unoccluded :114-116
    Hit h;
    if (pc.ctr) pc.ctr->shadow++;
    if (logged_intersect(pc, sray, 1.0f - 0.0001f, true, &h)) return spec_const(0.0f);
    Float p_l = p_choose * ls.pdf;
    if (lt.kind == SG_LIGHT_POINT) return ls.l * f / p_l;
    bsdf.layer_seed = layer_seed(rng, 2);
    Float p_b = bsdf.pdf(wo, wi);
    Float w_l = power_heuristic(p_l, p_b);
    return w_l * ls.l * f / p_l;
}

// PathIntegrator::li integrator.rs:748-895
static Spec path_li(const PathCtx& pc, Ray ray, AuxRays aux, Wavelengths& lambda, Rng& rng) {
    const SgSceneDesc* D = pc.sc->d;
    Spec L = spec_const(0.0f), beta = spec_const(1.0f);
    int depth = 0;
    Float p_b = 1.0f, eta_scale = 1.0f;
    bool specular_bounce = false, any_non_specular_bounces = false;
    LightSampleContext prev_ctx; prev_ctx.pi = p3fi_exact(v3(0, 0, 0)); prev_ctx.n = v3(0, 0, 0); prev_ctx.ns = v3(0, 0, 0);
    for (;;) {
        Hit hit;
        if (pc.ctr) pc.ctr->closest++;
        bool found = logged_intersect(pc, ray, F_INF, false, &hit);
        if (!found) {
            for (uint32_t i = 0; i < D->n_lights; ++i) {                          // :779-792
                const SgLight& lt = D->lights[i];
                if (!light_is_infinite(lt)) continue;
                Spec le = light_le(D, lt, ray.d, lambda);                         // light.rs:792-794, :907-911
                if (depth == 0 || specular_bounce) L = L + beta * le;
                else {
                    Float p_l = (1.0f / (Float)D->n_lights) * light_pdf_li(*pc.sc, lt, prev_ctx, ray.d);
                    Float w_b = power_heuristic(p_b, p_l);
                    L = L + beta * w_b * le;
                }
            }
            break;
        }
        SurfaceInteraction si = hit_interaction(*pc.sc, hit, ray);                 // sphere.rs:286-293, bilinear_patch.rs:496-509, primitive.rs:155-169, triangle.rs:529-535
        if (si.light >= 0) {                                                       // :798-813
            const SgLight& lt = D->lights[si.light];
            Spec le = light_l(D, lt, si.n, -ray.d, lambda);
            if (!spec_is_zero(le)) {
                if (depth == 0 || specular_bounce) L = L + beta * le;
                else {
                    Float p_l = (1.0f / (Float)D->n_lights) * light_pdf_li(*pc.sc, lt, prev_ctx, ray.d);
                    Float w_l = power_heuristic(p_b, p_l);
                    L = L + beta * w_l * le;
                }
            }
        }
        BSDF bsdf = get_bsdf(D, si, lambda, aux, pc.rp, layer_seed(rng, 5));      // :816
        if (pc.rp->option_flags & SG_OPT_FORCE_DIFFUSE) force_diffuse(bsdf, si.wo, rng);
        if (pc.rp->regularize && any_non_specular_bounces) { bsdf.mf.regularize(); bsdf.lay.mf.regularize(); bsdf.lay.mfb.regularize(); }  // :825-828 (LayeredBxDF::regularize: top + bottom, bxdf.rs:1616-1619)
        if (depth == pc.rp->max_depth) break;
        depth += 1;
        if (bsdf.flags() & (BX_DIFFUSE | BX_GLOSSY)) {                            // :837-841
            Spec ld = sample_ld(pc, si, bsdf, lambda, rng);
            L = L + beta * ld;
        }
        V3 wo = -ray.d;
        Float u = rng.get_1d();
        V2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
        BSDFSample bs;
        bsdf.layer_seed = layer_seed(rng, 3);
        if (!bsdf.sample_f(wo, u, u2, &bs)) break;
        beta = beta * (bs.f * abs_dot(bs.wi, si.sn) / bs.pdf);                    // :859
        if (bs.proportional) { bsdf.layer_seed = layer_seed(rng, 4); p_b = bsdf.pdf(wo, bs.wi); }   // :860-865
        else p_b = bs.pdf;
        specular_bounce = (bs.flags & BX_SPECULAR) != 0;
        any_non_specular_bounces |= !specular_bounce;
        if (bs.flags & BX_TRANSMISSION) eta_scale *= sqr(bs.eta);
        prev_ctx.pi = si.pi; prev_ctx.n = si.n; prev_ctx.ns = si.sn;
        if (D->n_textures > 0) aux = spawn_differentials(si, aux, bs.wi, bs.flags, bs.eta);   // spawn_ray_with_differentials :434-502
        ray.o = offset_ray_origin(si.pi, si.n, bs.wi); ray.d = bs.wi;             // spawn_ray interaction.rs:72-79
        if (g_debug_perturb != 0.0f && depth == 1) ray.d.x *= 1.0f + g_debug_perturb;  // debug only (orc_set_debug_perturb): sensitivity experiments
        if (std::isfinite(eta_scale)) {                                           // :878-891
            Spec rr_beta = beta * eta_scale;
            if (spec_max(rr_beta) < 1.0f && depth > 1) {
                Float q = fmax_(0.0f, 1.0f - spec_max(rr_beta));
                if (rng.get_1d() < q) break;
                beta = beta / (1.0f - q);
            }
        }
    }
    return L;
}

// Hit -> SurfaceInteraction (shared by the three integrators): Primitive::intersect results of the shapes on this path
static SurfaceInteraction hit_interaction(const Scene& sc, const Hit& hit, const Ray& ray) {
    const SgSceneDesc* D = sc.d;
    const SgPrimitive& prim = D->primitives[hit.prim];
    // TransformedPrimitive::intersect (primitive.rs:155-169): the shape is intersected with the instance-space ray, so its
    // interaction is built there (wo = -(M^-1 d)) and then mapped to render space with Transform::apply(SurfaceInteraction)
    V3 d = ray.d;
    if (hit.inst >= 0) {
        const float* mi = D->instances[hit.inst].primitive_from_render;
        d = v3(mi[0] * ray.d.x + mi[1] * ray.d.y + mi[2] * ray.d.z, mi[4] * ray.d.x + mi[5] * ray.d.y + mi[6] * ray.d.z,
               mi[8] * ray.d.x + mi[9] * ray.d.y + mi[10] * ray.d.z);
    }
    SurfaceInteraction si;
    if (prim.mesh == SG_PRIM_SPHERE) {                                         // Sphere::intersect sphere.rs:286-293
        V3 p_obj = v3(hit.th.b0, hit.th.b1, hit.th.b2);
        Float phi = std::atan2(p_obj.y, p_obj.x); if (phi < 0.0f) phi += 2.0f * PI_F;
        si = sphere_interaction(D, D->spheres[prim.tri], p_obj, phi, -d);
    } else if (D->meshes[prim.mesh].flags & SG_MESH_BILINEAR) {                 // BilinearPatch::intersect bilinear_patch.rs:496-509
        si = patch_interaction(sc, prim.mesh, prim.tri, hit.th.b0, hit.th.b1, -d);
    } else si = interaction_from_intersection(sc, prim.mesh, prim.tri, hit.th, -d);   // triangle.rs:529-535
    if (hit.inst >= 0) transform_interaction(D, D->instances[hit.inst], si);
    si.material = (int32_t)prim.material; si.light = prim.light;
    return si;
}
// sample_uniform_hemisphere sampling.rs:295-304 (around +z of the space it is used in)
static V3 sample_uniform_hemisphere(V2 u) {
    const Float z = u.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * PI_F * u.y;
    return v3(r * std::cos(phi), r * std::sin(phi), z);
}

// SimplePathIntegrator::li integrator.rs:585-727: no MIS, no Russian roulette; lights sampled with COMPLETE pdfs
// (allow_incomplete_pdf = false, :652-656) from the un-nudged LightSampleContext::from(&isect).
static Spec simple_path_li(const PathCtx& pc, Ray ray, AuxRays aux, Wavelengths& lambda, Rng& rng) {
    const SgSceneDesc* D = pc.sc->d;
    const bool sample_lights = (pc.rp->integrator_flags & SG_SIMPLEPATH_SAMPLE_LIGHTS) != 0, sample_bsdf = (pc.rp->integrator_flags & SG_SIMPLEPATH_SAMPLE_BSDF) != 0;
    Spec L = spec_const(0.0f), beta = spec_const(1.0f);
    bool specular_bounce = true;
    int depth = 0;
    while (!spec_is_zero(beta)) {
        Hit hit;
        if (pc.ctr) pc.ctr->closest++;
        if (!bvh_intersect(*pc.sc, ray, F_INF, false, &hit, pc.ctr)) {
            if (!sample_lights || specular_bounce)
                for (uint32_t i = 0; i < D->n_lights; ++i) if (light_is_infinite(D->lights[i])) L = L + beta * light_le(D, D->lights[i], ray.d, lambda);
            break;
        }
        SurfaceInteraction si = hit_interaction(*pc.sc, hit, ray);
        if ((!sample_lights || specular_bounce) && si.light >= 0) L = L + beta * light_l(D, D->lights[si.light], si.n, -ray.d, lambda);
        if (depth == pc.rp->max_depth) break;
        depth += 1;
        BSDF bsdf = get_bsdf(D, si, lambda, aux, pc.rp, layer_seed(rng, 5));
        if (pc.rp->option_flags & SG_OPT_FORCE_DIFFUSE) force_diffuse(bsdf, si.wo, rng);
        const V3 wo = -ray.d;
        if (sample_lights && D->n_lights > 0) {                                      // UniformLightSampler::sample_light light_sampler.rs:91-103
            const Float ul = rng.get_1d();
            Float fl = ul * (Float)D->n_lights;
            uint32_t li = fl != fl ? 0u : (fl <= 0.0f ? 0u : (fl >= 4294967296.0f ? 0xffffffffu : (uint32_t)fl));
            if (li > D->n_lights - 1) li = D->n_lights - 1;
            const Float p_choose = 1.0f / (Float)D->n_lights;
            V2 u_light; u_light.x = rng.get_1d(); u_light.y = rng.get_1d();
            LightSampleContext ctx; ctx.pi = si.pi; ctx.n = si.n; ctx.ns = si.sn;
            LightLiSample ls;
            if (light_sample_li(*pc.sc, D->lights[li], ctx, u_light, lambda, &ls, false) && !spec_is_zero(ls.l) && ls.pdf > 0.0f) {
                bsdf.layer_seed = layer_seed(rng, 1);
                const Spec f = bsdf.f(wo, ls.wi) * abs_dot(ls.wi, si.sn);
                if (!spec_is_zero(f)) {
                    Ray sray = spawn_ray_to_both_offset(si.pi, si.n, ls.p_light, ls.n_light);
                    Hit h;
                    if (pc.ctr) pc.ctr->shadow++;
                    if (!bvh_intersect(*pc.sc, sray, 1.0f - 0.0001f, true, &h, pc.ctr)) L = L + beta * f * ls.l / (p_choose * ls.pdf);
                }
            }
        } else if (sample_lights) (void)rng.get_1d();                                // sample_light(u) still draws u; no light -> None before get_2d
        if (sample_bsdf) {
            const Float u = rng.get_1d();
            V2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
            BSDFSample bs;
            bsdf.layer_seed = layer_seed(rng, 3);
            if (!bsdf.sample_f(wo, u, u2, &bs)) break;
            beta = beta * (bs.f * abs_dot(bs.wi, si.sn) / bs.pdf);
            specular_bounce = (bs.flags & BX_SPECULAR) != 0;
            ray.o = offset_ray_origin(si.pi, si.n, bs.wi); ray.d = bs.wi;
        } else {
            const int flags = bsdf.flags();
            const bool refl = flags & BX_REFLECTION, trans = flags & BX_TRANSMISSION;
            V2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
            V3 wi; const Float pdf = INV_4PI;                                        // uniform_sphere_pdf and (sic) uniform_hemisphere_pdf, sampling.rs:291-293,306-308
            if (refl && trans) wi = sample_uniform_sphere(u2);
            else {
                wi = sample_uniform_hemisphere(u2);
                if ((refl && dot(wo, si.n) * dot(wi, si.n) < 0.0f) || (trans && dot(wo, si.n) * dot(wi, si.n) > 0.0f)) wi = -wi;
            }
            bsdf.layer_seed = layer_seed(rng, 3);
            beta = beta * (bsdf.f(wo, wi) * abs_dot(wi, si.sn) / pdf);
            specular_bounce = false;
            ray.o = offset_ray_origin(si.pi, si.n, wi); ray.d = wi;
        }
        aux.has = false;                                                             // Interaction::spawn_ray: a plain Ray, no differentials
    }
    return L;
}

// RandomWalkIntegrator::li_random_walk integrator.rs:493-567, recursion as written
static Spec random_walk_li(const PathCtx& pc, Ray ray, AuxRays aux, Wavelengths& lambda, Rng& rng, int depth) {
    const SgSceneDesc* D = pc.sc->d;
    Hit hit;
    if (pc.ctr) pc.ctr->closest++;
    if (!bvh_intersect(*pc.sc, ray, F_INF, false, &hit, pc.ctr)) {
        Spec le = spec_const(0.0f);
        for (uint32_t i = 0; i < D->n_lights; ++i) if (light_is_infinite(D->lights[i])) le = le + light_le(D, D->lights[i], ray.d, lambda);
        return le;
    }
    SurfaceInteraction si = hit_interaction(*pc.sc, hit, ray);
    const V3 wo = -ray.d;
    const Spec le = si.light >= 0 ? light_l(D, D->lights[si.light], si.n, wo, lambda) : spec_const(0.0f);
    if (depth == pc.rp->max_depth) return le;
    BSDF bsdf = get_bsdf(D, si, lambda, aux, pc.rp, layer_seed(rng, 5));
    V2 u; u.x = rng.get_1d(); u.y = rng.get_1d();
    const V3 wp = sample_uniform_sphere(u);
    bsdf.layer_seed = layer_seed(rng, 1);
    const Spec f = bsdf.f(wo, wp);
    if (spec_is_zero(f)) return le;
    const Spec fcos = f * abs_dot(wp, si.sn);
    Ray next; next.o = offset_ray_origin(si.pi, si.n, wp); next.d = wp;
    AuxRays none;
    return le + fcos * random_walk_li(pc, next, none, lambda, rng, depth + 1) / (1.0f / (4.0f * PI_F));
}

// evaluate_pixel_sample integrator.rs:326-396 + get_camera_sample sampling.rs:347-371 + BoxFilter::sample filter.rs:99-105
static void camera_stage(const SgSceneDesc* D, const SgRenderParams* rp, int px, int py, Rng& rng, Wavelengths* lambda, Ray* ray, Float* weight, AuxRays* aux = nullptr) {
    Float lu = (rp->option_flags & SG_OPT_DISABLE_WAVELENGTH_JITTER) ? 0.5f : rng.get_1d();
    *lambda = sample_visible(lu);
    V2 pu; pu.x = rng.get_1d(); pu.y = rng.get_1d();                  // get_pixel_2d, always consumed
    CameraSample cs;
    if (rp->option_flags & SG_OPT_DISABLE_PIXEL_JITTER) {
        cs.p_film.x = (Float)px + 0.5f; cs.p_film.y = (Float)py + 0.5f;
        cs.p_lens.x = 0.5f; cs.p_lens.y = 0.5f; cs.time = 0.5f; cs.filter_weight = 1.0f;
    } else {
        Float rx = D->film.filter_radius[0], ry = D->film.filter_radius[1];
        V2 fp = {lerp(pu.x, -rx, rx), lerp(pu.y, -ry, ry)};          // BoxFilter::sample
        cs.p_film.x = (Float)px + fp.x + 0.5f; cs.p_film.y = (Float)py + fp.y + 0.5f;
        cs.p_lens.x = rng.get_1d(); cs.p_lens.y = rng.get_1d();
        cs.time = rng.get_1d();
        cs.filter_weight = 1.0f;
    }
    *ray = camera_generate_ray(D->camera, cs, aux);
    if (aux && aux->has && !(rp->option_flags & SG_OPT_DISABLE_PIXEL_JITTER)) {       // integrator.rs:355-361 + Ray::scale_differentials ray.rs:137-145
        Float s = fmax_(0.125f, 1.0f / std::sqrt((Float)rp->samples_per_pixel));
        aux->rxo = ray->o + (aux->rxo - ray->o) * s; aux->ryo = ray->o + (aux->ryo - ray->o) * s;
        aux->rxd = ray->d + (aux->rxd - ray->d) * s; aux->ryd = ray->d + (aux->ryd - ray->d) * s;
    }
    *weight = cs.filter_weight;
}

static void eval_sample(const PathCtx& pc, int px, int py, Rng& rng, SgFilmPixel* film) {
    const SgSceneDesc* D = pc.sc->d;
    Wavelengths lambda; Ray ray; Float weight; AuxRays aux;
    camera_stage(D, pc.rp, px, py, rng, &lambda, &ray, &weight, D->n_textures > 0 ? &aux : nullptr);
    Spec L;                                               // camera_ray.weight == 1 (camera.rs:997-1000)
    if (pc.rp->integrator == SG_INTEGRATOR_SIMPLE_PATH) L = simple_path_li(pc, ray, aux, lambda, rng);
    else if (pc.rp->integrator == SG_INTEGRATOR_RANDOM_WALK) L = random_walk_li(pc, ray, aux, lambda, rng, 0);
    else L = path_li(pc, ray, aux, lambda, rng);
    int W = D->film.pixel_bounds[2] - D->film.pixel_bounds[0];
    SgFilmPixel* pxl = film + (size_t)(py - D->film.pixel_bounds[1]) * W + (px - D->film.pixel_bounds[0]);
    film_add_sample(D, pxl, L, lambda, weight);
}

}  // namespace

extern "C" {

const char* orc_header(void) {
    return "shimmer oracle: CPU restatement, TEST INFRASTRUCTURE ONLY; pinned against the reference's unit-test "
           "vectors; rand::SmallRng and num-complex restated from published algorithms (parity unpinned)";
}

void orc_sampler_fill(uint64_t seed, int raw, uint32_t pixel_index, uint32_t sample_index, int64_t n, float* out) {
    Rng r; r.seed_from_u64(raw ? seed : stream_key(seed, pixel_index, sample_index));
    for (int64_t i = 0; i < n; ++i) out[i] = r.get_1d();
}
void orc_rng_u64(uint64_t seed, int from_state, const uint64_t* state, int64_t n, uint64_t* out, uint64_t* state_out) {
    Rng r;
    if (from_state) for (int i = 0; i < 4; ++i) r.s[i] = state[i]; else r.seed_from_u64(seed);
    if (state_out) for (int i = 0; i < 4; ++i) state_out[i] = r.s[i];
    for (int64_t i = 0; i < n; ++i) out[i] = r.next_u64();
}

// aggregate.rs:207-468.  prim_bounds: 6 floats (min xyz, max xyz) per primitive in input order.
// out_nodes must hold 2*n-1 nodes, out_order n entries.  Returns node count.
int64_t orc_bvh_build(int64_t n, const float* prim_bounds, SgBvhNode* out_nodes, uint32_t* out_order) {
    std::vector<BuildPrim> prims((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        prims[i].index = (uint32_t)i;
        for (int a = 0; a < 3; ++a) { prims[i].bmin[a] = prim_bounds[6 * i + a]; prims[i].bmax[a] = prim_bounds[6 * i + 3 + a]; }
    }
    BvhBuilder b; b.nodes.reserve((size_t)(2 * n)); b.order.reserve((size_t)n);
    if (n > 0) b.build(prims.data(), (size_t)n);
    std::memcpy(out_nodes, b.nodes.data(), b.nodes.size() * sizeof(SgBvhNode));
    std::memcpy(out_order, b.order.data(), b.order.size() * sizeof(uint32_t));
    return (int64_t)b.nodes.size();
}

void orc_trace(const SgSceneDesc* desc, int64_t n, const float* o, const float* d, const float* t_max, int any_hit,
               SgHit* out, SgStats* stats, int n_threads) {
    Scene sc(desc);
    if (n_threads < 1) n_threads = 1;
    std::vector<Counters> ctrs((size_t)n_threads);
    auto work = [&](int tid) {
        Counters& c = ctrs[tid];
        for (int64_t i = tid; i < n; i += n_threads) {
            Ray r; r.o = v3(o[3 * i], o[3 * i + 1], o[3 * i + 2]); r.d = v3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
            Hit h;
            bool found = bvh_intersect(sc, r, t_max[i], any_hit != 0, &h, &c);
            SgHit& oh = out[i];
            std::memset(&oh, 0, sizeof oh);
            if (!found) { oh.prim = -1; continue; }
            if (any_hit) { oh.prim = 0; continue; }
            oh.prim = h.prim; oh.t = h.th.t; oh.b0 = h.th.b0; oh.b1 = h.th.b1; oh.b2 = h.th.b2;
            const SgPrimitive& pr = desc->primitives[h.prim];
            if (pr.mesh == SG_PRIM_SPHERE) {
                V3 p_obj = v3(h.th.b0, h.th.b1, h.th.b2);
                Float phi = std::atan2(p_obj.y, p_obj.x); if (phi < 0.0f) phi += 2.0f * PI_F;
                SurfaceInteraction ssi = sphere_interaction(desc, desc->spheres[pr.tri], p_obj, phi, -r.d);
                oh.ng[0] = ssi.n.x; oh.ng[1] = ssi.n.y; oh.ng[2] = ssi.n.z;
                continue;
            }
            if (desc->meshes[pr.mesh].flags & SG_MESH_BILINEAR) {
                SurfaceInteraction psi = patch_interaction(sc, pr.mesh, pr.tri, h.th.b0, h.th.b1, -r.d);
                V3 q[4]; sc.patch_points(pr.mesh, pr.tri, q);                   // geometric normal before shading-normal face-forwarding
                (void)q;
                oh.ng[0] = psi.n.x; oh.ng[1] = psi.n.y; oh.ng[2] = psi.n.z;
                continue;
            }
            SurfaceInteraction si = interaction_from_intersection(sc, pr.mesh, pr.tri, h.th, -r.d);
            // geometric normal as produced by triangle.rs:407-412 (before any shading-normal face-forwarding)
            V3 p0, p1, p2; sc.tri_points(pr.mesh, pr.tri, &p0, &p1, &p2);
            V3 ng = normalize(cross(p0 - p2, p1 - p2));
            const SgMesh& m = desc->meshes[pr.mesh];
            if (((m.flags & SG_MESH_REVERSE_ORIENTATION) != 0) ^ ((m.flags & SG_MESH_SWAPS_HANDEDNESS) != 0)) ng = -ng;
            (void)si;
            oh.ng[0] = ng.x; oh.ng[1] = ng.y; oh.ng[2] = ng.z;
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (auto& c : ctrs) { stats->nodes_visited += c.nodes; stats->tris_tested += c.tris; }
        if (any_hit) stats->shadow_rays = (uint64_t)n; else stats->closest_hit_rays = (uint64_t)n;
    }
}

void orc_camera_rays(const SgSceneDesc* desc, const SgRenderParams* rp, int64_t n, const int32_t* pixel_xy,
                     const int32_t* sample_index, float* out_rays, float* out_lambda) {
    for (int64_t i = 0; i < n; ++i) {
        int px = pixel_xy[2 * i], py = pixel_xy[2 * i + 1];
        Rng rng; rng.seed_from_u64(stream_key(rp->seed, (uint32_t)(py * desc->film.full_resolution[0] + px), (uint32_t)sample_index[i]));
        Wavelengths lam; Ray ray; Float w;
        camera_stage(desc, rp, px, py, rng, &lam, &ray, &w);
        float* r = out_rays + 6 * i;
        r[0] = ray.o.x; r[1] = ray.o.y; r[2] = ray.o.z; r[3] = ray.d.x; r[4] = ray.d.y; r[5] = ray.d.z;
        for (int k = 0; k < 4; ++k) { out_lambda[8 * i + k] = lam.lambda[k]; out_lambda[8 * i + 4 + k] = lam.pdf[k]; }
    }
}

// ImageTileIntegrator::render integrator.rs:227-321.
// stream_mode 0: deterministic (pixel,sample) streams (shared with the CUDA path).
// stream_mode 1: the reference's behaviour -- one generator per worker thread, cloned from the
//                same seed, consumed in tile-scheduling order (integrator.rs:252-263).
// film is ADDED to (caller zeroes it).  Returns seconds spent in the tile loop.
double orc_render(const SgSceneDesc* desc, const SgRenderParams* rp, SgFilmPixel* film, SgStats* stats,
                  int n_threads, int stream_mode) {
    Scene sc(desc);
    if (n_threads < 1) n_threads = 1;
    const int x0 = desc->film.pixel_bounds[0], y0 = desc->film.pixel_bounds[1];
    const int x1 = desc->film.pixel_bounds[2], y1 = desc->film.pixel_bounds[3];
    // Tile::tile(bounds, 8, 8) tile.rs:21-104: row-major tiles, remainders kept
    struct TileB { int x0, y0, x1, y1; };
    std::vector<TileB> tiles;
    for (int ty = y0; ty < y1; ty += 8) for (int tx = x0; tx < x1; tx += 8)
        tiles.push_back({tx, ty, std::min(tx + 8, x1), std::min(ty + 8, y1)});
    std::vector<Counters> ctrs((size_t)n_threads);
    std::vector<Rng> thread_rng((size_t)n_threads);
    for (auto& r : thread_rng) r.seed_from_u64(rp->seed);
    auto t_begin = std::chrono::steady_clock::now();
    // waves 1,1,2,4,...,64 (:231-233,306-308), restricted to [sample_begin, sample_end)
    int wave_start = 0, wave_end = 1, next_wave = 1;
    const int spp_hi = rp->sample_end;
    while (wave_start < spp_hi) {
        int ws = std::max(wave_start, rp->sample_begin), we = std::min(wave_end, spp_hi);
        if (ws < we) {
            std::atomic<size_t> next_tile(0);
            auto work = [&](int tid) {
                PathCtx pc; pc.sc = &sc; pc.rp = rp; pc.ctr = &ctrs[tid];
                for (;;) {
                    size_t ti = next_tile.fetch_add(1);
                    if (ti >= tiles.size()) break;
                    const TileB& t = tiles[ti];
                    for (int x = t.x0; x < t.x1; ++x) for (int y = t.y0; y < t.y1; ++y)       // x outer, y inner :257-258
                        for (int s = ws; s < we; ++s) {
                            if (stream_mode == 0) {
                                Rng rng; rng.seed_from_u64(stream_key(rp->seed, (uint32_t)(y * desc->film.full_resolution[0] + x), (uint32_t)s));
                                eval_sample(pc, x, y, rng, film);
                            } else {
                                eval_sample(pc, x, y, thread_rng[tid], film);
                            }
                        }
                }
            };
            std::vector<std::thread> th;
            for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
            work(0);
            for (auto& t : th) t.join();
        }
        wave_start = wave_end;
        wave_end = std::min(spp_hi, wave_end + next_wave);
        next_wave = std::min(2 * next_wave, 64);
    }
    double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
    if (stats) {
        std::memset(stats, 0, sizeof *stats);
        for (auto& c : ctrs) { stats->nodes_visited += c.nodes; stats->tris_tested += c.tris; stats->closest_hit_rays += c.closest; stats->shadow_rays += c.shadow; }
        stats->camera_paths = (uint64_t)(x1 - x0) * (uint64_t)(y1 - y0) * (uint64_t)std::max(0, rp->sample_end - rp->sample_begin);
        stats->render_ms = secs * 1e3;
    }
    return secs;
}

// RgbFilm::get_pixel_rgb film.rs:720-738 (splat term is zero on this path)
void orc_film_develop(const SgSceneDesc* desc, const SgFilmPixel* film, int64_t n, float* out_rgb) {
    const float* M = desc->film.output_rgb_from_sensor_rgb;
    for (int64_t i = 0; i < n; ++i) {
        Float rgb[3] = {(Float)film[i].rgb_sum[0], (Float)film[i].rgb_sum[1], (Float)film[i].rgb_sum[2]};
        if (film[i].weight_sum != 0.0) for (int c = 0; c < 3; ++c) rgb[c] /= (Float)film[i].weight_sum;
        for (int r = 0; r < 3; ++r) out_rgb[3 * i + r] = M[3 * r] * rgb[0] + M[3 * r + 1] * rgb[1] + M[3 * r + 2] * rgb[2];
    }
}

// half 2.2.1 `f16::from_f32` (software path; Cargo.lock pins half 2.2.1): IEEE 754 binary32 -> binary16, round to nearest
// even, overflow -> inf, NaN keeps a quiet payload; restated from the published algorithm (third-party crate, not in
// /root/reference) and pinned against numpy's float16 conversion in tests/test_oracle_kat.py.
static uint16_t f16_bits_from_f32(float f) {
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u, man = x & 0x007fffffu;
    const int32_t exp = (int32_t)((x >> 23) & 0xffu);
    if (exp == 255) return (uint16_t)(sign | 0x7c00u | (man ? (0x0200u | (man >> 13)) : 0u));
    const int32_t he = exp - 127 + 15;
    if (he >= 31) return (uint16_t)(sign | 0x7c00u);
    if (he <= 0) {
        if (14 - he > 24) return (uint16_t)sign;                              // too small: +-0
        const uint32_t m = man | 0x00800000u;
        const int shift = 14 - he;                                            // 14..24
        uint32_t hm = m >> shift;
        const uint32_t round_bit = 1u << (shift - 1);
        if ((m & round_bit) && (m & (3u * round_bit - 1u))) hm++;             // RNE: round bit and (sticky or odd)
        return (uint16_t)(sign | hm);
    }
    uint32_t h = sign | ((uint32_t)he << 10) | (man >> 13);
    if ((man & 0x1000u) && (man & 0x2fffu)) h++;                              // may carry into the exponent (-> inf): correct
    return (uint16_t)h;
}
static float f32_from_f16_bits(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16; const uint32_t e = (h >> 10) & 31u, m = h & 0x3ffu;
    uint32_t x;
    if (e == 0) {
        if (m == 0) x = sign;
        else { int k = 0; uint32_t mm = m; while (!(mm & 0x400u)) { mm <<= 1; ++k; } x = sign | ((uint32_t)(127 - 15 - k + 1) << 23) | ((mm & 0x3ffu) << 13); }
    } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
    else x = sign | ((e - 15 + 127) << 23) | (m << 13);
    float f; std::memcpy(&f, &x, 4); return f;
}
void orc_f16_round(int64_t n, const float* in, float* out, uint16_t* bits) {
    for (int64_t i = 0; i < n; ++i) { const uint16_t b = f16_bits_from_f32(in[i]); if (bits) bits[i] = b; out[i] = f32_from_f16_bits(b); }
}
// RgbFilm::get_image film.rs:647-707 + Image::set_channel image.rs:648-661 + the row order of write_pfm image.rs:1350
void orc_film_get_image(const SgSceneDesc* desc, const SgFilmPixel* film, int32_t w, int32_t h, uint32_t flags, float* out_rgb) {
    const float* M = desc->film.output_rgb_from_sensor_rgb;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        const int64_t i = (int64_t)y * w + x;
        Float rgb[3] = {(Float)film[i].rgb_sum[0], (Float)film[i].rgb_sum[1], (Float)film[i].rgb_sum[2]};
        if (film[i].weight_sum != 0.0) for (int c = 0; c < 3; ++c) rgb[c] /= (Float)film[i].weight_sum;
        Float o[3];
        for (int r = 0; r < 3; ++r) o[r] = M[3 * r] * rgb[0] + M[3 * r + 1] * rgb[1] + M[3 * r + 2] * rgb[2];
        if (flags & SG_IMAGE_FP16) {
            const Float max_f16 = 65504.0f;
            Float mx = -INFINITY;
            for (int c = 0; c < 3; ++c) mx = std::fmax(mx, o[c]);            // Float::max ignores NaN
            if (mx > max_f16) {
                if (o[0] > max_f16) o[0] = max_f16;
                if (o[1] > max_f16) o[0] = max_f16;                           // sic, film.rs:683-685
                if (o[2] > max_f16) o[2] = max_f16;
            }
        }
        const int64_t dst = (flags & SG_IMAGE_BOTTOM_UP) ? (int64_t)(h - 1 - y) * w + x : i;
        for (int c = 0; c < 3; ++c) {
            Float v = o[c];
            if (std::isnan(v)) v = 0.0f;
            if (flags & SG_IMAGE_FP16) v = f32_from_f16_bits(f16_bits_from_f32(v));
            out_rgb[3 * dst + c] = v;
        }
    }
}

// ---- known-answer entry points (tests/test_oracle_kat.py) --------------------------
float orc_difference_of_products(float a, float b, float c, float d) { return difference_of_products(a, b, c, d); }
float orc_lerp(float t, float a, float b) { return lerp(t, a, b); }
float orc_next_float_up(float v) { return next_float_up(v); }
float orc_next_float_down(float v) { return next_float_down(v); }
float orc_gamma(int n) { return gamma_n(n); }
float orc_visible_wavelengths_pdf(float l) { return visible_wavelengths_pdf(l); }
float orc_sample_visible_wavelengths(float u) { return sample_visible_wavelengths(u); }
float orc_tr_d(float ax, float ay, const float* wm) { return TR::make(ax, ay).d(v3(wm[0], wm[1], wm[2])); }
float orc_tr_g(float ax, float ay, const float* wo, const float* wi) { return TR::make(ax, ay).g(v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2])); }
float orc_fresnel_dielectric(float c, float eta) { return fresnel_dielectric(c, eta); }
float orc_fresnel_complex(float c, float eta, float k) { return fresnel_complex(c, cx(eta, k)); }
float orc_blackbody(float lambda, float t) { return blackbody(lambda, t); }
float orc_spectrum_get(const SgSceneDesc* d, int id, float lambda) { return spectrum_get(d, id, lambda); }
// ---- texture / differential KAT entry points ----
void orc_rotate_from_to(const float* from, const float* to, float* out9) {
    rotate_from_to(v3(from[0], from[1], from[2]), v3(to[0], to[1], to[2]), out9);
}
float orc_sigmoid_poly_get(const float* c3, float lambda) { return sigmoid_poly_get(c3, lambda); }
void orc_rgb2spec_fetch(const SgSceneDesc* d, const float* rgb, float* out3) { rgb2spec_fetch(d, rgb, out3); }
// n lookups; q = u v dudx dudy dvdx dvdy per lookup; lambda4 per lookup; out4 per lookup (float textures: value replicated)
void orc_texture_eval(const SgSceneDesc* d, int tex, int as_float, int64_t n, const float* q, const float* lambda4, float* out4) {
    for (int64_t i = 0; i < n; ++i) {
        TexCoordCtx c; c.uv.x = q[6 * i]; c.uv.y = q[6 * i + 1]; c.dudx = q[6 * i + 2]; c.dudy = q[6 * i + 3]; c.dvdx = q[6 * i + 4]; c.dvdy = q[6 * i + 5];
        if (as_float) { Float v = eval_float_texture(d, tex, c); for (int k = 0; k < 4; ++k) out4[4 * i + k] = v; }
        else {
            Wavelengths w; for (int k = 0; k < 4; ++k) { w.lambda[k] = lambda4[4 * i + k]; w.pdf[k] = 1.0f; }
            Spec s = eval_spectrum_texture(d, tex, c, w);
            for (int k = 0; k < 4; ++k) out4[4 * i + k] = s.v[k];
        }
    }
}
// same with a full TextureEvalContext: pdp = p, dpdx, dpdy (9 floats per lookup) for the non-UV mappings (texture.rs:938-1035)
void orc_texture_eval_p(const SgSceneDesc* d, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* lambda4, float* out4) {
    for (int64_t i = 0; i < n; ++i) {
        TexCoordCtx c; c.uv.x = q[6 * i]; c.uv.y = q[6 * i + 1]; c.dudx = q[6 * i + 2]; c.dudy = q[6 * i + 3]; c.dvdx = q[6 * i + 4]; c.dvdy = q[6 * i + 5];
        c.p = v3(pdp[9 * i], pdp[9 * i + 1], pdp[9 * i + 2]); c.dpdx = v3(pdp[9 * i + 3], pdp[9 * i + 4], pdp[9 * i + 5]); c.dpdy = v3(pdp[9 * i + 6], pdp[9 * i + 7], pdp[9 * i + 8]);
        if (as_float) { Float v = eval_float_texture(d, tex, c); for (int k = 0; k < 4; ++k) out4[4 * i + k] = v; }
        else {
            Wavelengths w; for (int k = 0; k < 4; ++k) { w.lambda[k] = lambda4[4 * i + k]; w.pdf[k] = 1.0f; }
            Spec s = eval_spectrum_texture(d, tex, c, w);
            for (int k = 0; k < 4; ++k) out4[4 * i + k] = s.v[k];
        }
    }
}
// same with TextureEvalContext::n too (3 floats per lookup; pdp may be null): the direction-mix textures (texture.rs:295-310,:810-826)
void orc_texture_eval_ctx(const SgSceneDesc* d, int tex, int as_float, int64_t n, const float* q, const float* pdp, const float* nrm, const float* lambda4, float* out4) {
    for (int64_t i = 0; i < n; ++i) {
        TexCoordCtx c; c.uv.x = q[6 * i]; c.uv.y = q[6 * i + 1]; c.dudx = q[6 * i + 2]; c.dudy = q[6 * i + 3]; c.dvdx = q[6 * i + 4]; c.dvdy = q[6 * i + 5];
        if (pdp) { c.p = v3(pdp[9 * i], pdp[9 * i + 1], pdp[9 * i + 2]); c.dpdx = v3(pdp[9 * i + 3], pdp[9 * i + 4], pdp[9 * i + 5]); c.dpdy = v3(pdp[9 * i + 6], pdp[9 * i + 7], pdp[9 * i + 8]); }
        if (nrm) c.n = v3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
        if (as_float) { Float v = eval_float_texture(d, tex, c); for (int k = 0; k < 4; ++k) out4[4 * i + k] = v; }
        else {
            Wavelengths w; for (int k = 0; k < 4; ++k) { w.lambda[k] = lambda4[4 * i + k]; w.pdf[k] = 1.0f; }
            Spec s = eval_spectrum_texture(d, tex, c, w);
            for (int k = 0; k < 4; ++k) out4[4 * i + k] = s.v[k];
        }
    }
}
// ---- Image::generate_pyramid (image.rs:699-787) with Image::float_resize_up (:1007-1111) / resample_weights (:1113-1141) ----
// `image`: width x height x n_channels linear f32 texels (what convert_to_format(Float) yields).  Levels are written back to back
// into `out`; returns the number of levels.  As written: resample_weights evaluates the windowed sinc at `first_pixel + 0.5` for
// all four taps (pbrt: first_pixel + j + 0.5), so after normalisation every tap weighs ~0.25.
static inline Float orc_sin_over_x(Float x) { if (1.0f - x * x == 1.0f) return 1.0f; return std::sin(x) / x; }      // math.rs:413-420
static inline Float orc_sinc(Float x) { return orc_sin_over_x(PI_F * x); }
static inline Float orc_windowed_sinc(Float x, Float radius, Float tau) { if (std::fabs(x) > radius) return 0.0f; return orc_sinc(x) * orc_sinc(x / tau); }
struct OrcResampleWeight { int32_t first_pixel; Float weight[4]; };
static std::vector<OrcResampleWeight> orc_resample_weights(int old_res, int new_res) {
    std::vector<OrcResampleWeight> wt(new_res);
    const Float filter_radius = 2.0f, tau = 2.0f;
    for (int i = 0; i < new_res; ++i) {
        const Float center = ((Float)i + 0.5f) * (Float)old_res / (Float)new_res;
        wt[i].first_pixel = std::max(0, f2i(std::floor(center - filter_radius + 0.5f)));
        for (int j = 0; j < 4; ++j) { const Float pos = (Float)wt[i].first_pixel + 0.5f; wt[i].weight[j] = orc_windowed_sinc(pos - center, filter_radius, tau); }
        const Float inv = 1.0f / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
        for (int j = 0; j < 4; ++j) wt[i].weight[j] *= inv;
    }
    return wt;
}
static inline uint32_t orc_next_pow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; }
int32_t orc_image_pyramid_levels(int32_t width, int32_t height, int32_t* res_xy /* 2 per level, may be null */) {
    uint32_t w = (uint32_t)width, h = (uint32_t)height;
    if ((w & (w - 1)) || (h & (h - 1))) { w = orc_next_pow2(w); h = orc_next_pow2(h); }
    const int32_t n_levels = 1 + (int32_t)std::log2((Float)std::max(w, h));
    int32_t rx = (int32_t)w, ry = (int32_t)h;
    for (int32_t l = 0; l < n_levels; ++l) {
        if (res_xy) { res_xy[2 * l] = rx; res_xy[2 * l + 1] = ry; }
        rx = std::max(1, (rx + 1) / 2); ry = std::max(1, (ry + 1) / 2);
    }
    return n_levels;
}
int32_t orc_image_generate_pyramid(const float* image, int32_t width, int32_t height, int32_t nc, int32_t wrap, float* out) {
    std::vector<float> cur(image, image + (size_t)width * height * nc);
    int32_t rx = width, ry = height;
    if (((uint32_t)rx & ((uint32_t)rx - 1)) || ((uint32_t)ry & ((uint32_t)ry - 1))) {                    // float_resize_up
        const int32_t nx = (int32_t)orc_next_pow2((uint32_t)rx), ny = (int32_t)orc_next_pow2((uint32_t)ry);
        // the reference asserts new_res > resolution in BOTH dimensions (image.rs:1009-1010)
        const std::vector<OrcResampleWeight> xw = orc_resample_weights(rx, nx), yw = orc_resample_weights(ry, ny);
        auto texel = [&](int32_t x, int32_t y, int c) -> Float {                                          // copy_rect_out + remap_pixel_coords (:134-177)
            int32_t p[2] = {x, y}; const int32_t res[2] = {rx, ry};
            for (int k = 0; k < 2; ++k) {
                if (p[k] >= 0 && p[k] < res[k]) continue;
                if (wrap == SG_WRAP_CLAMP) p[k] = p[k] < 0 ? 0 : res[k] - 1;
                else { int32_t r = p[k] - (p[k] / res[k]) * res[k]; p[k] = r < 0 ? r + res[k] : r; }        // repeat (black panics in the reference)
            }
            return cur[((size_t)p[1] * rx + p[0]) * nc + c];
        };
        std::vector<float> rs((size_t)nx * ny * nc);
        for (int32_t y = 0; y < ny; ++y) for (int32_t x = 0; x < nx; ++x) for (int c = 0; c < nc; ++c) {
            const OrcResampleWeight& wx = xw[x]; const OrcResampleWeight& wy = yw[y];
            Float col[4];
            for (int j = 0; j < 4; ++j) {
                const int32_t yy = wy.first_pixel + j;
                col[j] = wx.weight[0] * texel(wx.first_pixel, yy, c) + wx.weight[1] * texel(wx.first_pixel + 1, yy, c) +
                         wx.weight[2] * texel(wx.first_pixel + 2, yy, c) + wx.weight[3] * texel(wx.first_pixel + 3, yy, c);
            }
            rs[((size_t)y * nx + x) * nc + c] = fmax_(0.0f, wy.weight[0] * col[0] + wy.weight[1] * col[1] + wy.weight[2] * col[2] + wy.weight[3] * col[3]);
        }
        cur.swap(rs); rx = nx; ry = ny;
    }
    const int32_t n_levels = 1 + (int32_t)std::log2((Float)std::max(rx, ry));
    size_t off = 0;
    for (int32_t l = 0; l < n_levels; ++l) {
        std::memcpy(out + off, cur.data(), cur.size() * sizeof(float)); off += cur.size();
        if (l == n_levels - 1) break;
        const int32_t nx = std::max(1, (rx + 1) / 2), ny = std::max(1, (ry + 1) / 2);
        std::vector<float> nxt((size_t)nx * ny * nc);
        const size_t d1 = rx == 1 ? 0 : (size_t)nc, d2 = ry == 1 ? 0 : (size_t)nc * rx;                   // src_deltas :737-753
        for (int32_t y = 0; y < ny; ++y) for (int32_t x = 0; x < nx; ++x) for (int c = 0; c < nc; ++c) {
            const size_t src = ((size_t)(2 * y) * rx + 2 * x) * nc + c;
            nxt[((size_t)y * nx + x) * nc + c] = 0.25f * (cur[src] + cur[src + d1] + cur[src + d2] + cur[src + d1 + d2]);
        }
        cur.swap(nxt); rx = nx; ry = ny;
    }
    return n_levels;
}
// debug: every ray the path integrator traces for one (pixel, sample): 10 floats per ray (see PathCtx::ray_log); returns the count
int64_t orc_path_rays(const SgSceneDesc* desc, const SgRenderParams* rp, int px, int py, int sample, int64_t max_rays, float* out) {
    Scene sc(desc);
    std::vector<float> log;
    PathCtx pc{&sc, rp, nullptr, &log};
    Rng rng; rng.seed_from_u64(stream_key(rp->seed, (uint32_t)(py * desc->film.full_resolution[0] + px), (uint32_t)sample));
    std::vector<SgFilmPixel> film((size_t)(desc->film.pixel_bounds[2] - desc->film.pixel_bounds[0]) * (desc->film.pixel_bounds[3] - desc->film.pixel_bounds[1]));
    std::memset(film.data(), 0, film.size() * sizeof(SgFilmPixel));
    eval_sample(pc, px, py, rng, film.data());
    const int64_t n = std::min<int64_t>((int64_t)log.size() / 10, max_rays);
    std::memcpy(out, log.data(), (size_t)n * 10 * sizeof(float));
    return n;
}
void orc_set_debug_perturb(float rel) { g_debug_perturb = rel; }
void orc_approximate_dp_dxy(const SgSceneDesc* d, const float* p, const float* n, int spp, uint32_t option_flags, float* out6) {
    V3 dpdx, dpdy;
    approximate_dp_dxy(d->camera, v3(p[0], p[1], p[2]), v3(n[0], n[1], n[2]), spp, option_flags, &dpdx, &dpdy);
    out6[0] = dpdx.x; out6[1] = dpdx.y; out6[2] = dpdx.z; out6[3] = dpdy.x; out6[4] = dpdy.y; out6[5] = dpdy.z;
}

void orc_spectrum_sample(const SgSceneDesc* d, int id, const float* lambda4, float* out4) {
    Wavelengths w; for (int i = 0; i < 4; ++i) { w.lambda[i] = lambda4[i]; w.pdf[i] = 1.0f; }
    Spec s = spectrum_sample(d, id, w); for (int i = 0; i < 4; ++i) out4[i] = s.v[i];
}
// dielectric BxDF::sample_f in local space: out = f[4], wi[3], pdf, flags, eta ; returns 0 if None
int orc_dielectric_sample_f(float eta, float ax, float ay, const float* wo, float uc, const float* u2, float* out) {
    BSDF b; b.kind = SG_MATERIAL_DIELECTRIC; b.eta = eta; b.mf = TR::make(ax, ay); b.r = spec_const(0); b.k = spec_const(0);
    BSDFSample bs; V2 u = {u2[0], u2[1]};
    if (!b.sample_local(v3(wo[0], wo[1], wo[2]), uc, u, &bs)) return 0;
    for (int i = 0; i < 4; ++i) out[i] = bs.f.v[i];
    out[4] = bs.wi.x; out[5] = bs.wi.y; out[6] = bs.wi.z; out[7] = bs.pdf; out[8] = (float)bs.flags; out[9] = bs.eta;
    return 1;
}
// generic local-space BxDF evaluation: kind, params(r[4],k[4],eta,ax,ay) ; out f[4], pdf
void orc_bxdf_eval(int kind, const float* prm, const float* wo, const float* wi, float* out) {
    BSDF b; b.kind = kind;
    for (int i = 0; i < 4; ++i) { b.r.v[i] = prm[i]; b.k.v[i] = prm[4 + i]; }
    b.eta = prm[8]; b.mf = TR::make(prm[9], prm[10]);
    Spec f = b.f_local(v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2]));
    for (int i = 0; i < 4; ++i) out[i] = f.v[i];
    out[4] = b.pdf_local(v3(wo[0], wo[1], wo[2]), v3(wi[0], wi[1], wi[2]));
}
int orc_bxdf_sample(int kind, const float* prm, const float* wo, float uc, const float* u2, float* out) {
    BSDF b; b.kind = kind;
    for (int i = 0; i < 4; ++i) { b.r.v[i] = prm[i]; b.k.v[i] = prm[4 + i]; }
    b.eta = prm[8]; b.mf = TR::make(prm[9], prm[10]);
    BSDFSample bs; V2 u = {u2[0], u2[1]};
    if (!b.sample_local(v3(wo[0], wo[1], wo[2]), uc, u, &bs)) return 0;
    for (int i = 0; i < 4; ++i) out[i] = bs.f.v[i];
    out[4] = bs.wi.x; out[5] = bs.wi.y; out[6] = bs.wi.z; out[7] = bs.pdf; out[8] = (float)bs.flags; out[9] = bs.eta;
    return 1;
}
// Triangle::sample / sample_with_context on a free-standing triangle (triangle.rs:773-848 tests)
int orc_tri_intersect(const float* o, const float* d, float t_max, const float* p, float* out4) {
    Ray r; r.o = v3(o[0], o[1], o[2]); r.d = v3(d[0], d[1], d[2]);
    TriHit th;
    if (!intersect_triangle(r, t_max, v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), v3(p[6], p[7], p[8]), &th)) return 0;
    out4[0] = th.b0; out4[1] = th.b1; out4[2] = th.b2; out4[3] = th.t; return 1;
}
int orc_bounds_intersect(const float* bmin, const float* bmax, const float* o, const float* d, float t_max) {
    SgBvhNode n; std::memset(&n, 0, sizeof n);
    for (int a = 0; a < 3; ++a) { n.bmin[a] = bmin[a]; n.bmax[a] = bmax[a]; }
    V3 inv = v3(1.0f / d[0], 1.0f / d[1], 1.0f / d[2]);
    int neg[3] = {inv.x < 0.0f, inv.y < 0.0f, inv.z < 0.0f};
    return bounds_intersect_p_cached(n, v3(o[0], o[1], o[2]), t_max, inv, neg) ? 1 : 0;
}
// sample a light from a context: out = l[4], wi[3], pdf, p_light mid[3], n[3]
int orc_light_sample(const SgSceneDesc* desc, int light, const float* ctx_p, const float* ctx_n, const float* ctx_ns,
                     const float* u2, const float* lambda4, float* out) {
    Scene sc(desc);
    LightSampleContext ctx; ctx.pi = p3fi_exact(v3(ctx_p[0], ctx_p[1], ctx_p[2])); ctx.n = v3(ctx_n[0], ctx_n[1], ctx_n[2]); ctx.ns = v3(ctx_ns[0], ctx_ns[1], ctx_ns[2]);
    Wavelengths w; for (int i = 0; i < 4; ++i) { w.lambda[i] = lambda4[i]; w.pdf[i] = 1.0f; }
    LightLiSample ls; V2 u = {u2[0], u2[1]};
    if (!light_sample_li(sc, desc->lights[light], ctx, u, w, &ls)) return 0;
    for (int i = 0; i < 4; ++i) out[i] = ls.l.v[i];
    out[4] = ls.wi.x; out[5] = ls.wi.y; out[6] = ls.wi.z; out[7] = ls.pdf;
    V3 pm = p3fi_mid(ls.p_light); out[8] = pm.x; out[9] = pm.y; out[10] = pm.z;
    out[11] = ls.n_light.x; out[12] = ls.n_light.y; out[13] = ls.n_light.z;
    return 1;
}
float orc_light_pdf(const SgSceneDesc* desc, int light, const float* ctx_p, const float* ctx_n, const float* ctx_ns, const float* wi) {
    Scene sc(desc);
    LightSampleContext ctx; ctx.pi = p3fi_exact(v3(ctx_p[0], ctx_p[1], ctx_p[2])); ctx.n = v3(ctx_n[0], ctx_n[1], ctx_n[2]); ctx.ns = v3(ctx_ns[0], ctx_ns[1], ctx_ns[2]);
    return light_pdf_li(sc, desc->lights[light], ctx, v3(wi[0], wi[1], wi[2]));
}
// allow_incomplete_pdf = false variants (SimplePathIntegrator, integrator.rs:652-656) and Light::le of the infinite lights
int orc_light_sample_complete(const SgSceneDesc* desc, int light, const float* ctx_p, const float* u2, const float* lambda4, float* out) {
    Scene sc(desc);
    LightSampleContext ctx; ctx.pi = p3fi_exact(v3(ctx_p[0], ctx_p[1], ctx_p[2])); ctx.n = v3(0, 0, 0); ctx.ns = v3(0, 0, 0);
    Wavelengths w; for (int i = 0; i < 4; ++i) { w.lambda[i] = lambda4[i]; w.pdf[i] = 1.0f; }
    LightLiSample ls; V2 u = {u2[0], u2[1]};
    if (!light_sample_li(sc, desc->lights[light], ctx, u, w, &ls, false)) return 0;
    for (int i = 0; i < 4; ++i) out[i] = ls.l.v[i];
    out[4] = ls.wi.x; out[5] = ls.wi.y; out[6] = ls.wi.z; out[7] = ls.pdf;
    return 1;
}
float orc_light_pdf_complete(const SgSceneDesc* desc, int light, const float* wi) {
    Scene sc(desc);
    LightSampleContext ctx; ctx.pi = p3fi_exact(v3(0, 0, 0)); ctx.n = v3(0, 0, 0); ctx.ns = v3(0, 0, 0);
    return light_pdf_li(sc, desc->lights[light], ctx, v3(wi[0], wi[1], wi[2]), false);
}
void orc_light_le(const SgSceneDesc* desc, int light, const float* ray_d, const float* lambda4, float* out4) {
    Wavelengths w; for (int i = 0; i < 4; ++i) { w.lambda[i] = lambda4[i]; w.pdf[i] = 1.0f; }
    Spec s = light_le(desc, desc->lights[light], v3(ray_d[0], ray_d[1], ray_d[2]), w);
    for (int i = 0; i < 4; ++i) out4[i] = s.v[i];
}
void orc_equal_area_square_to_sphere(const float* p2, float* out3) {
    V2 p = {p2[0], p2[1]}; V3 w = equal_area_square_to_sphere(p); out3[0] = w.x; out3[1] = w.y; out3[2] = w.z;
}
void orc_equal_area_sphere_to_square(const float* d3, float* out2) {
    V2 p = equal_area_sphere_to_square(v3(d3[0], d3[1], d3[2])); out2[0] = p.x; out2[1] = p.y;
}

}  // extern "C"
