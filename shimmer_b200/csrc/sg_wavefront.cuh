// Wavefront state + kernels of the B200 path tracer.
//
// One wavefront = up to `capacity` camera paths resident in HBM as structure-of-arrays
// (float4-packed so every access is one LDG.128/STG.128).  Per bounce the stream runs
//   trace_closest -> shade_miss / shade<Diffuse|Conductor|Dielectric> -> trace_shadow
// with device-side queues: kernels read their work count from device counters, so a whole
// batch (max_depth+1 bounces) is enqueued without a single host synchronisation.
// Queue appends are warp-aggregated (__ballot_sync + one atomicAdd per warp and queue).
#pragma once
#include <cuda_fp16.h>
#include "sg_shading.cuh"
#include "sg_texture.cuh"
#include "sg_envmap.cuh"
#include "sg_trace2.cuh"
#include "sg_sphere_surface.cuh"
#include "sg_patch_light.cuh"

namespace sg {

struct PathState {
    float4* ray_o;      // o.xyz
    float4* ray_d;      // d.xyz
    float4* hit_b;      // b0 b1 b2 t
    int*    hit_prim;
    int*    hit_inst;   // instance index of the hit or -1; allocated only for scenes with object instances
    uint32_t* mat_override;  // material chosen by MixMaterial::choose_material for the current hit; allocated only for scenes with Mix materials
    float4* L;
    float4* beta;
    float4* lambda;
    float4* lpdf;
    ulonglong2* rng_a;  // xoshiro256++ s0 s1
    ulonglong2* rng_b;  //               s2 s3
    uint32_t* pixel;    // linear index inside the film window
    uint32_t* flags;    // [0:8) depth, bit 8 specular_bounce, bit 9 any_non_specular_bounces
    float2* pb_eta;     // p_b, eta_scale
    float4* ctx0;       // prev LightSampleContext: pi.lo.xyz, pi.hi.x
    float4* ctx1;       //   pi.hi.yz, n.xy
    float4* ctx2;       //   n.z, ns.xyz
    float4* sh_o;       // shadow ray origin
    float4* sh_d;       // shadow ray direction (unnormalised p_to - p_from)
    float4* sh_L;       // beta * Ld, added to L if the shadow ray is unoccluded
    // auxiliary (differential) rays, allocated only for scenes with image textures: flags bit 10 = present
    float4* aux0;       // rx_origin.xyz, rx_direction.x
    float4* aux1;       // rx_direction.yz, ry_origin.xy
    float4* aux2;       // ry_origin.z, ry_direction.xyz
    // staged shading of textured scenes (k_shade STAGE 1 -> 2): the BSDF record of the current hit, allocated only then
    float4* rec[6];     // pi.lo.xyz pi.hi.x | pi.hi.yz n.xy | n.z sn.xyz | fx.xyz wo_si.x | wo_si.yz - - | reflectance
};
static constexpr int kPathBytes = 16 * 16 + 4 + 4 + 4 + 8;   // per-path HBM footprint (276 B; +48 B with image textures)
static constexpr uint32_t kFlagSpecular = 256u, kFlagNonSpecular = 512u, kFlagAux = 1024u;
SGD AuxRays aux_load(const PathState& st, uint32_t path) {
    const float4 a = st.aux0[path], b = st.aux1[path], c = st.aux2[path];
    AuxRays r; r.has = true;
    r.rxo = f3(a.x, a.y, a.z); r.rxd = f3(a.w, b.x, b.y); r.ryo = f3(b.z, b.w, c.x); r.ryd = f3(c.y, c.z, c.w);
    return r;
}
SGD void aux_store(const PathState& st, uint32_t path, const AuxRays& r) {
    st.aux0[path] = make_float4(r.rxo.x, r.rxo.y, r.rxo.z, r.rxd.x);
    st.aux1[path] = make_float4(r.rxd.y, r.rxd.z, r.ryo.x, r.ryo.y);
    st.aux2[path] = make_float4(r.ryo.z, r.ryd.x, r.ryd.y, r.ryd.z);
}

enum { Q_MISS = 0, Q_DIFFUSE = 1, Q_CONDUCTOR = 2, Q_DIELECTRIC = 3, Q_COATED = 4, Q_THIN = 5, Q_COATED_CONDUCTOR = 6, Q_MIX = 7, Q_NKINDS = 8 };
// per-depth counter block (uint32 x 16)
enum { C_NRAY = 0, C_NSHADE = 1 /*..8*/, C_NSHADOW = 9, C_CUR_CLOSEST = 10, C_CUR_SHADOW = 11, C_STRIDE = 16 };

struct Queues {
    uint32_t* ray[2];
    uint32_t* shade[Q_NKINDS];
    uint32_t* shadow;
    uint32_t* counters;     // (max_depth + 2) * C_STRIDE
    uint32_t* sorted;       // textured scenes: one shade queue re-ordered by material (k_sort_queue_*), else nullptr
    uint32_t* sort_hist;    // 2 x 64 words: bucket counts, bucket cursors
    // order-preserving shade queues (k_queue_count / _scan / _scatter): the closest-hit kernel records the destination queue of
    // ray-queue entry i in ray_kind[i] instead of appending; nullptr = append with atomics at retire time
    uint8_t*  ray_kind;     // capacity bytes (0xff = none)
    uint32_t* block_counts; // ceil(capacity / 2048) x Q_NKINDS words: per-block counts, turned into exclusive bases by k_queue_scan
};

struct RenderConst {
    uint64_t seed;
    int32_t  sample_begin;
    int32_t  spp;
    int32_t  max_depth;
    int32_t  regularize;
    uint32_t option_flags;
    int32_t  win_x0, win_y0, win_w, win_h;
    int32_t  full_res_x;
    int32_t  n_samples;          // sample_end - sample_begin of this call
    int32_t  path_order;         // 0 = sample-major row-major (debug), 1 = pixel-major tiled
    int32_t  integrator;         // SgIntegratorKind
    int32_t  integrator_flags;   // SG_SIMPLEPATH_*
    int32_t  shade_sync;         // shade kernels: bit mask of the stage barriers that keep a CTA's warps in the same code region
    int32_t  shade_sync_tex;     //   (untextured / textured variants)
};

// Path order of a wavefront.  Wavefront slot g -> (pixel, sample): pixel-major, so the samples of one pixel
// sit in adjacent lanes (identical origin, near-identical direction at depth 0, neighbouring hit points at
// every later depth), and pixels follow 8x4 tiles inside 32x32 blocks so that the ~100k paths resident on
// the GPU at any moment cover a compact screen region.  Edge tiles are clipped; the map is a bijection for
// any window size.  The reference's order is rayon's tile schedule (integrator.rs:235-263), i.e. arbitrary.
SGD void pixel_from_order(uint32_t ord, int w, int h, int& x, int& y) {
    const uint32_t W = (uint32_t)w, H = (uint32_t)h;
    const uint32_t by = ord / (W * 32u); uint32_t r = ord - by * W * 32u;
    const uint32_t bh = min(32u, H - by * 32u);
    if (bh == 32u && (r >> 10) < (W >> 5)) {                // a full 32x32 block (all but the right / bottom edge): shifts only
        const uint32_t bx = r >> 10; r &= 1023u;
        x = (int)(bx * 32u + ((r >> 5) & 3u) * 8u + (r & 7u)); y = (int)(by * 32u + (r >> 7) * 4u + ((r >> 3) & 3u));
        return;
    }
    const uint32_t bx = r / (32u * bh); r -= bx * 32u * bh;
    const uint32_t bw = min(32u, W - bx * 32u);
    const uint32_t ty = r / (bw * 4u); r -= ty * bw * 4u;
    const uint32_t th = min(4u, bh - ty * 4u);
    const uint32_t tx = r / (8u * th); r -= tx * 8u * th;
    const uint32_t tw = min(8u, bw - tx * 8u);
    const uint32_t yi = r / tw, xi = r - yi * tw;
    x = (int)(bx * 32u + tx * 8u + xi); y = (int)(by * 32u + ty * 4u + yi);
}

struct DevStats { unsigned long long closest, shadow, nodes, tris, nodes_closest, tris_closest; };

#define SG_SHADOW_TMAX 0.9999f        /* 1.0 - SHADOW_EPISLON, integrator.rs:66,115 */
#ifndef SG_TRACE_THREADS
#define SG_TRACE_THREADS 128
#endif
static constexpr int kTraceThreads = SG_TRACE_THREADS;
// Resident CTAs per SM the traversal kernels are compiled for: 9 x 128 threads caps ptxas at 56 registers, which is
// what the closest-hit kernel needs without spilling and what the shared-memory stack (20 levels) leaves room for;
// the instanced variants carry more lane state and are given 8 (64 registers).
#ifndef SG_TRACE_MIN_BLOCKS
#define SG_TRACE_MIN_BLOCKS 9
#endif
// 1: postponed-leaf scheduling for the triangle-only traversal kernels (trace_persistent_post); 0: in-order loop everywhere.
// Measured on C2 (B200, round 1): bit-identical hits, but 2415 vs 2573 Mrays/s closest-hit and 2043 vs 2165 Mrays/s any-hit --
// the extra boxes visited under a stale t_max and the heavier lane state cost more than the fuller triangle phases save --
// so the in-order loop stays the default; the variant is kept for A/B runs (-DSG_TRACE_POSTPONE=1).
#ifndef SG_TRACE_POSTPONE
#define SG_TRACE_POSTPONE 0
#endif
// 7 blocks (72 registers) for the instanced variants: 836 vs 843 Mrays/s closest-hit on C4 -- no gain over 8, kept at 8.
#ifndef SG_TRACE_MIN_BLOCKS_INST
#define SG_TRACE_MIN_BLOCKS_INST 8
#endif

// ---- camera ray generation: evaluate_pixel_sample integrator.rs:326-362 ----
static __global__ void __launch_bounds__(256) k_generate(const __grid_constant__ DScene sc, PathState st, Queues q, RenderConst rc,
                                                  unsigned long long first_item, uint32_t first_ord, uint32_t first_rem, uint32_t count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) q.counters[C_NRAY] = count;
    if (i >= count) return;
    const unsigned long long g = first_item + i;
    int s, wx, wy;
    if (rc.path_order == 0) {
        const unsigned long long npix = (unsigned long long)rc.win_w * (unsigned long long)rc.win_h;
        s = rc.sample_begin + (int)(g / npix);
        const uint32_t lin = (uint32_t)(g % npix);
        wx = (int)(lin % (uint32_t)rc.win_w); wy = (int)(lin / (uint32_t)rc.win_w);
    } else {
        // g / n_samples and g % n_samples without 64-bit division: the host split the batch's first item, the rest fits 32 bits
        const uint32_t t = first_rem + i, dq = t / (uint32_t)rc.n_samples;
        s = rc.sample_begin + (int)(t - dq * (uint32_t)rc.n_samples);
        pixel_from_order(first_ord + dq, rc.win_w, rc.win_h, wx, wy);
    }
    const uint32_t pix = (uint32_t)wy * (uint32_t)rc.win_w + (uint32_t)wx;
    const int px = rc.win_x0 + wx, py = rc.win_y0 + wy;
    Rng rng; rng.seed_from_u64(stream_key(rc.seed, (uint32_t)(py * rc.full_res_x + px), (uint32_t)s));
    Wavelengths lam; float3 o, d; float w;
    uint32_t flags0 = 0u;
    if (st.aux0 != nullptr) {                        // scene has image textures: camera ray differentials (camera.rs:1036-1079)
        AuxRays aux;
        camera_stage(sc, rc.option_flags, px, py, rng, lam, o, d, w, &aux);
        if (!(rc.option_flags & SG_OPT_DISABLE_PIXEL_JITTER)) {                    // integrator.rs:355-361, ray.rs:137-145
            const float sc_ = fmaxf(0.125f, 1.0f / sqrtf((float)rc.spp));
            aux.rxo = o + (aux.rxo - o) * sc_; aux.ryo = o + (aux.ryo - o) * sc_;
            aux.rxd = d + (aux.rxd - d) * sc_; aux.ryd = d + (aux.ryd - d) * sc_;
        }
        aux_store(st, i, aux);
        flags0 = kFlagAux;
    } else camera_stage(sc, rc.option_flags, px, py, rng, lam, o, d, w);
    // L = 0, beta = 1, p_b = eta_scale = 1 and the (unused) previous light-sample context are NOT written: at wavefront depth d
    // every queued path has depth d (one bounce per iteration), so the depth-0 kernels take those values as constants
    // (kernel argument `depth` == 0) and the first shade kernel stores them.  A path that escapes at depth 0 gets its L = 0
    // from the closest-hit kernel's retire step.  108 instead of 196 bytes per camera sample.
    st.ray_o[i] = make_float4(o.x, o.y, o.z, 0.0f);
    st.ray_d[i] = make_float4(d.x, d.y, d.z, 0.0f);
    st.lambda[i] = lam.lambda; st.lpdf[i] = lam.pdf;
    st.rng_a[i] = make_ulonglong2(rng.s0, rng.s1); st.rng_b[i] = make_ulonglong2(rng.s2, rng.s3);
    st.pixel[i] = pix;
    st.flags[i] = flags0;
    q.ray[0][i] = i;
}

// ---- closest-hit / any-hit traversal over a ray queue ----
// Persistent warps (grid = SMs x resident CTAs).  See sg_trace2.cuh for the traversal core.
// `IO` supplies rays and consumes results; retire() is called warp-convergently so it may use
// warp collectives (ballot-compacted appends into the material queues).
// Scheduling thresholds (lanes): run the triangle phase once this many lanes hold a leaf, and
// the retire/refill phase once this many lanes wait for a ray; kInteriorBurst interior steps
// are run back to back between votes.
template <bool ANY, bool COUNT, bool INST, class IO, class CursorT>
SGD void trace_persistent(const TraceScene& ts, IO& io, CursorT n, CursorT* cursor, uint32_t* s_mem,
                          uint32_t& cnt_nodes, uint32_t& cnt_tris, bool first_depth = false) {
    const int lane = threadIdx.x & 31;
    const int refill_threshold = first_depth ? ts.refill_threshold_d0 : ts.refill_threshold;
    const int interior_burst = first_depth ? ts.interior_burst_d0 : ts.interior_burst;
    uint2 spill[kSpillLevels];
    Stack S;
    S.stride = blockDim.x; S.levels = ts.smem_levels; S.spill = spill;
    S.template bind<ANY>(s_mem, (int)threadIdx.x);
    // INST kernels: the parked render-space ray (lane_save_ray) sits behind the stack levels (closest-hit: refs + entry distances, any-hit: refs)
    S.s_save = reinterpret_cast<float*>(s_mem + (size_t)ts.smem_levels * blockDim.x * (ANY ? 1 : 2)) + threadIdx.x;
    Lane L; L.cur = kEmptyRef; L.sp = 0; L.hit.prim = -1; L.hit.inst = -1; L.inst = -1; L.sp_base = 0; L.t_saved = 0.0f; L.inst_hit = false;
    bool has_ray = false, dead = false, finished = false;
    uint32_t m_dead = 0u;                                                        // lanes that found the queue exhausted (changes in the refill phase only)
    CursorT idx = 0;
    for (;;) {
        const bool is_leaf = has_ray && (L.cur & kLeafBit) != 0;
        const bool is_int = has_ray && !is_leaf;
        const uint32_t m_int = __ballot_sync(0xffffffffu, is_int);
        const uint32_t m_leaf = __ballot_sync(0xffffffffu, is_leaf);
        // lanes waiting for a ray: finished, or never had one.  Every lane is in exactly one state, so the triangle-only kernels
        // derive the mask from the other three (one vote less per iteration); the instanced kernels sit at their register cap and
        // spill with one more live value, they keep the vote
        uint32_t m_wait;
        if constexpr (INST) m_wait = __ballot_sync(0xffffffffu, finished || (!has_ray && !dead));
        else m_wait = ~(m_int | m_leaf | m_dead);
        if ((m_int | m_leaf) == 0u || __popc(m_wait) >= refill_threshold) {
            // ---- retire finished rays, claim new ones (warp-aggregated) ----
            if constexpr (!INST) L.hit.inst = -1;
            if (__ballot_sync(0xffffffffu, finished)) io.retire(finished, idx, L.hit, lane);
            finished = false;
            const bool need = !has_ray && !dead;
            const uint32_t mask = __ballot_sync(0xffffffffu, need);
            if (mask) {
                CursorT base = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) base = atomicAdd(cursor, (CursorT)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    idx = base + (CursorT)__popc(mask & ((1u << lane) - 1u));
                    if (idx < n) {
                        float3 o, d; float tmax;
                        io.load(idx, o, d, tmax);
                        lane_begin<ANY>(ts, L, o, d, tmax, cnt_nodes, COUNT);
                        if constexpr (INST) { L.hit.inst = -1; L.inst = -1; L.sp_base = 0; L.inst_hit = false; }
                        if (L.cur == kEmptyRef) finished = true; else has_ray = true;
                    } else dead = true;
                }
            }
            if constexpr (!INST) m_dead = __ballot_sync(0xffffffffu, dead);
            if (!__ballot_sync(0xffffffffu, has_ray || finished)) break;
            continue;
        }
        if (__popc(m_leaf) >= ts.leaf_threshold || m_int == 0u) {
            // ---- triangle phase ----
            if (is_leaf) {
                lane_step_leaf<ANY, COUNT, INST>(ts, L, S, cnt_nodes, cnt_tris, io, idx);
                if (L.cur == kEmptyRef) { finished = true; has_ray = false; }
            }
            continue;
        }
        // ---- interior phase ----
        if (is_int) {
#pragma unroll 1
            for (int k = 0; k < interior_burst; ++k) {
                lane_step_interior<ANY, COUNT, INST>(ts, L, S, cnt_nodes, io, idx);
                if constexpr (INST) {                                            // (register allocation of the instanced kernels prefers this form)
                    if (L.cur == kEmptyRef) { finished = true; has_ray = false; break; }
                    if (L.cur & kLeafBit) break;
                } else if (L.cur >= kEmptyRef) break;                            // one test for "finished" (== kEmptyRef) and "at a leaf" (bit 31)
            }
            if constexpr (!INST) if (L.cur == kEmptyRef) { finished = true; has_ray = false; }
        }
    }
}

// Postponed-leaf scheduling (see sg_trace2.cuh): lanes park their first leaf and keep traversing; the triangle phase runs
// for EVERY lane that holds a parked leaf once enough lanes are blocked on a second one (or have nothing else to do).
template <bool ANY, class IO, class CursorT>
SGD void trace_persistent_post(const TraceScene& ts, IO& io, CursorT n, CursorT* cursor, uint32_t* s_mem) {
    const int lane = threadIdx.x & 31;
    uint2 spill[kSpillLevels];
    Stack S;
    S.stride = blockDim.x; S.levels = ts.smem_levels; S.spill = spill;
    S.template bind<ANY>(s_mem, (int)threadIdx.x);
    Lane L; L.cur = kEmptyRef; L.sp = 0; L.hit.prim = -1; L.hit.inst = -1; L.inst = -1; L.sp_base = 0; L.t_saved = 0.0f; L.inst_hit = false;
    uint32_t pend = kEmptyRef; float pend_t = 0.0f;
    bool has_ray = false, dead = false, finished = false;
    CursorT idx = 0;
    for (;;) {
        const bool is_int = has_ray && L.cur < kBlockedRef;                       // an interior node to expand
        const bool is_blk = has_ray && !is_int;                                   // blocked on a 2nd leaf, or stack dry with a parked leaf
        const uint32_t m_int = __ballot_sync(0xffffffffu, is_int);
        const uint32_t m_blk = __ballot_sync(0xffffffffu, is_blk);
        const uint32_t m_wait = __ballot_sync(0xffffffffu, finished || (!has_ray && !dead));
        if ((m_int | m_blk) == 0u || __popc(m_wait) >= ts.refill_threshold) {
            if (__ballot_sync(0xffffffffu, finished)) io.retire(finished, idx, L.hit, lane);
            finished = false;
            const bool need = !has_ray && !dead;
            const uint32_t mask = __ballot_sync(0xffffffffu, need);
            if (mask) {
                CursorT base = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) base = atomicAdd(cursor, (CursorT)__popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need) {
                    idx = base + (CursorT)__popc(mask & ((1u << lane) - 1u));
                    if (idx < n) {
                        float3 o, d; float tmax;
                        io.load(idx, o, d, tmax);
                        uint32_t dummy = 0;
                        lane_begin<ANY>(ts, L, o, d, tmax, dummy, false);
                        pend = kEmptyRef;
                        if (L.cur == kEmptyRef) finished = true;
                        else {
                            has_ray = true;
                            if (L.cur & kLeafBit) lane_settle<ANY>(L, S, L.cur, -INFINITY, pend, pend_t);   // single-leaf tree
                        }
                    } else dead = true;
                }
            }
            if (!__ballot_sync(0xffffffffu, has_ray || finished)) break;
            continue;
        }
        if (__popc(m_blk) >= ts.leaf_threshold || m_int == 0u) {
            // ---- triangle phase: every parked leaf ----
            if (has_ray && pend != kEmptyRef) {
                const bool done = lane_test_pending<ANY>(ts, L, pend, pend_t);
                pend = kEmptyRef;
                if (ANY && done) L.cur = kEmptyRef;
                else if (L.cur == kBlockedRef) { float t; const uint32_t r = lane_pop_t<ANY>(L, S, t); lane_settle<ANY>(L, S, r, t, pend, pend_t); }
                if (L.cur == kEmptyRef && pend == kEmptyRef) { finished = true; has_ray = false; }
            }
            continue;
        }
        // ---- interior phase ----
        if (is_int) {
#pragma unroll 1
            for (int k = 0; k < ts.interior_burst; ++k) {
                lane_step_interior_post<ANY>(ts, L, S, pend, pend_t);
                if (L.cur >= kBlockedRef) break;
            }
            if (L.cur == kEmptyRef && pend == kEmptyRef) { finished = true; has_ray = false; }
        }
    }
}

struct ClosestIO {
    const DScene& sc; PathState st; Queues q; uint32_t* C; const uint32_t* queue; uint32_t queue_mask;
    bool first_depth;           // depth 0: nobody has written L yet (k_generate leaves it out) -- escaping paths get L = 0 here
    uint32_t path;
    SGD void load(uint32_t i, float3& o, float3& d, float& tmax) {
        path = first_depth ? i : queue[i];                      // the depth-0 ray queue is the identity (k_generate): one dependent load less
        const float4 o4 = st.ray_o[path], d4 = st.ray_d[path];
        o = f3(o4.x, o4.y, o4.z); d = f3(d4.x, d4.y, d4.z); tmax = INFINITY;
    }
    SGD void rebind(uint32_t i) { path = first_depth ? i : queue[i]; }           // two-rays-per-lane traversal: the slot being retired
    SGD void retire(bool fin, uint32_t idx, const HitRec& hit, int lane) {
        int kind = -1;
        if (fin) {
            st.hit_prim[path] = hit.prim;
            if (st.hit_inst) st.hit_inst[path] = hit.inst;
            if (hit.prim >= 0) {
                st.hit_b[path] = make_float4(hit.b0, hit.b1, hit.b2, hit.t);
                kind = 1 + (int)((__float_as_uint(__ldg(sc.tri_verts + 3 * (size_t)hit.prim).w) >> 28) & 7u);
            } else {
                kind = Q_MISS;
                if (first_depth) st.L[path] = spec1(0.0f);
            }
        }
        if (q.ray_kind != nullptr) {                                        // order-preserving queues: just note where entry idx goes
            if (fin) q.ray_kind[idx] = (uint8_t)kind;
            return;
        }
#pragma unroll
        for (int k = 0; k < Q_NKINDS; ++k) {
            if (!((queue_mask >> k) & 1u)) continue;                         // warp-uniform: material kinds the scene does not use cost no vote
            const uint32_t mask = __ballot_sync(0xffffffffu, kind == k);
            if (mask) {
                uint32_t qbase = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) qbase = atomicAdd(C + C_NSHADE + k, (uint32_t)__popc(mask));
                qbase = __shfl_sync(0xffffffffu, qbase, leader);
                if (kind == k) q.shade[k][qbase + __popc(mask & ((1u << lane) - 1u))] = path;
            }
        }
    }
};
struct ShadowIO {
    PathState st; const uint32_t* queue;
    uint32_t path;
    SGD void load(uint32_t i, float3& o, float3& d, float& tmax) {
        path = queue[i];
        const float4 o4 = st.sh_o[path], d4 = st.sh_d[path];
        o = f3(o4.x, o4.y, o4.z); d = f3(d4.x, d4.y, d4.z); tmax = SG_SHADOW_TMAX;
    }
    SGD void rebind(uint32_t i) { path = queue[i]; }
    SGD void retire(bool fin, uint32_t, const HitRec& hit, int) {
        if (fin && hit.prim < 0) st.L[path] = st.L[path] + st.sh_L[path];      // unoccluded: L += beta * Ld
    }
};

// ---- two rays per lane (triangle-only scenes, A/B: SG_TRACE_DUAL=1) ----
// The voted phase loop leaves lanes idle: a lane that reaches a leaf waits until enough lanes hold one, a finished lane waits for the
// next refill (interior phases run at ~19 of 32 lanes, triangle phases at ~6).  Here every lane owns TWO rays -- one in registers,
// one parked in shared memory with its own traversal stack -- and switches to the parked ray whenever the active one cannot take an
// interior step (it holds a leaf, is finished, or the slot is empty).  Each ray is still traversed exactly as in trace_persistent
// (same steps in the same order), so hits are bit-identical; only the interleaving changes.
static constexpr int kSpillDual = 64;
static constexpr int kParkWords = 19;
template <bool ANY, class IO>
SGD void trace_persistent_dual(const TraceScene& ts, IO& io, uint32_t n, uint32_t* cursor, uint32_t* s_mem, int levels, bool first_depth) {
    const int lane = threadIdx.x & 31;
    const int stride = blockDim.x;
    const int refill_threshold = first_depth ? 2 * ts.refill_threshold_d0 : ts.refill_threshold_dual;
    const int interior_burst = first_depth ? ts.interior_burst_d0 : ts.interior_burst;
    uint2 spill0[kSpillDual], spill1[kSpillDual];
    const int stack_words = levels * stride * (ANY ? 1 : 2);
    uint32_t* const base0 = s_mem;
    uint32_t* const base1 = s_mem + stack_words;
    float* const park = reinterpret_cast<float*>(s_mem + 2 * stack_words) + threadIdx.x;
    Stack S; S.stride = stride; S.levels = levels; S.s_save = nullptr;
    S.template bind<ANY>(base0, (int)threadIdx.x); S.spill = spill0;
    int which = 0;
    Lane L; L.cur = kEmptyRef; L.sp = 0; L.hit.prim = -1; L.hit.inst = -1; L.inst = -1; L.sp_base = 0; L.t_saved = 0.0f; L.inst_hit = false;
    L.o = f3(0.0f, 0.0f, 0.0f); L.inv_dir = f3(1.0f, 1.0f, 1.0f); L.rp.kx = 0; L.rp.ky = 1; L.rp.kz = 2; L.rp.sx = L.rp.sy = L.rp.sz = 0.0f;
    L.t_max = 0.0f; L.nx = L.ny = L.nz = 0; L.hit.t = L.hit.b0 = L.hit.b1 = L.hit.b2 = 0.0f;
    uint32_t idx = 0;
    // slot states: 0 empty (wants a ray), 1 interior node next, 2 holds a leaf, 3 finished (to retire), 4 dead (queue exhausted)
    int sa = 0, sb = 0;
    uint32_t dummy_n = 0, dummy_t = 0;
    auto swap_slots = [&]() {
        float w[kParkWords];
#pragma unroll
        for (int k = 0; k < kParkWords; ++k) w[k] = park[k * stride];
        park[0 * stride] = L.o.x; park[1 * stride] = L.o.y; park[2 * stride] = L.o.z;
        park[3 * stride] = L.inv_dir.x; park[4 * stride] = L.inv_dir.y; park[5 * stride] = L.inv_dir.z;
        park[6 * stride] = __int_as_float(L.rp.kx | (L.rp.ky << 2) | (L.rp.kz << 4));
        park[7 * stride] = L.rp.sx; park[8 * stride] = L.rp.sy; park[9 * stride] = L.rp.sz;
        park[10 * stride] = L.t_max; park[11 * stride] = __uint_as_float(L.cur); park[12 * stride] = __int_as_float(L.sp);
        park[13 * stride] = __int_as_float(L.hit.prim); park[14 * stride] = L.hit.t;
        park[15 * stride] = L.hit.b0; park[16 * stride] = L.hit.b1; park[17 * stride] = L.hit.b2;
        park[18 * stride] = __uint_as_float(idx);
        L.o = f3(w[0], w[1], w[2]); L.inv_dir = f3(w[3], w[4], w[5]);
        const int kk = __float_as_int(w[6]); L.rp.kx = kk & 3; L.rp.ky = (kk >> 2) & 3; L.rp.kz = (kk >> 4) & 3;
        L.rp.sx = w[7]; L.rp.sy = w[8]; L.rp.sz = w[9];
        L.t_max = w[10]; L.cur = __float_as_uint(w[11]); L.sp = __float_as_int(w[12]);
        L.hit.prim = __float_as_int(w[13]); L.hit.t = w[14]; L.hit.b0 = w[15]; L.hit.b1 = w[16]; L.hit.b2 = w[17];
        idx = __float_as_uint(w[18]);
        L.nx = L.inv_dir.x < 0.0f; L.ny = L.inv_dir.y < 0.0f; L.nz = L.inv_dir.z < 0.0f;
        const int t = sa; sa = sb; sb = t;
        which ^= 1;
        uint32_t* const b = which ? base1 : base0;
        S.template bind<ANY>(b, (int)threadIdx.x); S.spill = which ? spill1 : spill0;
    };
    auto state_of = [&]() { return L.cur == kEmptyRef ? 3 : ((L.cur & kLeafBit) ? 2 : 1); };
    for (;;) {
        if (sa != 1 && sb == 1) swap_slots();                                   // keep an interior ray active whenever the lane has one
        const uint32_t m_int = __ballot_sync(0xffffffffu, sa == 1);
        const uint32_t m_leaf_a = __ballot_sync(0xffffffffu, sa == 2), m_leaf_b = __ballot_sync(0xffffffffu, sb == 2);
        const uint32_t m_need_a = __ballot_sync(0xffffffffu, sa == 0 || sa == 3), m_need_b = __ballot_sync(0xffffffffu, sb == 0 || sb == 3);
        if ((m_int | m_leaf_a | m_leaf_b) == 0u || __popc(m_need_a) + __popc(m_need_b) >= refill_threshold) {
            // ---- retire finished rays, claim new ones: one slot per lane and round (the slot to serve is made the active one) ----
#pragma unroll 1
            for (int round = 0; round < 2; ++round) {
                if ((sb == 0 || sb == 3) && !(sa == 0 || sa == 3)) swap_slots();     // (also when the active slot is dead and a finished ray is parked)
                const bool fin = sa == 3;
                if (__ballot_sync(0xffffffffu, fin)) {
                    if (fin) io.rebind(idx);
                    L.hit.inst = -1;
                    io.retire(fin, idx, L.hit, lane);
                }
                if (fin) sa = 0;
                const bool need = sa == 0;
                const uint32_t mask = __ballot_sync(0xffffffffu, need);
                if (mask) {
                    uint32_t base = 0;
                    const int leader = __ffs(mask) - 1;
                    if (lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(mask));
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (need) {
                        idx = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                        if (idx < n) {
                            float3 o, d; float tmax;
                            io.load(idx, o, d, tmax);
                            lane_begin<ANY>(ts, L, o, d, tmax, dummy_n, false);
                            sa = state_of();
                        } else sa = 4;
                    }
                }
                if (!__ballot_sync(0xffffffffu, sb == 0 || sb == 3)) break;
            }
            if (!__ballot_sync(0xffffffffu, (sa >= 1 && sa <= 3) || (sb >= 1 && sb <= 3))) break;
            continue;
        }
        if (__popc(m_leaf_a) + __popc(m_leaf_b) >= ts.leaf_threshold_dual || m_int == 0u) {
            // ---- triangle phase: every lane that holds a leaf in either slot tests one ----
            if (sa != 2 && sb == 2) swap_slots();
            if (sa == 2) {
                lane_step_leaf<ANY, false, false>(ts, L, S, dummy_n, dummy_t, io, idx);
                sa = state_of();
            }
            continue;
        }
        // ---- interior phase ----
        if (sa == 1) {
#pragma unroll 1
            for (int k = 0; k < interior_burst; ++k) {
                lane_step_interior<ANY, false, false>(ts, L, S, dummy_n, io, idx);
                if (L.cur == kEmptyRef) { sa = 3; break; }
                if (L.cur & kLeafBit) { sa = 2; break; }
            }
        }
    }
}

template <bool ANY, bool COUNT, bool INST>
__global__ void __launch_bounds__(kTraceThreads, INST ? SG_TRACE_MIN_BLOCKS_INST : SG_TRACE_MIN_BLOCKS) k_trace(const __grid_constant__ DScene sc, const __grid_constant__ TraceScene ts,
                                                         PathState st, Queues q, int depth, DevStats* stats) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t* C = q.counters + depth * C_STRIDE;
    uint32_t cnt_nodes = 0, cnt_tris = 0;
    constexpr bool POST = SG_TRACE_POSTPONE && !COUNT && !INST;
    if (ANY) {
        ShadowIO io{st, q.shadow, 0};
        if constexpr (POST) trace_persistent_post<true>(ts, io, C[C_NSHADOW], C + C_CUR_SHADOW, s_mem);
        else trace_persistent<true, COUNT, INST>(ts, io, C[C_NSHADOW], C + C_CUR_SHADOW, s_mem, cnt_nodes, cnt_tris);
    } else {
        ClosestIO io{sc, st, q, C, q.ray[depth & 1], ts.queue_mask, depth == 0, 0};
        if constexpr (POST) trace_persistent_post<false>(ts, io, C[C_NRAY], C + C_CUR_CLOSEST, s_mem);
        else trace_persistent<false, COUNT, INST>(ts, io, C[C_NRAY], C + C_CUR_CLOSEST, s_mem, cnt_nodes, cnt_tris, depth == 0);
    }
    if (COUNT) {
        atomicAdd(&stats->nodes, (unsigned long long)cnt_nodes);
        atomicAdd(&stats->tris, (unsigned long long)cnt_tris);
        if (!ANY) {
            atomicAdd(&stats->nodes_closest, (unsigned long long)cnt_nodes);
            atomicAdd(&stats->tris_closest, (unsigned long long)cnt_tris);
        }
    }
}

#ifndef SG_TRACE_MIN_BLOCKS_DUAL
#define SG_TRACE_MIN_BLOCKS_DUAL 7
#endif
template <bool ANY>
__global__ void __launch_bounds__(kTraceThreads, SG_TRACE_MIN_BLOCKS_DUAL) k_trace_dual(const __grid_constant__ DScene sc, const __grid_constant__ TraceScene ts,
                                                         PathState st, Queues q, int depth, DevStats* stats) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t* C = q.counters + depth * C_STRIDE;
    if (ANY) {
        ShadowIO io{st, q.shadow, 0};
        trace_persistent_dual<true>(ts, io, C[C_NSHADOW], C + C_CUR_SHADOW, s_mem, ts.dual_levels_shadow, false);
    } else {
        ClosestIO io{sc, st, q, C, q.ray[depth & 1], ts.queue_mask, depth == 0, 0};
        trace_persistent_dual<false>(ts, io, C[C_NRAY], C + C_CUR_CLOSEST, s_mem, ts.dual_levels_closest, depth == 0);
    }
}

// ---- escaped rays: infinite lights, integrator.rs:776-794 ----
static __global__ void __launch_bounds__(128) k_shade_miss(const __grid_constant__ DScene sc, PathState st, Queues q, RenderConst rc, int depth) {
    const uint32_t* C = q.counters + depth * C_STRIDE;
    const uint32_t n = C[C_NSHADE + Q_MISS];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t path = q.shade[Q_MISS][i];
        const uint32_t fl = st.flags[path];
        const int pdepth = fl & 0xff; const bool specular_bounce = (fl >> 8) & 1u;
        Wavelengths lam; lam.lambda = st.lambda[path]; lam.pdf = st.lpdf[path];
        Spec L = st.L[path];                                                     // depth 0: zeroed by the closest-hit retire step
        const Spec beta = depth == 0 ? spec1(1.0f) : st.beta[path];              // depth 0: not in memory yet (see k_generate)
        const float p_b = depth == 0 ? 1.0f : st.pb_eta[path].x;
        const float4 rd4 = st.ray_d[path];
        const float3 rd = f3(rd4.x, rd4.y, rd4.z);
        for (int k = 0; k < sc.n_infinite; ++k) {
            const SgLight lt = sc.lights[sc.infinite_ids[k]];
            Spec le = infinite_le(sc, lt, rd, lam);                              // light.rs:792-794, :907-911
            if (rc.integrator != SG_INTEGRATOR_PATH) {                           // SimplePath integrator.rs:613-620 (specular_bounce starts true), RandomWalk :506-512
                if (rc.integrator == SG_INTEGRATOR_RANDOM_WALK || !(rc.integrator_flags & SG_SIMPLEPATH_SAMPLE_LIGHTS) || pdepth == 0 || specular_bounce) L = L + beta * le;
            } else if (pdepth == 0 || specular_bounce) L = L + beta * le;
            else {
                float p_l = (1.0f / (float)sc.n_lights) * infinite_pdf_li(sc, lt, rd, true);   // uniform: 0 (light.rs:770-781); image: compensated pdf (:882-892)
                float w_b = power_heuristic(p_b, p_l);
                L = L + beta * w_b * le;
            }
        }
        st.L[path] = L;
    }
}

// Seed of the LayeredBxDF's private generator: path-stream state x call site (1 = f in sample_ld, 2 = pdf in
// sample_ld, 3 = sample_f, 4 = pdf after sample_f); does not advance the path stream.  Same as the oracle.
SGD uint64_t layer_seed(const Rng& rng, uint64_t site) { return mix64(rng.s0 ^ (site * 0x9e3779b97f4a7c15ULL)); }

// Hit -> SurfaceInteraction for scenes that hold more than top-level triangles: Sphere::intersect (sphere.rs:286-293; hit_b
// carries p_obj), BilinearPatch::intersect (bilinear_patch.rs:496-509; hit_b carries (u, v)), TransformedPrimitive::intersect
// (primitive.rs:155-169) or a plain triangle.
template <bool TEX>
__device__ __noinline__ Surf surface_general(const DScene& sc, const TriGeo& geo, float4 hb, float3 rd, int inst, SurfTex* sx, float3& wo_si) {
    Surf s;
    // inside an instance the shape was intersected with the instance-space ray (primitive.rs:155-169)
    float3 rdi = rd;
    if (inst >= 0 && (geo.mesh & (kSphereBit | kPatchBit))) {
        const float* Mi = sc.instances[inst].mi;
        rdi = f3(Mi[0] * rd.x + Mi[1] * rd.y + Mi[2] * rd.z, Mi[4] * rd.x + Mi[5] * rd.y + Mi[6] * rd.z, Mi[8] * rd.x + Mi[9] * rd.y + Mi[10] * rd.z);
    }
    if (geo.mesh & kSphereBit) {
        const DSphere& S = sc.spheres[geo.mesh & ~kSphereBit];
        s = make_surface_sphere<TEX>(S, f3(hb.x, hb.y, hb.z), sx);
        transform_interaction<TEX>(sc, S.m, S.mi, rdi, s, sx, wo_si);
        if (inst >= 0) { const float3 w1 = wo_si; transform_interaction<TEX>(sc, sc.instances[inst].m, sc.instances[inst].mi, rd, s, sx, wo_si, &w1); }
    } else if (geo.mesh & kPatchBit) {
        s = make_surface_patch<TEX>(sc, geo.mesh & ~kPatchBit, hb.x, hb.y, sx);
        if (inst >= 0) transform_interaction<TEX>(sc, sc.instances[inst].m, sc.instances[inst].mi, rd, s, sx, wo_si);
    } else {
        s = make_surface<TEX>(sc, geo, hb.x, hb.y, hb.z, sx);
        if (inst >= 0) transform_interaction<TEX>(sc, sc.instances[inst].m, sc.instances[inst].mi, rd, s, sx, wo_si);
    }
    return s;
}

// ---- MixMaterial resolution (interaction.rs:206-221, MixMaterial::choose_material material.rs:1309-1330) ----
// Runs between the closest-hit kernel and the shade kernels of a depth: every path whose hit carries a Mix material gets its
// final material (st.mat_override) and is appended to that material kind's shade queue.  The stochastic choice draws from a
// generator seeded by the path stream's state (site 5; the stream itself does not advance) -- see SG_MATERIAL_MIX.
template <bool TEX>
__global__ void __launch_bounds__(128) k_resolve_mix(const __grid_constant__ DScene sc, PathState st, Queues q, RenderConst rc, int depth) {
    uint32_t* C = q.counters + depth * C_STRIDE;
    const uint32_t n = C[C_NSHADE + Q_MIX];
    const int lane = threadIdx.x & 31;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
        int kind = -1; uint32_t path = 0;
        if (i < n) {
            path = q.shade[Q_MIX][i];
            uint32_t material_id; int light_id;
            const TriGeo geo = geo_from_prim(sc, (uint32_t)st.hit_prim[path], material_id, light_id);
            float3 pdp[4];
            TexCoordCtx tc{make_float2(0.0f, 0.0f), 0.0f, 0.0f, 0.0f, 0.0f};
            bool have_ctx = false;
            Rng mr; mr.seed_from_u64(mix64(st.rng_a[path].x ^ (5ull * 0x9e3779b97f4a7c15ULL)));      // layer_seed(rng, 5)
            for (int guard = 0; guard < 64 && sc.materials[material_id].kind == SG_MATERIAL_MIX; ++guard) {
                const SgMaterial mm = sc.materials[material_id];
                float amt = mm.mix_amount;
                if (TEX && mm.tex_mix_amount >= 0) {
                    if (!have_ctx) {                                                               // the hit's TextureEvalContext, as k_shade builds it
                        const float4 rd4 = st.ray_d[path]; const float3 rd = f3(rd4.x, rd4.y, rd4.z);
                        const float4 hb = st.hit_b[path];
                        SurfTex sx; Surf s; float3 wo_si = -rd;
                        if (st.hit_inst == nullptr) s = make_surface<TEX>(sc, geo, hb.x, hb.y, hb.z, &sx);
                        else s = surface_general<TEX>(sc, geo, hb, rd, st.hit_inst[path], &sx, wo_si);
                        AuxRays aux; aux.has = false;
                        if (st.flags[path] & kFlagAux) aux = aux_load(st, path);
                        compute_differentials(sc, s, sx, aux, rc.spp, rc.option_flags);
                        tc = TexCoordCtx{sx.uv, sx.dudx, sx.dudy, sx.dvdx, sx.dvdy};
                        if (tex_needs_ctx(sc)) { pdp[0] = p3fi_mid(s.pi); pdp[1] = sx.dpdx; pdp[2] = sx.dpdy; pdp[3] = s.n; tc.pdp = pdp; }
                        have_ctx = true;
                    }
                    amt = eval_float_texture(sc, mm.tex_mix_amount, tc);
                }
                if (amt <= 0.0f) material_id = (uint32_t)mm.mix_materials[0];
                else if (amt >= 1.0f) material_id = (uint32_t)mm.mix_materials[1];
                else { const float u = mr.get_1d(); material_id = (uint32_t)(amt < u ? mm.mix_materials[0] : mm.mix_materials[1]); }
            }
            st.mat_override[path] = material_id;
            kind = 1 + sc.materials[material_id].kind;
        }
#pragma unroll 1
        for (int k = 1; k < Q_MIX; ++k) {
            const uint32_t mask = __ballot_sync(0xffffffffu, kind == k);
            if (mask) {
                uint32_t qbase = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) qbase = atomicAdd(C + C_NSHADE + k, (uint32_t)__popc(mask));
                qbase = __shfl_sync(0xffffffffu, qbase, leader);
                if (kind == k) q.shade[k][qbase + __popc(mask & ((1u << lane) - 1u))] = path;
            }
        }
    }
}

// ---- order-preserving shade queues ----
// Persistent warps finish their rays in any order, so appending hits to the shade queues at retire time (ballot + atomicAdd)
// scatters neighbouring path slots all over the queues -- and the shade kernels then read their 200-350 bytes of path state per
// hit uncoalesced (at depth >= 1 they wait on those loads: issue-active 27 % against 40 % at depth 0).  Instead the closest-hit
// kernel writes the destination queue of ray-queue entry i to ray_kind[i], and a three-kernel stream compaction builds every
// shade queue in RAY-QUEUE ORDER: per-block counts, an exclusive scan over the blocks, an ordered scatter.  Ray-queue order is
// path-slot order at depth 0 and warp-sized runs of neighbouring slots afterwards (the shade kernels append a warp's survivors
// together), so a shading warp touches a few contiguous stretches of every state array.  Cost: 5 bytes read per ray per pass.
static constexpr int kQueueBlock = 2048;          // entries per block: 256 threads x 8 consecutive entries
SGD void queue_load8(const Queues& q, uint32_t base, uint32_t n, int depth_parity, uint32_t kind8[8], uint32_t path8[8]) {
    const uint32_t* ray = q.ray[depth_parity];
    if (base + 8u <= n) {
        const uint2 kk = *reinterpret_cast<const uint2*>(q.ray_kind + base);
        const uint4 a = *reinterpret_cast<const uint4*>(ray + base), b = *reinterpret_cast<const uint4*>(ray + base + 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) { kind8[j] = (kk.x >> (8 * j)) & 0xffu; kind8[4 + j] = (kk.y >> (8 * j)) & 0xffu; }
        path8[0] = a.x; path8[1] = a.y; path8[2] = a.z; path8[3] = a.w; path8[4] = b.x; path8[5] = b.y; path8[6] = b.z; path8[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = base + j < n;
            kind8[j] = in ? q.ray_kind[base + j] : 0xffu; path8[j] = in ? ray[base + j] : 0u;
        }
    }
}
// per-thread counts of the 8 queues, 16 bits each, in two 64-bit words (a block holds at most 2048 entries of one kind)
SGD void queue_pack_counts(const uint32_t kind8[8], unsigned long long& lo, unsigned long long& hi) {
    lo = hi = 0ull;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t k = kind8[j];
        if (k < 4u) lo += 1ull << (16 * k); else if (k < 8u) hi += 1ull << (16 * (k - 4u));
    }
}
// count / scatter run on a fixed grid that strides over the 2048-entry blocks of the ray queue (its length n is only known on the
// device; late depths hold a few per cent of the wavefront, and a capacity-sized grid of early-exit blocks cost 0.13 ms per launch)
static __global__ void __launch_bounds__(256) k_queue_count(Queues q, int depth) {
    const uint32_t n = q.counters[depth * C_STRIDE + C_NRAY];
    __shared__ unsigned long long s_lo[8], s_hi[8];
    for (uint32_t vb = blockIdx.x; vb * (uint32_t)kQueueBlock < n; vb += gridDim.x) {
        const uint32_t base = vb * (uint32_t)kQueueBlock + threadIdx.x * 8u;
        uint32_t kind8[8], path8[8];
        queue_load8(q, base, n, depth & 1, kind8, path8);
        unsigned long long lo, hi;
        queue_pack_counts(kind8, lo, hi);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { lo += __shfl_xor_sync(0xffffffffu, lo, o); hi += __shfl_xor_sync(0xffffffffu, hi, o); }
        if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
        __syncthreads();
        if (threadIdx.x < Q_NKINDS) {
            unsigned long long tl = 0ull, th = 0ull;
            for (int w = 0; w < 8; ++w) { tl += s_lo[w]; th += s_hi[w]; }
            const int k = threadIdx.x;
            q.block_counts[(size_t)vb * Q_NKINDS + k] = (uint32_t)(((k < 4 ? tl : th) >> (16 * (k & 3))) & 0xffffull);
        }
        __syncthreads();
    }
}
// exclusive scan of the per-block counts of all eight queues at once: block_counts becomes block bases, the totals go to the depth's
// counters (what the shade kernels read as their work count).  One 1024-thread block; a thread owns `per` consecutive blocks (two
// 16-byte loads per block), the partial sums are scanned with warp shuffles + one pass over the 32 warp totals.
static_assert(Q_NKINDS == 8, "k_queue_scan reads a block's eight counts as two uint4");
static __global__ void __launch_bounds__(1024) k_queue_scan(Queues q, int depth, uint32_t queue_mask) {
    const uint32_t n = q.counters[depth * C_STRIDE + C_NRAY];
    const uint32_t n_blocks = (n + kQueueBlock - 1) / kQueueBlock;
    __shared__ uint32_t s_warp[32][Q_NKINDS];
    const uint32_t per = (n_blocks + 1023u) / 1024u;                         // consecutive blocks per thread
    const uint32_t b0 = min(threadIdx.x * per, n_blocks), b1 = min(b0 + per, n_blocks);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* bc = reinterpret_cast<uint4*>(q.block_counts);
    uint32_t sum[Q_NKINDS];
#pragma unroll
    for (int k = 0; k < Q_NKINDS; ++k) sum[k] = 0u;
    for (uint32_t b = b0; b < b1; ++b) {
        const uint4 lo = bc[2 * (size_t)b], hi = bc[2 * (size_t)b + 1];
        sum[0] += lo.x; sum[1] += lo.y; sum[2] += lo.z; sum[3] += lo.w; sum[4] += hi.x; sum[5] += hi.y; sum[6] += hi.z; sum[7] += hi.w;
    }
    uint32_t inc[Q_NKINDS];
#pragma unroll
    for (int k = 0; k < Q_NKINDS; ++k) {
        uint32_t v = sum[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        inc[k] = v;
        if (lane == 31) s_warp[warp][k] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < Q_NKINDS; ++k) {
            uint32_t v = s_warp[lane][k];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
            s_warp[lane][k] = v;                                               // inclusive over the warps
        }
    }
    __syncthreads();
    uint32_t run[Q_NKINDS];
#pragma unroll
    for (int k = 0; k < Q_NKINDS; ++k) run[k] = inc[k] - sum[k] + (warp > 0 ? s_warp[warp - 1][k] : 0u);   // exclusive base of this thread's first block
    for (uint32_t b = b0; b < b1; ++b) {
        const uint4 lo = bc[2 * (size_t)b], hi = bc[2 * (size_t)b + 1];
        bc[2 * (size_t)b] = make_uint4(run[0], run[1], run[2], run[3]); bc[2 * (size_t)b + 1] = make_uint4(run[4], run[5], run[6], run[7]);
        run[0] += lo.x; run[1] += lo.y; run[2] += lo.z; run[3] += lo.w; run[4] += hi.x; run[5] += hi.y; run[6] += hi.z; run[7] += hi.w;
    }
    if (threadIdx.x < Q_NKINDS)                                               // a queue this scene never feeds (TraceScene::queue_mask) stays empty
        q.counters[depth * C_STRIDE + C_NSHADE + threadIdx.x] = ((queue_mask >> threadIdx.x) & 1u) ? s_warp[31][threadIdx.x] : 0u;
}
static __global__ void __launch_bounds__(256) k_queue_scatter(Queues q, int depth) {
    const uint32_t n = q.counters[depth * C_STRIDE + C_NRAY];
    __shared__ unsigned long long s_lo[8], s_hi[8];
  for (uint32_t vb = blockIdx.x; vb * (uint32_t)kQueueBlock < n; vb += gridDim.x) {
    const uint32_t base = vb * (uint32_t)kQueueBlock + threadIdx.x * 8u;
    uint32_t kind8[8], path8[8];
    queue_load8(q, base, n, depth & 1, kind8, path8);
    unsigned long long lo, hi;
    queue_pack_counts(kind8, lo, hi);
    // exclusive prefix of the packed counts over the block's 256 threads (entry order = thread order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long ilo = lo, ihi = hi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long a = __shfl_up_sync(0xffffffffu, ilo, o), b = __shfl_up_sync(0xffffffffu, ihi, o);
        if (lane >= o) { ilo += a; ihi += b; }
    }
    if (lane == 31) { s_lo[warp] = ilo; s_hi[warp] = ihi; }
    __syncthreads();
    unsigned long long plo = ilo - lo, phi = ihi - hi;
    for (int w = 0; w < warp; ++w) { plo += s_lo[w]; phi += s_hi[w]; }
    uint32_t pos[Q_NKINDS];
#pragma unroll
    for (int k = 0; k < Q_NKINDS; ++k)
        pos[k] = q.block_counts[(size_t)vb * Q_NKINDS + k] + (uint32_t)(((k < 4 ? plo : phi) >> (16 * (k & 3))) & 0xffffull);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t k = kind8[j];
        if (k < (uint32_t)Q_NKINDS) {
            uint32_t p = 0;
#pragma unroll
            for (int kk = 0; kk < Q_NKINDS; ++kk) if (k == (uint32_t)kk) p = pos[kk]++;
            q.shade[k][p] = path8[j];
        }
    }
    __syncthreads();                                                        // s_lo / s_hi are reused by the next block of entries
  }
}

// ---- shade queue of one material KIND re-ordered by material ID (textured scenes) ----
// The closest-hit kernel bins hits by material kind only.  In a textured scene the materials of one kind run different code
// (EWA / trilinear / bilinear lookups, bump mapping, untextured) and the textured shade kernels are 17 k instructions long: with
// every CTA working on a mix of materials the SMs' instruction caches thrash (ncu on C4: `no_instruction` = 70 % of all stall
// samples, issue-active 12 %).  A counting sort over 64 material buckets makes a grid sweep of the shade kernel cover ONE material
// (almost) everywhere, so that all warps of an SM fetch the same instructions.  Nothing in the result depends on the order.
SGD uint32_t material_bucket(const DScene& sc, const PathState& st, uint32_t path) {
    const uint32_t mat = __float_as_uint(__ldg(sc.tri_verts + 3 * (size_t)st.hit_prim[path]).w) & 0x7fffffu;
    return (mat ^ (mat >> 6)) & 63u;
}
static __global__ void __launch_bounds__(256) k_sort_queue_count(const __grid_constant__ DScene sc, PathState st, Queues q, int depth, int qk) {
    __shared__ uint32_t s_hist[64];
    if (threadIdx.x < 64) s_hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t n = q.counters[depth * C_STRIDE + C_NSHADE + qk];
    const uint32_t* queue = q.shade[qk];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&s_hist[material_bucket(sc, st, queue[i])], 1u);
    __syncthreads();
    if (threadIdx.x < 64 && s_hist[threadIdx.x]) atomicAdd(&q.sort_hist[threadIdx.x], s_hist[threadIdx.x]);
}
static __global__ void __launch_bounds__(256) k_sort_queue_scatter(const __grid_constant__ DScene sc, PathState st, Queues q, int depth, int qk) {
    __shared__ uint32_t s_cnt[64], s_base[64], s_prefix[64];
    const uint32_t n = q.counters[depth * C_STRIDE + C_NSHADE + qk];
    const uint32_t* queue = q.shade[qk];
    if (threadIdx.x < 32) {                                                          // exclusive scan of the 64 global bucket counts
        const int lane = threadIdx.x;
        const uint32_t a = q.sort_hist[2 * lane], b = q.sort_hist[2 * lane + 1];
        uint32_t incl = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        s_prefix[2 * lane] = incl - a - b; s_prefix[2 * lane + 1] = incl - b;
    }
    // every block takes the same chunks as in the count pass would be unnecessary: any partition works, ranks are reserved per block
    const uint32_t chunk = 256u * 8u;
    for (uint32_t c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {
        if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t e_path[8], e_slot[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t i = c0 + k * 256u + threadIdx.x;
            e_slot[k] = 0xffffffffu; e_path[k] = 0u;
            if (i < n) {
                e_path[k] = queue[i];
                const uint32_t b = material_bucket(sc, st, e_path[k]);
                e_slot[k] = (b << 16) | atomicAdd(&s_cnt[b], 1u);                    // rank inside the block's share of the bucket (<= 2048)
            }
        }
        __syncthreads();
        if (threadIdx.x < 64) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? s_prefix[threadIdx.x] + atomicAdd(&q.sort_hist[64 + threadIdx.x], s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) if (e_slot[k] != 0xffffffffu) q.sorted[s_base[e_slot[k] >> 16] + (e_slot[k] & 0xffffu)] = e_path[k];
        __syncthreads();
    }
}

// ---- surface shading, one kernel per material kind: PathIntegrator::li body integrator.rs:796-891
//      + sample_ld :897-963 ----
// lean variants: CTAs of SG_SHADE_THREADS threads, SG_SHADE_MIN_BLOCKS of them per SM.  16 warps per SM (<= 128 registers, which the
// diffuse / conductor kernels do not even reach); 20 warps (96 registers) was measured slower in round 1 -- C2 557.5 -> 545.4,
// C3 543.6 -> 522.9 Mpaths/s.
#ifndef SG_SHADE_THREADS
#define SG_SHADE_THREADS 256
#endif
#ifndef SG_SHADE_MIN_BLOCKS
#define SG_SHADE_MIN_BLOCKS (512 / SG_SHADE_THREADS)
#endif
// textured variants (C4): 128-thread CTAs, resident blocks per SM measured on one box, same build otherwise -- 3 (168 regs): 176.1,
// 4 (128): 186.8, 5 (96): 188.4, 6 (80): 186.5 Mpaths/s.  Round 2 (staged shading, material sort): 2 CTAs of 256 threads (128 regs, no
// spills) 371.5 against 359.1 for 5 x 128 (96 regs) and 359.4 for 3 x 256 (80 regs); stage barriers change nothing here (r02_c33).
#ifndef SG_SHADE_THREADS_TEX
#define SG_SHADE_THREADS_TEX 256
#endif
#ifndef SG_SHADE_MIN_BLOCKS_TEX
#define SG_SHADE_MIN_BLOCKS_TEX (512 / SG_SHADE_THREADS_TEX)
#endif
// TEX = the scene has image textures (or a non-zero constant displacement): screen-space differentials, texture lookups, bump /
// normal mapping and specular ray-differential propagation (sg_texture.cuh) are compiled in.  TEX = false is the lean variant
// for untextured scenes lit by triangle emitters: no call sites on its hot path.
// PATH = false: the SimplePathIntegrator (integrator.rs:570-728) / RandomWalkIntegrator (:458-568) bodies, chosen at run time by
// rc.integrator; instantiated with TEX = true only (that variant is a superset: it also renders untextured scenes).
// LG = the scene has lights that are not triangle emitters (see light_sample_li<GENERAL>); LG implies TEX.
// FD = Options::force_diffuse (interaction.rs:258-273): a separate set of instantiations (general superset only), so the regular
// kernels carry none of it.
template <bool FD, class A, class B> SGD auto& pick_bsdf(A& a, B& b) { if constexpr (FD) return b; else return a; }
// STAGE (textured scenes, Diffuse materials, path integrator): 0 = the whole body in one kernel; 1 = get_bsdf only -- surface,
// differentials, bump / normal map, texture lookups -- leaving a 96-byte BSDF record per hit (PathState::rec); 2 = everything
// else -- emission + MIS, sample_ld, BSDF sampling, Russian roulette, the queue appends -- from that record, compiled WITHOUT any
// texture code (TEX = false).  The fused textured kernel is 17 k instructions and instruction-fetch bound (ncu on C4:
// `no_instruction` the top stall even after the material sort, issue-active 21 %); the two stages each fit the instruction
// caches far better.  Same arithmetic in the same order per path, so the films are bit-identical to STAGE 0's.
template <int KIND, bool TEX, bool PATH = true, bool LG = TEX, bool FD = false, int STAGE = 0>
__global__ void __launch_bounds__(TEX ? SG_SHADE_THREADS_TEX : SG_SHADE_THREADS, TEX ? SG_SHADE_MIN_BLOCKS_TEX : SG_SHADE_MIN_BLOCKS) k_shade(const __grid_constant__ DScene sc, PathState st, Queues q, RenderConst rc, int depth) {
    uint32_t* C = q.counters + depth * C_STRIDE;
    uint32_t* Cn = C + C_STRIDE;
    const int qk = 1 + KIND;
    const uint32_t n = C[C_NSHADE + qk];
    const uint32_t* queue = q.shade[qk];
    uint32_t* q_next = q.ray[(depth + 1) & 1];
    const int lane = threadIdx.x & 31;
    // Work distribution: a CTA takes chunks of kThreads x ITEMS queue entries.  Textured scenes (1024-entry chunks) first sort the chunk by
    // material in shared memory (counting sort over 64 hash buckets), so that the 32 lanes of a warp run the same texture
    // filter and bump-map code: shade queues are binned by material KIND only, and a warp of mixed EWA / bilinear / untextured
    // lanes runs at a few active lanes per instruction.  Nothing in the result depends on the processing order.
    constexpr uint32_t kThreads = TEX ? (uint32_t)SG_SHADE_THREADS_TEX : (uint32_t)SG_SHADE_THREADS;
    constexpr int ITEMS = TEX ? 1024 / (int)kThreads : 1;
    constexpr uint32_t kChunk = kThreads * ITEMS;
    __shared__ uint32_t s_sorted[TEX ? 1024 : 1];
    __shared__ uint32_t s_hist[64], s_off[64];
    for (uint32_t chunk = blockIdx.x * kChunk; chunk < n; chunk += gridDim.x * kChunk) {
      if constexpr (TEX) {
        if (threadIdx.x < 64) s_hist[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t e_path[ITEMS], e_slot[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t i = chunk + k * kThreads + threadIdx.x;
            e_slot[k] = 0xffffffffu; e_path[k] = 0u;
            if (i < n) {
                e_path[k] = queue[i];
                const uint32_t mat = __float_as_uint(__ldg(sc.tri_verts + 3 * (size_t)st.hit_prim[e_path[k]]).w) & 0x7fffffu;
                const uint32_t bucket = (mat ^ (mat >> 6)) & 63u;
                e_slot[k] = (bucket << 16) | atomicAdd(&s_hist[bucket], 1u);          // rank inside the bucket (chunk <= 1024 entries)
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {                                                      // exclusive scan of the 64 bucket counts
            const uint32_t a = s_hist[2 * lane], b = s_hist[2 * lane + 1];
            uint32_t incl = a + b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            s_off[2 * lane] = incl - a - b; s_off[2 * lane + 1] = incl - b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) if (e_slot[k] != 0xffffffffu) s_sorted[s_off[e_slot[k] >> 16] + (e_slot[k] & 0xffffu)] = e_path[k];
        __syncthreads();
      }
#pragma unroll 1
      for (int it = 0; it < ITEMS; ++it) {
        const uint32_t i = chunk + it * kThreads + threadIdx.x;
        // stage barriers (full chunks only -- every thread of the CTA has an item and reaches them)
        const int sync_mask = (chunk + kChunk <= n) ? (TEX ? rc.shade_sync_tex : rc.shade_sync) : 0;
#define SG_STAGE_SYNC(bit) do { if (sync_mask & (bit)) __syncthreads(); } while (0)
        bool want_shadow = false, want_next = false;
        uint32_t path = 0;
        if (i < n) {
            path = TEX ? s_sorted[it * kThreads + threadIdx.x] : queue[i];
            const float4 rd4 = st.ray_d[path];
            const float3 rd = f3(rd4.x, rd4.y, rd4.z);
            const float3 wo = -rd;
            const int prim_id = st.hit_prim[path];
            const float4 hb = st.hit_b[path];
            uint32_t material_id; int light_id;
            const TriGeo geo = geo_from_prim(sc, (uint32_t)prim_id, material_id, light_id);
            uint32_t fl = st.flags[path];
            int pdepth = fl & 0xff; bool specular_bounce = (fl >> 8) & 1u; bool any_non_specular = (fl >> 9) & 1u;
            Wavelengths lam; lam.lambda = st.lambda[path]; lam.pdf = st.lpdf[path];
            // depth 0 (warp-uniform kernel argument): L = 0, beta = 1, p_b = eta_scale = 1 are not in memory yet (see k_generate)
            Spec L = (STAGE == 1 || depth == 0) ? spec1(0.0f) : st.L[path]; Spec beta = (STAGE == 1 || depth == 0) ? spec1(1.0f) : st.beta[path];
            float2 pbe = (STAGE == 1 || depth == 0) ? make_float2(1.0f, 1.0f) : st.pb_eta[path];
            float p_b = pbe.x, eta_scale = pbe.y;

            SurfTex sx;
            Surf s;
            float3 wo_si = wo;                                   // SurfaceInteraction::wo (what sample_ld reads, integrator.rs:905-917)
            BSDF<KIND> mb;
            if constexpr (STAGE == 2) {                          // the record stage 1 left: interaction + shading frame + BSDF parameters
                const float4 r0 = st.rec[0][path], r1 = st.rec[1][path], r2 = st.rec[2][path], r3 = st.rec[3][path], r4 = st.rec[4][path];
                s.pi.lo = f3(r0.x, r0.y, r0.z); s.pi.hi = f3(r0.w, r1.x, r1.y);
                s.n = f3(r1.z, r1.w, r2.x); s.sn = f3(r2.y, r2.z, r2.w);
                s.sdpdu = s.sdpdv = f3(0.0f, 0.0f, 0.0f);
                mb.fx = f3(r3.x, r3.y, r3.z); wo_si = f3(r3.w, r4.x, r4.y);
                mb.r = st.rec[5][path];
            } else {
            // scenes with instances / spheres / patches resolve the hit out of line, so that triangle-only scenes (hit_inst ==
            // nullptr) keep a small kernel: the shade kernels are sensitive to instruction-cache footprint
            if (st.hit_inst == nullptr) s = make_surface<TEX>(sc, geo, hb.x, hb.y, hb.z, &sx);
            else s = surface_general<TEX>(sc, geo, hb, rd, st.hit_inst[path], &sx, wo_si);
            }

            SG_STAGE_SYNC(1);
            const bool simple = !PATH && rc.integrator == SG_INTEGRATOR_SIMPLE_PATH, walk = !PATH && !simple;
            const bool sample_lights = (rc.integrator_flags & SG_SIMPLEPATH_SAMPLE_LIGHTS) != 0, sample_bsdf_dir = (rc.integrator_flags & SG_SIMPLEPATH_SAMPLE_BSDF) != 0;
            // emission + MIS against light sampling, :798-813
            if (STAGE != 1 && light_id >= 0) {
                const SgLight lt = sc.lights[light_id];
                Spec le = light_l(sc, lt, s.n, wo, lam);
                if (!spec_zero(le)) {
                    if (!PATH) {                                         // SimplePath :625-629 (specular_bounce starts true), RandomWalk :519-520
                        if (walk || !sample_lights || pdepth == 0 || specular_bounce) L = L + beta * le;
                    } else if (pdepth == 0 || specular_bounce) L = L + beta * le;
                    else {
                        LightCtx pc;
                        const float4 c0 = st.ctx0[path], c1 = st.ctx1[path], c2 = st.ctx2[path];
                        pc.pi.lo = f3(c0.x, c0.y, c0.z); pc.pi.hi = f3(c0.w, c1.x, c1.y);
                        pc.n = f3(c1.z, c1.w, c2.x); pc.ns = f3(c2.y, c2.z, c2.w);
                        float p_l = (1.0f / (float)sc.n_lights) * light_pdf_li<LG>(sc, (uint32_t)light_id, lt, geo, pc, rd);
                        float w_l = power_heuristic(p_b, p_l);
                        L = L + beta * w_l * le;
                    }
                }
            }

            SG_STAGE_SYNC(2);
            // get_bsdf, interaction.rs:187-278 + Material::get_bsdf
            AuxRays aux; aux.has = false;
            if constexpr (STAGE == 2) {                          // built by stage 1; a DiffuseBxDF has no other parameters
                static_assert(STAGE != 2 || KIND == SG_MATERIAL_DIFFUSE, "staged shading carries the Diffuse BSDF record only");
                mb.k = spec1(0.0f); mb.eta = 1.0f; mb.mf = TR::make(0.0f, 0.0f);
                mb.fz = s.sn; mb.fy = cross3(mb.fz, mb.fx);
            } else {
            if (geo.kind == SG_MATERIAL_MIX) material_id = st.mat_override[path];                   // resolved by k_resolve_mix (interaction.rs:206-221)
            const SgMaterial mat = sc.materials[material_id];
            if (TEX) {
                if (sc.n_textures > 0) {
                    if (fl & kFlagAux) aux = aux_load(st, path);
                    compute_differentials(sc, s, sx, aux, rc.spp, rc.option_flags);                 // interaction.rs:201
                }
                if (mat.flags & SG_MAT_HAS_DISPLACEMENT) {
                    if (mat.tex_displacement >= 0 || mat.displacement != 0.0f) bump_map(sc, mat.tex_displacement, mat.displacement, s, sx);
                } else if (mat.normal_map >= 0) normal_map(sc, mat.normal_map, s, sx);              // only without a displacement: interaction.rs:229-244
            }
            if ((mat.flags & SG_MAT_HAS_DISPLACEMENT) || (TEX && mat.normal_map >= 0)) apply_constant_bump(s);
            float3 pdp[4];
            TexCoordCtx tc;
            if (TEX) {
                tc = TexCoordCtx{sx.uv, sx.dudx, sx.dudy, sx.dvdx, sx.dvdy};
                if (tex_needs_ctx(sc)) { pdp[0] = p3fi_mid(s.pi); pdp[1] = sx.dpdx; pdp[2] = sx.dpdy; pdp[3] = s.n; tc.pdp = pdp; }
            }
            // texture-valued parameters (SgMaterialTextures).  The scalars are plain values outside the branch, so kernels of scenes
            // without the table keep them in registers; `mv` (whose address escapes) is only touched inside it, and its spectra are
            // only read back where ov_mask says a texture supplied them.
            MatTexValues mv;
            float p_ur = mat.u_roughness, p_vr = mat.v_roughness, p_thickness = mat.thickness, p_g = mat.g, p_ur2 = mat.u_roughness2, p_vr2 = mat.v_roughness2;
            uint32_t ov_mask = 0u;
            if (TEX && KIND != SG_MATERIAL_DIFFUSE && sc.material_textures != nullptr) {
                mv.mask = 0u; mv.ur = p_ur; mv.vr = p_vr; mv.thickness = p_thickness; mv.g = p_g; mv.ur2 = p_ur2; mv.vr2 = p_vr2;
                resolve_material_textures(sc, material_id, tc, lam, mv);
                p_ur = mv.ur; p_vr = mv.vr; p_thickness = mv.thickness; p_g = mv.g; p_ur2 = mv.ur2; p_vr2 = mv.vr2; ov_mask = mv.mask;
            }
            mb.r = spec1(0.0f); mb.k = spec1(0.0f); mb.eta = 1.0f; mb.mf = TR::make(0.0f, 0.0f);
            if (KIND == SG_MATERIAL_DIFFUSE) {
                mb.r = spec_clamp(TEX && mat.tex_reflectance >= 0 ? eval_spectrum_texture(sc, mat.tex_reflectance, tc, lam)
                                                                    : spectrum_sample(sc, mat.spec_a, lam), 0.0f, 1.0f);          // material.rs:307-310
            } else if (KIND == SG_MATERIAL_COATED_DIFFUSE) {                                    // material.rs:917-963
                mb.lay.r = spec_clamp(TEX && mat.tex_reflectance >= 0 ? eval_spectrum_texture(sc, mat.tex_reflectance, tc, lam)
                                                                        : spectrum_sample(sc, mat.spec_a, lam), 0.0f, 1.0f);
                float ur = p_ur, vr = p_vr;
                if (mat.flags & SG_MAT_REMAP_ROUGHNESS) { ur = sqrtf(ur); vr = sqrtf(vr); }
                mb.lay.mf = TR::make(ur, vr);
                mb.lay.thickness = p_thickness;
                float se = spectrum_get(sc, mat.spec_c, lam.lambda.x);
                if (sc.spectra[mat.spec_c].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lam);
                if (se == 0.0f) se = 1.0f;
                mb.lay.eta = se;
                mb.lay.albedo = spec_clamp((ov_mask & 2u) ? mv.b : spectrum_sample(sc, mat.spec_b, lam), 0.0f, 1.0f);
                mb.lay.g = clampf(p_g, -1.0f, 1.0f);
                mb.lay.max_depth = mat.max_depth; mb.lay.n_samples = mat.n_samples;
            } else if (KIND == SG_MATERIAL_COATED_CONDUCTOR) {                                 // material.rs:1188-1260
                float iur = p_ur, ivr = p_vr;
                if (mat.flags & SG_MAT_REMAP_ROUGHNESS) { iur = sqrtf(iur); ivr = sqrtf(ivr); }
                mb.lay.mf = TR::make(iur, ivr);
                mb.lay.thickness = p_thickness;
                float ieta = spectrum_get(sc, mat.spec_c, lam.lambda.x);
                if (sc.spectra[mat.spec_c].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lam);
                if (ieta == 0.0f) ieta = 1.0f;
                mb.lay.eta = ieta;
                Spec ce, ck;
                if (!(mat.flags & SG_MAT_CONDUCTOR_REFLECTANCE)) { ce = (ov_mask & 1u) ? mv.a : spectrum_sample(sc, mat.spec_a, lam); ck = (ov_mask & 4u) ? mv.d : spectrum_sample(sc, mat.spec_d, lam); }
                else {                                                                         // :1225-1233
                    const Spec r = spec_clamp((ov_mask & 1u) ? mv.a : spectrum_sample(sc, mat.spec_a, lam), 0.0f, 0.9999f);
                    ce = spec1(1.0f);
                    ck = make_float4(2.0f * sqrtf(r.x) / sqrtf(fmaxf(0.0f, 1.0f - r.x)), 2.0f * sqrtf(r.y) / sqrtf(fmaxf(0.0f, 1.0f - r.y)),
                                     2.0f * sqrtf(r.z) / sqrtf(fmaxf(0.0f, 1.0f - r.z)), 2.0f * sqrtf(r.w) / sqrtf(fmaxf(0.0f, 1.0f - r.w)));
                }
                mb.lay.ce = ce / ieta; mb.lay.ck = ck / ieta;
                float cur = p_ur2, cvr = p_vr2;
                if (mat.flags & SG_MAT_REMAP_ROUGHNESS) { cur = sqrtf(iur); cvr = sqrtf(ivr); } // sic: roughness_to_alpha(iurough), material.rs:1237-1241
                mb.lay.mfb = TR::make(cur, cvr);
                mb.lay.r = spec1(0.0f);
                mb.lay.albedo = spec_clamp((ov_mask & 2u) ? mv.b : spectrum_sample(sc, mat.spec_b, lam), 0.0f, 1.0f);
                mb.lay.g = clampf(p_g, -1.0f, 1.0f);
                mb.lay.max_depth = mat.max_depth; mb.lay.n_samples = mat.n_samples;
            } else {
                float ur = p_ur, vr = p_vr;
                if (mat.flags & SG_MAT_REMAP_ROUGHNESS) { ur = sqrtf(ur); vr = sqrtf(vr); }     // roughness_to_alpha
                if (KIND == SG_MATERIAL_CONDUCTOR) {
                    mb.r = (ov_mask & 1u) ? mv.a : spectrum_sample(sc, mat.spec_a, lam); mb.k = (ov_mask & 2u) ? mv.b : spectrum_sample(sc, mat.spec_b, lam);
                } else {
                    float se = spectrum_get(sc, mat.spec_a, lam.lambda.x);                      // material.rs:609-624
                    if (sc.spectra[mat.spec_a].kind != SG_SPECTRUM_CONSTANT) terminate_secondary(lam);
                    if (se == 0.0f) se = 1.0f;
                    mb.eta = se;
                }
                mb.mf = TR::make(ur, vr);
            }
            mb.fx = normalize3(s.sdpdu); mb.fz = s.sn; mb.fy = cross3(mb.fz, mb.fx);
            }   // STAGE != 2
            if constexpr (STAGE == 1) {                          // hand the hit over to stage 2
                st.rec[0][path] = make_float4(s.pi.lo.x, s.pi.lo.y, s.pi.lo.z, s.pi.hi.x);
                st.rec[1][path] = make_float4(s.pi.hi.y, s.pi.hi.z, s.n.x, s.n.y);
                st.rec[2][path] = make_float4(s.n.z, s.sn.x, s.sn.y, s.sn.z);
                st.rec[3][path] = make_float4(mb.fx.x, mb.fx.y, mb.fx.z, wo_si.x);
                st.rec[4][path] = make_float4(wo_si.y, wo_si.z, 0.0f, 0.0f);
                st.rec[5][path] = mb.r;
            } else {
            // Options::force_diffuse: DiffuseBxDF(rho_hd(si.wo, [get_1d], [get_2d])) on the same frame (interaction.rs:258-273, bxdf.rs:49-71:
            // one BxDF-level sample, no BSDF-level rejection tests, kept when pdf > 0); sampled BEFORE any regularisation, as there.
            BSDF<SG_MATERIAL_DIFFUSE> db;
            if constexpr (FD) {
                Rng frng; { ulonglong2 ra = st.rng_a[path], rb = st.rng_b[path]; frng.s0 = ra.x; frng.s1 = ra.y; frng.s2 = rb.x; frng.s3 = rb.y; }
                mb.layer_seed = layer_seed(frng, 6);
                const float fuc = frng.get_1d();
                float2 fu2; fu2.x = frng.get_1d(); fu2.y = frng.get_1d();
                st.rng_a[path] = make_ulonglong2(frng.s0, frng.s1); st.rng_b[path] = make_ulonglong2(frng.s2, frng.s3);
                Spec fr = spec1(0.0f);
                const float3 wol = mb.to_local(wo_si);
                if (wol.z != 0.0f) {
                    BSDFSample fbs; bool fprop = false;
                    if (mb.sample_local(wol, fuc, fu2, fbs, fprop) && fbs.pdf > 0.0f) fr = fr + fbs.f * fabsf(fbs.wi.z) / fbs.pdf;
                    fr = fr / 1.0f;
                }
                db.r = fr; db.k = spec1(0.0f); db.eta = 1.0f; db.mf = TR::make(0.0f, 0.0f); db.fx = mb.fx; db.fy = mb.fy; db.fz = mb.fz;
            }
            if (PATH && rc.regularize && any_non_specular) {                                    // :825-828
                mb.mf.regularize();
                if (KIND == SG_MATERIAL_COATED_DIFFUSE || KIND == SG_MATERIAL_COATED_CONDUCTOR) mb.lay.mf.regularize();
                if (KIND == SG_MATERIAL_COATED_CONDUCTOR) mb.lay.mfb.regularize();          // LayeredBxDF::regularize bxdf.rs:1616-1619
            }
            auto& bsdf = pick_bsdf<FD>(mb, db);

            bool alive = pdepth != rc.max_depth;                                               // :830-833
            if (alive) {
                pdepth += 1;
                Rng rng; { ulonglong2 a = st.rng_a[path], b = st.rng_b[path]; rng.s0 = a.x; rng.s1 = a.y; rng.s2 = b.x; rng.s3 = b.y; }
                const int bflags = bsdf.flags();
                if constexpr (!PATH) {
                    if (simple && sample_lights) {                                             // integrator.rs:644-672
                        const float ul = rng.get_1d();
                        if (sc.n_lights > 0) {
                            float2 u_light; u_light.x = rng.get_1d(); u_light.y = rng.get_1d();
                            uint32_t li = __float2uint_rz(ul * (float)sc.n_lights);
                            if (li > sc.n_lights - 1) li = sc.n_lights - 1;
                            const float p_choose = 1.0f / (float)sc.n_lights;
                            const SgLight lt = sc.lights[li];
                            LightCtx ctx; ctx.pi = s.pi; ctx.n = s.n; ctx.ns = s.sn;           // LightSampleContext::from(&isect): no nudge
                            LightSample ls;
                            if (light_sample_li<true>(sc, li, lt, ctx, u_light, lam, ls, false) && !spec_zero(ls.l) && ls.pdf > 0.0f) {
                                bsdf.layer_seed = layer_seed(rng, 1);
                                const Spec f = bsdf.f(wo, ls.wi) * absdot3(ls.wi, s.sn);          // wo = -ray.d (:646), not intr.wo
                                if (!spec_zero(f)) {
                                    float3 pf = offset_ray_origin(s.pi, s.n, p3fi_mid(ls.p_light) - p3fi_mid(s.pi));
                                    float3 pt = offset_ray_origin(ls.p_light, ls.n_light, pf - p3fi_mid(ls.p_light));
                                    float3 sd = pt - pf;
                                    st.sh_o[path] = make_float4(pf.x, pf.y, pf.z, 0.0f);
                                    st.sh_d[path] = make_float4(sd.x, sd.y, sd.z, 0.0f);
                                    st.sh_L[path] = beta * f * ls.l / (p_choose * ls.pdf);     // :669
                                    want_shadow = true;
                                }
                            }
                        }
                    }
                    float3 wi_new = f3(0.0f, 0.0f, 0.0f);
                    if (simple && sample_bsdf_dir) {                                           // :675-691
                        const float u = rng.get_1d();
                        float2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
                        BSDFSample bs; bool prop = false;
                        bsdf.layer_seed = layer_seed(rng, 3);
                        alive = bsdf.sample_f(wo, u, u2, bs, prop);
                        if (alive) { beta = beta * (bs.f * absdot3(bs.wi, s.sn) / bs.pdf); specular_bounce = (bs.flags & BX_SPECULAR) != 0; wi_new = bs.wi; }
                    } else if (simple) {                                                       // :692-721 uniform sphere / hemisphere sampling
                        const bool refl = bflags & BX_REFLECTION, trans = bflags & BX_TRANSMISSION;
                        float2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
                        const float phi = 2.0f * kPi * u2.y;
                        if (refl && trans) { const float z = 1.0f - 2.0f * u2.x, r = safe_sqrt(1.0f - z * z); wi_new = f3(r * cosf(phi), r * sinf(phi), z); }
                        else {
                            const float z = u2.x, r = safe_sqrt(1.0f - z * z); wi_new = f3(r * cosf(phi), r * sinf(phi), z);   // sample_uniform_hemisphere sampling.rs:295-304
                            if ((refl && dot3(wo, s.n) * dot3(wi_new, s.n) < 0.0f) || (trans && dot3(wo, s.n) * dot3(wi_new, s.n) > 0.0f)) wi_new = -wi_new;
                        }
                        bsdf.layer_seed = layer_seed(rng, 3);
                        beta = beta * (bsdf.f(wo, wi_new) * absdot3(wi_new, s.sn) / kInv4Pi);    // uniform_hemisphere_pdf() == 1/(4 pi) too (sampling.rs:306-308)
                        specular_bounce = false;
                    } else {                                                                   // RandomWalk :535-566
                        float2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
                        const float z = 1.0f - 2.0f * u2.x, r = safe_sqrt(1.0f - z * z), phi = 2.0f * kPi * u2.y;
                        wi_new = f3(r * cosf(phi), r * sinf(phi), z);
                        bsdf.layer_seed = layer_seed(rng, 1);
                        const Spec f = bsdf.f(wo, wi_new);
                        if (spec_zero(f)) alive = false;
                        else beta = beta * (f * absdot3(wi_new, s.sn)) / (1.0f / (4.0f * kPi));
                        specular_bounce = false;
                    }
                    if (alive && simple && spec_zero(beta)) alive = false;                     // `while !beta.is_zero()` :606
                    if (alive) {
                        const float3 no = offset_ray_origin(s.pi, s.n, wi_new);                // Interaction::spawn_ray: no differentials
                        st.ray_o[path] = make_float4(no.x, no.y, no.z, 0.0f);
                        st.ray_d[path] = make_float4(wi_new.x, wi_new.y, wi_new.z, 0.0f);
                        aux.has = false;
                    }
                } else {
                SG_STAGE_SYNC(4);
                // ---- sample_ld :897-963 ----
#ifdef SG_EXP_SKIP_NEE      /* timing experiment only (tools/r02_exp_split.sh): wrong images */
                if (false) {
#else
                if (bflags & (BX_DIFFUSE | BX_GLOSSY)) {
#endif
                    LightCtx ctx; ctx.pi = s.pi; ctx.n = s.n; ctx.ns = s.sn;
                    const bool refl = bflags & BX_REFLECTION, trans = bflags & BX_TRANSMISSION;
                    if (refl && !trans) ctx.pi = p3fi_exact(offset_ray_origin(s.pi, s.n, wo_si));
                    else if (trans && !refl) ctx.pi = p3fi_exact(offset_ray_origin(s.pi, s.n, -wo_si));
                    const float ul = rng.get_1d();
                    float2 u_light; u_light.x = rng.get_1d(); u_light.y = rng.get_1d();
                    if (sc.n_lights > 0) {
                        // UniformLightSampler::sample_light light_sampler.rs:91-103 (`as usize` saturates, NaN -> 0)
                        uint32_t li = __float2uint_rz(ul * (float)sc.n_lights);
                        if (li > sc.n_lights - 1) li = sc.n_lights - 1;
                        const float p_choose = 1.0f / (float)sc.n_lights;
                        const SgLight lt = sc.lights[li];
                        LightSample ls;
                        if (light_sample_li<LG>(sc, li, lt, ctx, u_light, lam, ls) && !spec_zero(ls.l) && ls.pdf != 0.0f) {
                            bsdf.layer_seed = layer_seed(rng, 1);
                            Spec f = bsdf.f(wo_si, ls.wi) * absdot3(ls.wi, s.sn);
                            if (!spec_zero(f)) {
                                const float p_l = p_choose * ls.pdf;
                                Spec ld;
                                if (lt.kind == SG_LIGHT_POINT) ld = ls.l * f / p_l;
                                else {
                                    bsdf.layer_seed = layer_seed(rng, 2);
                                    float p_bsdf = bsdf.pdf(wo_si, ls.wi);
                                    float w_l = power_heuristic(p_l, p_bsdf);
                                    ld = w_l * ls.l * f / p_l;
                                }
                                // Ray::spawn_ray_to_both_offset ray.rs:83-99
                                float3 pf = offset_ray_origin(s.pi, s.n, p3fi_mid(ls.p_light) - p3fi_mid(s.pi));
                                float3 pt = offset_ray_origin(ls.p_light, ls.n_light, pf - p3fi_mid(ls.p_light));
                                float3 sd = pt - pf;
                                st.sh_o[path] = make_float4(pf.x, pf.y, pf.z, 0.0f);
                                st.sh_d[path] = make_float4(sd.x, sd.y, sd.z, 0.0f);
                                st.sh_L[path] = beta * ld;
                                want_shadow = true;
                            }
                        }
                    }
                }
                SG_STAGE_SYNC(8);
                // ---- BSDF sampling :843-875 ----
                const float u = rng.get_1d();
                float2 u2; u2.x = rng.get_1d(); u2.y = rng.get_1d();
                BSDFSample bs; bool prop = false;
                bsdf.layer_seed = layer_seed(rng, 3);
#ifdef SG_EXP_SKIP_SAMPLE   /* timing experiment only: every path ends after its first vertex */
                alive = false;
#else
                alive = bsdf.sample_f(wo, u, u2, bs, prop);
#endif
                if (alive) {
                    beta = beta * (bs.f * absdot3(bs.wi, s.sn) / bs.pdf);
                    if (prop) { bsdf.layer_seed = layer_seed(rng, 4); p_b = bsdf.pdf(wo, bs.wi); }      // pdf_is_proportional :860-865
                    else p_b = bs.pdf;
                    specular_bounce = (bs.flags & BX_SPECULAR) != 0;
                    any_non_specular = any_non_specular || !specular_bounce;
                    if (bs.flags & BX_TRANSMISSION) eta_scale *= sqr(bs.eta);
                    st.ctx0[path] = make_float4(s.pi.lo.x, s.pi.lo.y, s.pi.lo.z, s.pi.hi.x);
                    st.ctx1[path] = make_float4(s.pi.hi.y, s.pi.hi.z, s.n.x, s.n.y);
                    st.ctx2[path] = make_float4(s.n.z, s.sn.x, s.sn.y, s.sn.z);
                    if (TEX && sc.n_textures > 0) {                                            // spawn_ray_with_differentials :434-502
                        aux = spawn_differentials(s, sx, aux, wo, bs.wi, bs.flags, bs.eta);
                        if (aux.has) aux_store(st, path, aux);
                    }
                    const float3 no = offset_ray_origin(s.pi, s.n, bs.wi);                     // Interaction::spawn_ray
                    st.ray_o[path] = make_float4(no.x, no.y, no.z, 0.0f);
                    st.ray_d[path] = make_float4(bs.wi.x, bs.wi.y, bs.wi.z, 0.0f);
                    // Russian roulette :878-891
                    if (isfinite(eta_scale)) {
                        const float m = spec_max(beta * eta_scale);
                        if (m < 1.0f && pdepth > 1) {
                            const float qq = fmaxf(0.0f, 1.0f - m);
                            if (rng.get_1d() < qq) alive = false;
                            else beta = beta / (1.0f - qq);
                        }
                    }
                }
                }   // PATH
                st.rng_a[path] = make_ulonglong2(rng.s0, rng.s1); st.rng_b[path] = make_ulonglong2(rng.s2, rng.s3);
            }
            st.L[path] = L;
            if (alive) {
                st.beta[path] = beta;
                st.pb_eta[path] = make_float2(p_b, eta_scale);
                st.flags[path] = (uint32_t)pdepth | (specular_bounce ? kFlagSpecular : 0u) | (any_non_specular ? kFlagNonSpecular : 0u) |
                                 (TEX && aux.has ? kFlagAux : 0u);
            }
            // terminate_secondary (get_bsdf of a dispersive material).  SimplePath / RandomWalk return at `depth == max_depth` BEFORE
            // get_bsdf (integrator.rs:520-523, :628-631), so their last vertex leaves the wavelengths alone; PathIntegrator::li builds
            // the BSDF first and tests the depth afterwards (:815-833).
            if ((KIND == SG_MATERIAL_DIELECTRIC || KIND == SG_MATERIAL_COATED_DIFFUSE || KIND == SG_MATERIAL_THIN_DIELECTRIC || KIND == SG_MATERIAL_COATED_CONDUCTOR) &&
                (PATH || (int)(fl & 0xffu) != rc.max_depth)) st.lpdf[path] = lam.pdf;
            want_next = alive;
            }   // STAGE != 1
        }
        __syncwarp();
        {   // warp-aggregated queue appends
            uint32_t mask = __ballot_sync(0xffffffffu, want_shadow);
            if (mask) {
                uint32_t qb = 0; const int leader = __ffs(mask) - 1;
                if (lane == leader) qb = atomicAdd(C + C_NSHADOW, (uint32_t)__popc(mask));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (want_shadow) q.shadow[qb + __popc(mask & ((1u << lane) - 1u))] = path;
            }
            mask = __ballot_sync(0xffffffffu, want_next);
            if (mask) {
                uint32_t qb = 0; const int leader = __ffs(mask) - 1;
                if (lane == leader) qb = atomicAdd(Cn + C_NRAY, (uint32_t)__popc(mask));
                qb = __shfl_sync(0xffffffffu, qb, leader);
                if (want_next) q_next[qb + __popc(mask & ((1u << lane) - 1u))] = path;
            }
        }
        SG_STAGE_SYNC(16);
#undef SG_STAGE_SYNC
      }
      if constexpr (TEX) __syncthreads();                                            // s_sorted / s_hist are reused by the next chunk
    }
}

// ---- RgbFilm::add_sample for every path of the batch (film.rs:548-574) ----
// `rgb_sum[c] += (w * rgb[c]) as f64; weight_sum += w` with f64 atomics (RED.ADD.F64): samples of one pixel live
// in different wavefront slots.  Paths are pixel-major, so a warp usually holds 32 samples of ONE pixel: those
// are summed in f64 across the warp first (one atomic per channel per warp instead of 32 same-address ones).
static __global__ void __launch_bounds__(256) k_film(const __grid_constant__ DScene sc, PathState st, uint32_t count, double* film) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < count;
    float rgb[3] = {0.0f, 0.0f, 0.0f};
    uint32_t pixel = 0xffffffffu;
    if (live) {
        const Spec L = st.L[i];
        // a black sample contributes rgb = +0 whatever its wavelengths (0 / pdf, times the sensor curves, summed from +0):
        // skip its 32 bytes of wavelength state and 12 sensor lookups -- half of C5's samples are background
        if (L.x != 0.0f || L.y != 0.0f || L.z != 0.0f || L.w != 0.0f) {
            Wavelengths lam; lam.lambda = st.lambda[i]; lam.pdf = st.lpdf[i];
            film_sample_rgb(sc, L, lam, rgb);
        }
        pixel = st.pixel[i];
    }
    const float weight = 1.0f;                                                      // BoxFilter::sample weight, filter.rs:99-105
    double v[4] = {(double)(weight * rgb[0]), (double)(weight * rgb[1]), (double)(weight * rgb[2]), live ? (double)weight : 0.0};
    // Lanes holding samples of the same pixel form aligned power-of-two runs (n_samples consecutive slots per pixel: 8 lanes each
    // when a 4K batch carries 8 spp, the whole warp at >= 32 spp): butterfly-sum as far as every lane's partner shares its pixel,
    // then the first lane of each run issues the four atomics (16 instead of 128 per warp at 8 spp).
    int run = 1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t other = __shfl_xor_sync(0xffffffffu, pixel, o);
        if (!__all_sync(0xffffffffu, other == pixel)) break;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        run = o << 1;
    }
    if (live && ((threadIdx.x & 31) & (run - 1)) == 0) {
        double* px = film + 4 * (size_t)pixel;
        for (int c = 0; c < 4; ++c) atomicAdd(px + c, v[c]);
    }
}

// Scene upload: evaluate the degenerate-triangle test of triangle.rs:181 once per triangle record (the same device expression the
// traversal used to evaluate per test) and keep the answer in bit 31 of the record's first .w word.
static __global__ void k_mark_degenerate(float4* tri_verts, uint32_t n_prims) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims) return;
    float4* tv = tri_verts + 3 * (size_t)i;
    const float4 v0 = tv[0], v1 = tv[1], v2 = tv[2];
    if (((__float_as_uint(v0.w) >> 28) & 7u) == kKindInstance) return;                  // TransformedPrimitive record
    if (__float_as_uint(v2.w) & (kSphereBit | kPatchBit)) return;                         // sphere / bilinear-patch record
    if (triangle_is_degenerate(f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z)))
        tv[0].w = __uint_as_float(__float_as_uint(v0.w) | kDegenerateBit);
}

static __global__ void k_accum_stats(const uint32_t* counters, int n_depths, DevStats* stats) {
    unsigned long long c = 0, s = 0;
    for (int d = 0; d < n_depths; ++d) { c += counters[d * C_STRIDE + C_NRAY]; s += counters[d * C_STRIDE + C_NSHADOW]; }
    atomicAdd(&stats->closest, c); atomicAdd(&stats->shadow, s);              // two wavefronts (streams) may finish a batch at the same time
}

// ---- free-standing ray batches: the ray-cast parity / roofline entry (sg_trace) ----
template <bool ANY>
struct RaysIO {
    const DScene& sc; const float* o; const float* d; const float* tmax; SgHit* out;
    SGD void load(unsigned long long i, float3& ro, float3& rd, float& t) {
        ro = f3(o[3 * i], o[3 * i + 1], o[3 * i + 2]); rd = f3(d[3 * i], d[3 * i + 1], d[3 * i + 2]); t = tmax[i];
    }
    SGD void retire(bool fin, unsigned long long i, const HitRec& hit, int) {
        if (!fin) return;
        SgHit h; h.prim = -1; h.t = 0.0f; h.b0 = h.b1 = h.b2 = 0.0f; h.ng[0] = h.ng[1] = h.ng[2] = 0.0f;
        if (hit.prim >= 0) {
            if (ANY) h.prim = 0;
            else {
                h.prim = hit.prim; h.t = hit.t; h.b0 = hit.b0; h.b1 = hit.b1; h.b2 = hit.b2;
                const float4 v0 = __ldg(sc.tri_verts + 3 * (size_t)hit.prim), v1 = __ldg(sc.tri_verts + 3 * (size_t)hit.prim + 1),
                             v2 = __ldg(sc.tri_verts + 3 * (size_t)hit.prim + 2);
                if (__float_as_uint(v2.w) & kSphereBit) {                            // geometric normal of the render-space interaction
                    const DSphere& S = sc.spheres[__float_as_uint(v2.w) & ~(kSphereBit | kLastInLeaf)];
                    Surf ss = make_surface_sphere<false>(S, f3(hit.b0, hit.b1, hit.b2), nullptr);
                    float3 wo_si, rd = f3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
                    transform_interaction<false>(sc, S.m, S.mi, rd, ss, nullptr, wo_si);
                    h.ng[0] = ss.n.x; h.ng[1] = ss.n.y; h.ng[2] = ss.n.z;
                    out[i] = h;
                    return;
                }
                if (__float_as_uint(v2.w) & kPatchBit) {
                    const Surf ps = make_surface_patch<false>(sc, __float_as_uint(v2.w) & ~(kPatchBit | kLastInLeaf), hit.b0, hit.b1, nullptr);
                    h.ng[0] = ps.n.x; h.ng[1] = ps.n.y; h.ng[2] = ps.n.z;
                    out[i] = h;
                    return;
                }
                const float3 p0 = f3(v0.x, v0.y, v0.z), p1 = f3(v1.x, v1.y, v1.z), p2 = f3(v2.x, v2.y, v2.z);
                float3 ng = normalize3(cross3(p0 - p2, p1 - p2));                    // triangle.rs:407-412
                const uint32_t mflags = sc.meshes[__float_as_uint(v2.w) & ~kLastInLeaf].flags;
                if (((mflags & SG_MESH_REVERSE_ORIENTATION) != 0) != ((mflags & SG_MESH_SWAPS_HANDEDNESS) != 0)) ng = -ng;
                h.ng[0] = ng.x; h.ng[1] = ng.y; h.ng[2] = ng.z;
            }
        }
        out[i] = h;
    }
};
template <bool ANY, bool COUNT, bool INST>
__global__ void __launch_bounds__(kTraceThreads, INST ? SG_TRACE_MIN_BLOCKS_INST : SG_TRACE_MIN_BLOCKS) k_trace_rays(const __grid_constant__ DScene sc, const __grid_constant__ TraceScene ts,
                                                              long long n, const float* __restrict__ o, const float* __restrict__ d,
                                                              const float* __restrict__ tmax, SgHit* __restrict__ out,
                                                              unsigned long long* cursor, DevStats* stats) {
    extern __shared__ __align__(16) uint32_t s_mem[];
    uint32_t cnt_nodes = 0, cnt_tris = 0;
    RaysIO<ANY> io{sc, o, d, tmax, out};
    if constexpr (SG_TRACE_POSTPONE && !COUNT && !INST) trace_persistent_post<ANY>(ts, io, (unsigned long long)n, cursor, s_mem);
    else trace_persistent<ANY, COUNT, INST>(ts, io, (unsigned long long)n, cursor, s_mem, cnt_nodes, cnt_tris);
    if (COUNT) {
        atomicAdd(&stats->nodes, (unsigned long long)cnt_nodes);
        atomicAdd(&stats->tris, (unsigned long long)cnt_tris);
    }
}

// ---- KAT kernels ----
static __global__ void k_sampler_fill(uint64_t seed, int raw, uint32_t pixel, uint32_t sample, long long n, float* out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Rng r; r.seed_from_u64(raw ? seed : stream_key(seed, pixel, sample));
    for (long long i = 0; i < n; ++i) out[i] = r.get_1d();
}
static __global__ void k_camera_rays(const __grid_constant__ DScene sc, RenderConst rc, long long n, const int* pixel_xy, const int* sample_index,
                              float* out_rays, float* out_lambda) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int px = pixel_xy[2 * i], py = pixel_xy[2 * i + 1];
    Rng rng; rng.seed_from_u64(stream_key(rc.seed, (uint32_t)(py * rc.full_res_x + px), (uint32_t)sample_index[i]));
    Wavelengths lam; float3 o, d; float w;
    camera_stage(sc, rc.option_flags, px, py, rng, lam, o, d, w);
    float* r = out_rays + 6 * i;
    r[0] = o.x; r[1] = o.y; r[2] = o.z; r[3] = d.x; r[4] = d.y; r[5] = d.z;
    float* l = out_lambda + 8 * i;
    l[0] = lam.lambda.x; l[1] = lam.lambda.y; l[2] = lam.lambda.z; l[3] = lam.lambda.w;
    l[4] = lam.pdf.x; l[5] = lam.pdf.y; l[6] = lam.pdf.z; l[7] = lam.pdf.w;
}
static __global__ void k_texture_eval(const __grid_constant__ DScene sc, int tex, int as_float, long long n, const float* q, const float* pdp_in, const float* nrm_in, const float* lambda, float* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    TexCoordCtx c; c.uv = make_float2(q[6 * i], q[6 * i + 1]); c.dudx = q[6 * i + 2]; c.dudy = q[6 * i + 3]; c.dvdx = q[6 * i + 4]; c.dvdy = q[6 * i + 5];
    float3 pdp[4] = {f3(0.0f, 0.0f, 0.0f), f3(0.0f, 0.0f, 0.0f), f3(0.0f, 0.0f, 0.0f), f3(0.0f, 0.0f, 0.0f)};
    if (pdp_in) {
        const float* r = pdp_in + 9 * i;
        pdp[0] = f3(r[0], r[1], r[2]); pdp[1] = f3(r[3], r[4], r[5]); pdp[2] = f3(r[6], r[7], r[8]);
    }
    if (nrm_in) pdp[3] = f3(nrm_in[3 * i], nrm_in[3 * i + 1], nrm_in[3 * i + 2]);
    if (pdp_in || nrm_in) c.pdp = pdp;
    Spec s;
    if (as_float) s = spec1(eval_float_texture(sc, tex, c));
    else {
        Wavelengths w; w.lambda = make_float4(lambda[4 * i], lambda[4 * i + 1], lambda[4 * i + 2], lambda[4 * i + 3]); w.pdf = spec1(1.0f);
        s = eval_spectrum_texture(sc, tex, c, w);
    }
    out[4 * i] = s.x; out[4 * i + 1] = s.y; out[4 * i + 2] = s.z; out[4 * i + 3] = s.w;
}
// RgbFilm::get_pixel_rgb film.rs:720-738
static __global__ void k_film_develop(const __grid_constant__ DScene sc, const double* film, long long n, float* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float rgb[3] = {(float)film[4 * i], (float)film[4 * i + 1], (float)film[4 * i + 2]};
    const double ws = film[4 * i + 3];
    if (ws != 0.0) { rgb[0] /= (float)ws; rgb[1] /= (float)ws; rgb[2] /= (float)ws; }
    const float* M = sc.film.output_rgb_from_sensor_rgb;
    for (int r = 0; r < 3; ++r) out[3 * i + r] = M[3 * r] * rgb[0] + M[3 * r + 1] * rgb[1] + M[3 * r + 2] * rgb[2];
}
// RgbFilm::get_image film.rs:647-707 + Image::set_channel image.rs:648-661 (see sg_film_get_image in shimmer_gpu.h)
static __global__ void k_film_image(const __grid_constant__ DScene sc, const double* film, int w, int h, uint32_t flags, float* out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)w * h) return;
    float rgb[3] = {(float)film[4 * i], (float)film[4 * i + 1], (float)film[4 * i + 2]};
    const double ws = film[4 * i + 3];
    if (ws != 0.0) { rgb[0] /= (float)ws; rgb[1] /= (float)ws; rgb[2] /= (float)ws; }
    const float* M = sc.film.output_rgb_from_sensor_rgb;
    float o[3];
    for (int r = 0; r < 3; ++r) o[r] = M[3 * r] * rgb[0] + M[3 * r + 1] * rgb[1] + M[3 * r + 2] * rgb[2];
    if (flags & SG_IMAGE_FP16) {
        const float max_f16 = 65504.0f;
        // the fold uses Float::max, which ignores NaN operands
        if (fmaxf(fmaxf(fmaxf(-INFINITY, o[0]), o[1]), o[2]) > max_f16) {
            if (o[0] > max_f16) o[0] = max_f16;
            if (o[1] > max_f16) o[0] = max_f16;                         // sic: film.rs:683-685 assigns r
            if (o[2] > max_f16) o[2] = max_f16;
        }
    }
    const int y = (int)(i / w), x = (int)(i - (long long)y * w);
    const long long dst = (flags & SG_IMAGE_BOTTOM_UP) ? (long long)(h - 1 - y) * w + x : i;
    for (int c = 0; c < 3; ++c) {
        float v = o[c];
        if (isnan(v)) v = 0.0f;
        if (flags & SG_IMAGE_FP16) v = __half2float(__float2half_rn(v));
        out[3 * dst + c] = v;
    }
}

}  // namespace sg
