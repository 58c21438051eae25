// Traversal kernel core, generation 2: the hot loop of the path tracer.
//
// Same visiting order, same box/triangle arithmetic and therefore the same hits, ties and visit
// counts as BvhAggregate::intersect / intersect_predicate (aggregate.rs:71-203) -- but laid out
// for a GPU:
//   * Node64: an interior node stores BOTH children's bounds (64 B = 4 x LDG.128), so one fetch
//     feeds two slab tests and the dependent-load chain per level is halved.  The reference
//     tests a node's bounds when it is visited/popped; here the near child is tested at once and
//     the far child's entry distance is pushed with it, re-checked against the current t_max on pop -- exactly the `t_min < ray_t_max` term of intersect_p_cached
//     (bounding_box.rs:563), the only term that depends on t_max.
//   * warp-voted phase scheduling (sg_wavefront.cuh trace_persistent): every iteration the warp
//     ballots its lanes' states (interior / holding a leaf / finished) and runs ONE phase for
//     all lanes that want it -- interior steps by default, triangle tests once enough lanes hold
//     a leaf, retire+refill once enough lanes finished -- so box-test lanes never idle behind a
//     130-instruction triangle test and finished lanes never wait for the slowest ray.
//   * persistent warps with per-lane replacement: finished lanes claim new rays with one
//     warp-aggregated atomicAdd.
//   * per-thread traversal stack in shared memory, level-major (bank-conflict-free), sized from
//     the tree depth computed at upload.
#pragma once
#include "sg_scene.cuh"
#include "sg_sphere.cuh"
#include "sg_patch.cuh"

namespace sg {

static constexpr uint32_t kLeafBit = 0x80000000u;
static constexpr uint32_t kFailBit = 0x40000000u;     // COUNT builds only: far child failed its box test
static constexpr uint32_t kEmptyRef = 0x7fffffffu;
static constexpr uint32_t kLastInLeaf = 0x80000000u;  // flag in tri_verts[3*i+2].w (mesh id word)

// slab test of bounding_box.rs:520-564 split into its t_max-independent part (return value) and
// the entry distance the `t_min < ray_t_max` term needs.
SGD bool slab_entry(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                    float3 o, float3 inv_dir, int nx, int ny, int nz, float& t_entry) {
    const float k = 1.0f + 2.0f * gamma_n(3);
    float t_min = ((nx ? bmaxx : bminx) - o.x) * inv_dir.x;
    float t_max = ((nx ? bminx : bmaxx) - o.x) * inv_dir.x;
    float ty_min = ((ny ? bmaxy : bminy) - o.y) * inv_dir.y;
    float ty_max = ((ny ? bminy : bmaxy) - o.y) * inv_dir.y;
    t_max *= k; ty_max *= k;
    bool ok = !(t_min > ty_max || ty_min > t_max);
    if (ty_min > t_min) t_min = ty_min;
    if (ty_max < t_max) t_max = ty_max;
    float tz_min = ((nz ? bmaxz : bminz) - o.z) * inv_dir.z;
    float tz_max = ((nz ? bminz : bmaxz) - o.z) * inv_dir.z;
    tz_max *= k;
    ok = ok && !(t_min > tz_max || tz_min > t_max);
    if (tz_min > t_min) t_min = tz_min;
    if (tz_max < t_max) t_max = tz_max;
    t_entry = t_min;
    return ok && t_max > 0.0f;
}

struct TraceScene {
    const float4* node64;       // 4 float4 per interior node
    const float4* tri_verts;    // 3 float4 per primitive (see sg_scene.cuh)
    float root_bmin[3], root_bmax[3];
    uint32_t root_ref;          // kEmptyRef for an empty scene
    int stack_depth;            // entries per thread (tree depth)
    int smem_levels;            // entries per thread held in shared memory; deeper entries spill to local memory
    int leaf_threshold;         // warp-vote scheduling knobs (see trace_persistent)
    int refill_threshold;
    int interior_burst;
    int refill_threshold_d0;    // the same knobs for the CLOSEST-HIT launch of depth 0 (camera rays: a warp's rays are near-identical, so
    int interior_burst_d0;      //   waiting for most lanes before a refill keeps them in lockstep)
    int prefetch;               // unused (SG_TRACE_PREFETCH is a build-time switch)
    int refill_threshold_dual, leaf_threshold_dual;      // two-rays-per-lane kernels (k_trace_dual): thresholds count SLOTS (64 per warp)
    int dual_levels_closest, dual_levels_shadow;         //   shared-memory stack levels per ray
    const DInstance* instances; // object instancing (INST kernels only)
    const float4* patch_verts;  // bilinear patches (INST kernels only)
    const DSphere* spheres;     // sphere shapes (INST kernels only: the "general" kernels handle everything that is not a triangle)
    uint32_t scene_flags;
    uint32_t queue_mask;        // bit k: shade queue k (sg_wavefront.cuh Q_*) can receive hits in this scene -- the retire step skips the others
};

// Per-thread traversal stack: the first `levels` entries live in shared memory (level-major, so a warp's
// accesses to one level hit 32 consecutive banks), deeper ones in a small local-memory array.  Capping the
// shared part is what lets 8 CTAs (32 warps) share an SM on deep trees; the spill path is rarely taken.
static constexpr int kSpillLevels = 48;
struct Stack {
    uint32_t a_ref;         // shared-window byte address of this thread's level-0 slot: 32-bit st.shared / ld.shared, where a generic
                            //   pointer costs a window-base lookup and a 64-bit address per access.  Closest-hit stacks hold (ref, entry
                            //   distance) PAIRS, 8 bytes per entry and level-major (one 64-bit access per push / pop); any-hit stacks hold refs.
    int stride; int levels;
    // `stacks`: start of this stack's region (8-byte aligned); the thread's level-0 slot is `tid` entries in
    template <bool ANY> SGD void bind(uint32_t* stacks, int tid) { a_ref = (uint32_t)__cvta_generic_to_shared(stacks + tid * (ANY ? 1 : 2)); }
    uint2* spill;
    float* s_save;          // INST kernels: 10 words per thread (stride apart) holding the render-space ray while the lane is inside an instance
    template <bool ANY> SGD void put(int sp, uint32_t ref, float t) const {
        if (sp < levels) {
            if (ANY) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_ref + (uint32_t)(sp * stride) * 4u), "r"(ref));
            else asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_ref + (uint32_t)(sp * stride) * 8u), "r"(ref), "r"(__float_as_uint(t)));
        }
        else spill[sp - levels] = make_uint2(ref, __float_as_uint(t));
    }
    template <bool ANY> SGD void get(int sp, uint32_t& ref, float& t) const {
        if (sp < levels) {
            if (ANY) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ref) : "r"(a_ref + (uint32_t)(sp * stride) * 4u));
            else { uint32_t tb; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(ref), "=r"(tb) : "r"(a_ref + (uint32_t)(sp * stride) * 8u)); t = __uint_as_float(tb); }
        }
        else { const uint2 e = spill[sp - levels]; ref = e.x; t = __uint_as_float(e.y); }
    }
};

// One lane's traversal state.
struct Lane {
    float3 o, inv_dir;
    RayPre rp;
    float t_max;
    uint32_t cur;
    int sp;
    int nx, ny, nz;
    HitRec hit;
    // object instancing (INST kernels only): the instance being traversed (-1: top level), the stack level its
    // traversal started at, the top-level t_max to restore on the way out, and whether it produced a hit
    int inst; int sp_base; float t_saved; bool inst_hit;
};

SGD void lane_set_ray(Lane& L, float3 o, float3 d) {
    L.o = o;
    L.inv_dir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                 // aggregate.rs:76
    L.nx = L.inv_dir.x < 0.0f; L.ny = L.inv_dir.y < 0.0f; L.nz = L.inv_dir.z < 0.0f;
    L.rp = ray_precompute(d);
}

template <bool ANY>
SGD void lane_begin(const TraceScene& ts, Lane& L, float3 o, float3 d, float t_max, uint32_t& n_nodes, bool count) {
    lane_set_ray(L, o, d);
    L.t_max = t_max; L.sp = 0; L.hit.prim = -1;
    L.cur = kEmptyRef;
    if (ts.root_ref != kEmptyRef) {
        float te;
        if (count) n_nodes++;
        bool ok = slab_entry(ts.root_bmin[0], ts.root_bmin[1], ts.root_bmin[2], ts.root_bmax[0], ts.root_bmax[1], ts.root_bmax[2],
                             o, L.inv_dir, L.nx, L.ny, L.nz, te);
        if (ok && te < t_max) L.cur = ts.root_ref;
    }
}

// Transform::apply_ray_inverse (transform.rs:701-723, inverse = true) / apply_ray (:515-532, inverse = false) with Some(t_max):
// the origin is shifted to the edge of its error box and t_max shrinks by dt -- a real shift for the inverse variant only.
SGD void instance_ray(const DInstance& I, bool inverse, float3& o, float3& d, float& t_max) {
    const float* m = inverse ? I.mi : I.m;
    const float x = o.x, y = o.y, z = o.z;
    P3fi oi;
    if (inverse) {          // apply_ray_inverse: Point3fi transform of an exact point, error term without the translation column (:650-662)
        const float xp = (m[0] * x + m[1] * y) + (m[2] * z + m[3]);
        const float yp = (m[4] * x + m[5] * y) + (m[6] * z + m[7]);
        const float zp = (m[8] * x + m[9] * y) + (m[10] * z + m[11]);
        const float3 err = f3(gamma_n(3) * (fabsf(m[0] * x) + fabsf(m[1] * y) + fabsf(m[2] * z)), gamma_n(3) * (fabsf(m[4] * x) + fabsf(m[5] * y) + fabsf(m[6] * z)),
                              gamma_n(3) * (fabsf(m[8] * x) + fabsf(m[9] * y) + fabsf(m[10] * z)));
        oi = p3fi_make(f3(xp, yp, zp), err);
    } else {                // apply_ray (:515-532): the Point3f overload (apply_point_helper, left to right) -> zero-width interval, dt = 0
        oi = p3fi_exact(f3(((m[0] * x + m[1] * y) + m[2] * z) + m[3], ((m[4] * x + m[5] * y) + m[6] * z) + m[7], ((m[8] * x + m[9] * y) + m[10] * z) + m[11]));
    }
    const float3 dd = f3(m[0] * d.x + m[1] * d.y + m[2] * d.z, m[4] * d.x + m[5] * d.y + m[6] * d.z, m[8] * d.x + m[9] * d.y + m[10] * d.z);
    const float ls = len2(dd);
    if (ls > 0.0f) {
        const float dt = dot3(abs3(dd), p3fi_err(oi)) / ls;
        t_max = t_max - dt;
        const float3 off = dd * dt;
        oi.lo = f3(next_down(oi.lo.x + off.x), next_down(oi.lo.y + off.y), next_down(oi.lo.z + off.z));   // interval.rs:353-356
        oi.hi = f3(next_up(oi.hi.x + off.x), next_up(oi.hi.y + off.y), next_up(oi.hi.z + off.z));
    }
    o = p3fi_mid(oi); d = dd;
}

// TransformedPrimitive::intersect / intersect_predicate (primitive.rs:155-175): move the lane into the instance's space
// and start on the object's BvhAggregate.  The closest-hit path uses the inverse transform; the predicate uses the FORWARD
// one in the reference (unless SG_SCENE_FIX_INSTANCING).
// The render-space ray of a lane that enters an instance is parked in shared memory and restored on the way out: re-deriving it
// (queue -> path -> ray_o / ray_d: three dependent global loads, then six IEEE divides for 1 / d and the shear constants) was a
// large part of the instanced kernels' instance-transition cost, paid at leaf-phase lane counts.  Same values, so same results.
SGD void lane_save_ray(const Lane& L, const Stack& S) {
    float* p = S.s_save; const int st = S.stride;
    p[0] = L.o.x; p[st] = L.o.y; p[2 * st] = L.o.z; p[3 * st] = L.inv_dir.x; p[4 * st] = L.inv_dir.y; p[5 * st] = L.inv_dir.z;
    p[6 * st] = L.rp.sx; p[7 * st] = L.rp.sy; p[8 * st] = L.rp.sz; p[9 * st] = __int_as_float(L.rp.kz);
}
SGD void lane_restore_ray(Lane& L, const Stack& S) {
    const float* p = S.s_save; const int st = S.stride;
    L.o = f3(p[0], p[st], p[2 * st]); L.inv_dir = f3(p[3 * st], p[4 * st], p[5 * st]);
    L.rp.sx = p[6 * st]; L.rp.sy = p[7 * st]; L.rp.sz = p[8 * st]; L.rp.kz = __float_as_int(p[9 * st]);
    L.rp.kx = L.rp.kz + 1; if (L.rp.kx == 3) L.rp.kx = 0;
    L.rp.ky = L.rp.kx + 1; if (L.rp.ky == 3) L.rp.ky = 0;
    L.nx = L.inv_dir.x < 0.0f; L.ny = L.inv_dir.y < 0.0f; L.nz = L.inv_dir.z < 0.0f;
}

template <bool ANY, bool COUNT>
SGD void lane_enter_instance(const TraceScene& ts, Lane& L, const Stack& S, uint32_t inst_id, float3 o, float3 d, uint32_t& n_nodes) {
    const DInstance& I = ts.instances[inst_id];
    lane_save_ray(L, S);
    float tm = L.t_max;
    instance_ray(I, !ANY || (ts.scene_flags & SG_SCENE_FIX_INSTANCING) != 0, o, d, tm);
    L.t_saved = L.t_max; L.t_max = tm; L.inst = (int)inst_id; L.sp_base = L.sp; L.inst_hit = false;
    // like lane_set_ray, but the watertight test's shear constants (three more IEEE divides) wait until a triangle of the object is
    // actually tested -- most instance visits end in the object's upper BVH levels: the direction is parked in rp.s*, kz = -1
    L.o = o;
    L.inv_dir = f3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    L.nx = L.inv_dir.x < 0.0f; L.ny = L.inv_dir.y < 0.0f; L.nz = L.inv_dir.z < 0.0f;
    L.rp.kz = -1; L.rp.sx = d.x; L.rp.sy = d.y; L.rp.sz = d.z;
    L.cur = I.root_ref;
    if (I.root_has_bounds) {
        float te;
        if (COUNT) n_nodes++;
        const bool ok = slab_entry(I.bmin[0], I.bmin[1], I.bmin[2], I.bmax[0], I.bmax[1], I.bmax[2], o, L.inv_dir, L.nx, L.ny, L.nz, te);
        if (!(ok && te < tm)) L.cur = kEmptyRef;
    }
}

// Pops the next node whose entry distance is still in range; kEmptyRef when the stack runs dry.  When an instance's
// part of the stack is exhausted the lane returns to render space (the original ray is re-read through `io`) and goes
// on with the top-level entries below.
template <bool ANY, bool COUNT, bool INST, class IO, class IdxT>
SGD uint32_t lane_pop(Lane& L, const Stack& S, uint32_t& n_nodes, IO& io, IdxT idx) {
    if constexpr (!INST) {
        while (L.sp > 0) {
            --L.sp;
            uint32_t ref; float t = 0.0f;
            S.get<ANY>(L.sp, ref, t);
            if (COUNT) n_nodes++;                                   // the reference tests the bounds at pop time
            if (COUNT && (ref & kFailBit)) continue;
            if (ANY) return ref;                                    // t_max never shrinks for the predicate
            if (t < L.t_max) return ref;
        }
        return kEmptyRef;
    } else {
        for (;;) {
            while (L.sp > L.sp_base) {
                --L.sp;
                uint32_t ref; float t = 0.0f;
                S.get<ANY>(L.sp, ref, t);
                if (COUNT) n_nodes++;
                if (COUNT && (ref & kFailBit)) continue;
                if (ANY) return ref;
                if (t < L.t_max) return ref;
            }
            if (L.inst < 0) return kEmptyRef;
            lane_restore_ray(L, S);                                 // back to the render-space ray
            if (!L.inst_hit) L.t_max = L.t_saved;                   // else t_max = si.t_hit of the instanced hit (primitive.rs:162)
            L.inst = -1; L.sp_base = 0;
        }
    }
}

// One interior step: fetch a Node64, test both children, push the far one, move to the near one
// (or pop).  Leaves `L.cur` at an interior ref, a leaf ref, or kEmptyRef (ray finished).
template <bool ANY, bool COUNT, bool INST, class IO, class IdxT>
SGD void lane_step_interior(const TraceScene& ts, Lane& L, const Stack& S, uint32_t& n_nodes, IO& io, IdxT idx) {
    const float4* nd = ts.node64 + 4 * (size_t)L.cur;
    const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
    const uint32_t ref0 = __float_as_uint(q3.x), ref1 = __float_as_uint(q3.y), axis = __float_as_uint(q3.z) & 3u;
    // near child by dir_is_neg[axis], aggregate.rs:119-127 (a shift instead of a two-level select: the select compiled to a divergent branch)
    const int neg = (int)((((uint32_t)L.nx | ((uint32_t)L.ny << 1) | ((uint32_t)L.nz << 2)) >> axis) & 1u);
    const uint32_t near_ref = neg ? ref1 : ref0, far_ref = neg ? ref0 : ref1;
#ifndef SG_TRACE_PREFETCH
#define SG_TRACE_PREFETCH 0     /* measured on C2: -4 % (r02_sweep2); a build-time switch so that the hot loop carries no test for it */
#endif
    if (SG_TRACE_PREFETCH) {
        // start pulling the near child's node (or triangle) towards L1 while the two slab tests run
        const float4* nxt = (near_ref & kLeafBit) ? ts.tri_verts + 3 * (size_t)(near_ref & ~kLeafBit) : ts.node64 + 4 * (size_t)near_ref;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(nxt));
    }
    float t0, t1;
    const bool ok0 = slab_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, L.o, L.inv_dir, L.nx, L.ny, L.nz, t0);
    const bool ok1 = slab_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, L.o, L.inv_dir, L.nx, L.ny, L.nz, t1);
    const bool near_ok = neg ? ok1 : ok0, far_ok = neg ? ok0 : ok1;
    const float near_t = neg ? t1 : t0, far_t = neg ? t0 : t1;
    // The reference tests the far child against t_max when it is POPPED, and t_max is not
    // monotone: intersect_triangle compares t_scaled with t_max*det, so an accepted hit can
    // round to a t up to (1+2^-24)^3 above the previous t_max (vertex/edge ties).  Pruning at
    // push time must therefore leave slack; 2^-10 covers > 5000 successive tie increases and
    // costs almost no extra pushes.  The decisive `entry < t_max` test is made on pop.
    const bool far_take = far_ok && (ANY ? far_t < L.t_max : far_t <= L.t_max * 1.0009765625f);
    if (far_take || COUNT) {
        S.put<ANY>(L.sp, far_take ? far_ref : (far_ref | kFailBit), far_t);
        L.sp++;
    }
    if (COUNT) n_nodes++;                                               // near child's bounds test
    if (near_ok && near_t < L.t_max) L.cur = near_ref;
    else L.cur = lane_pop<ANY, COUNT, INST>(L, S, n_nodes, io, idx);
}

// The leaf's primitives (aggregate.rs:99-110), then pop.  Returns true when an any-hit ray is done.
template <bool ANY, bool COUNT, bool INST, class IO, class IdxT>
SGD void lane_step_leaf(const TraceScene& ts, Lane& L, const Stack& S, uint32_t& n_nodes, uint32_t& n_tris, IO& io, IdxT idx) {
    uint32_t pi = L.cur & ~kLeafBit;
    for (;;) {
        const float4* tv = ts.tri_verts + 3 * (size_t)pi;
        const float4 v0 = __ldg(tv), v1 = __ldg(tv + 1), v2 = __ldg(tv + 2);
        if constexpr (INST) if (((__float_as_uint(v0.w) >> 28) & 7u) == kKindInstance) {
            // a TransformedPrimitive: the rest of this leaf (if any) resumes from the stack after the instance
            if (!(__float_as_uint(v2.w) & kLastInLeaf)) {
                S.put<ANY>(L.sp, kLeafBit | (pi + 1), -INFINITY); L.sp++;
                if (COUNT) n_nodes--;                               // the resume entry is not a node visit
            }
            float3 o, d; float tm;
            io.load(idx, o, d, tm);
            lane_enter_instance<ANY, COUNT>(ts, L, S, __float_as_uint(v1.w), o, d, n_nodes);
            if (L.cur == kEmptyRef) L.cur = lane_pop<ANY, COUNT, INST>(L, S, n_nodes, io, idx);
            return;
        }
        if (COUNT) n_tris++;
        float b0, b1, b2, t;
        bool hit_prim;
        if (INST && (__float_as_uint(v2.w) & kSphereBit)) {
            // Shape::Sphere: works on the current space's ray (the render-space ray re-read through `io`, moved into the instance
            // when the lane is inside one); the hit record carries p_obj
            float3 o, d; float tm;
            io.load(idx, o, d, tm);
            if (L.inst >= 0) instance_ray(ts.instances[L.inst], !ANY || (ts.scene_flags & SG_SCENE_FIX_INSTANCING) != 0, o, d, tm);
            float3 p_obj;
            hit_prim = sphere_basic_intersect(ts.spheres[__float_as_uint(v2.w) & ~(kSphereBit | kLastInLeaf)], o, d, L.t_max, p_obj, t);
            b0 = p_obj.x; b1 = p_obj.y; b2 = p_obj.z;
        } else if (INST && (__float_as_uint(v2.w) & kPatchBit)) {
            // Shape::BilinearPatch: hit record carries (u, v)
            float3 o, d; float tm;
            io.load(idx, o, d, tm);
            if (L.inst >= 0) instance_ray(ts.instances[L.inst], !ANY || (ts.scene_flags & SG_SCENE_FIX_INSTANCING) != 0, o, d, tm);
            const float4* pv = ts.patch_verts + 4 * (size_t)(__float_as_uint(v2.w) & ~(kPatchBit | kLastInLeaf));
            const float4 a0 = __ldg(pv), a1 = __ldg(pv + 1), a2 = __ldg(pv + 2), a3 = __ldg(pv + 3);
            hit_prim = intersect_blp(o, d, L.t_max, f3(a0.x, a0.y, a0.z), f3(a1.x, a1.y, a1.z), f3(a2.x, a2.y, a2.z), f3(a3.x, a3.y, a3.z), b0, b1, t);
            b2 = 0.0f;
        } else {
            if constexpr (INST) if (L.rp.kz < 0) L.rp = ray_precompute(f3(L.rp.sx, L.rp.sy, L.rp.sz));    // deferred by lane_enter_instance
            hit_prim = !(__float_as_uint(v0.w) & kDegenerateBit) && intersect_triangle<false>(L.o, L.rp, L.t_max, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), b0, b1, b2, t);
        }
        if (hit_prim) {
            L.hit.prim = (int)pi; L.hit.t = t; L.hit.b0 = b0; L.hit.b1 = b1; L.hit.b2 = b2;
            if constexpr (INST) { L.hit.inst = L.inst; L.inst_hit = true; }
            if (ANY) { L.cur = kEmptyRef; return; }
            L.t_max = t;
        }
        if (__float_as_uint(v2.w) & kLastInLeaf) break;
        ++pi;
    }
    L.cur = lane_pop<ANY, COUNT, INST>(L, S, n_nodes, io, idx);
}

// ---- postponed-leaf variant (triangle-only kernels) -------------------------------------------------------------
// A lane that reaches a leaf does not stop for it: the leaf is parked in `pend` (with its box entry distance) and the lane
// keeps traversing; only a SECOND leaf blocks it (that one goes back on the stack and is re-popped, and thereby re-tested
// against the then-current t_max, after the parked leaf has been tested).  Leaves are still tested in exactly the
// reference's order, and a parked leaf is re-validated with `entry < t_max` before its triangles are tested -- a leaf's
// entry distance is >= every ancestor's (child bounds are subsets and the slab arithmetic is monotone), so that one
// comparison is equivalent to the reference having culled any box on the way down with the updated t_max.  Hits, ties
// and `t` are therefore identical to aggregate.rs:71-203; only the set of boxes looked at grows a little (stale t_max
// while a leaf is parked), which is why the COUNT builds (reference-order visit counters) keep the in-order loop.
static constexpr uint32_t kBlockedRef = 0x7ffffffeu;

template <bool ANY>
SGD uint32_t lane_pop_t(Lane& L, const Stack& S, float& t_out) {
    while (L.sp > 0) {
        --L.sp;
        uint32_t ref; float t = 0.0f;
        S.get<ANY>(L.sp, ref, t);
        if (ANY || t < L.t_max) { t_out = t; return ref; }
    }
    return kEmptyRef;
}
// Park / block on leaves until `ref` is an interior node, kBlockedRef or kEmptyRef.
template <bool ANY>
SGD void lane_settle(Lane& L, const Stack& S, uint32_t ref, float t, uint32_t& pend, float& pend_t) {
    while (ref & kLeafBit) {
        if (pend == kEmptyRef) { pend = ref; pend_t = t; ref = lane_pop_t<ANY>(L, S, t); }
        else { S.put<ANY>(L.sp, ref, t); L.sp++; ref = kBlockedRef; }
    }
    L.cur = ref;
}
template <bool ANY>
SGD void lane_step_interior_post(const TraceScene& ts, Lane& L, const Stack& S, uint32_t& pend, float& pend_t) {
    const float4* nd = ts.node64 + 4 * (size_t)L.cur;
    const float4 q0 = __ldg(nd), q1 = __ldg(nd + 1), q2 = __ldg(nd + 2), q3 = __ldg(nd + 3);
    const uint32_t ref0 = __float_as_uint(q3.x), ref1 = __float_as_uint(q3.y), axis = __float_as_uint(q3.z) & 3u;
    const int neg = axis == 0 ? L.nx : (axis == 1 ? L.ny : L.nz);
    const uint32_t near_ref = neg ? ref1 : ref0, far_ref = neg ? ref0 : ref1;
    float t0, t1;
    const bool ok0 = slab_entry(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, L.o, L.inv_dir, L.nx, L.ny, L.nz, t0);
    const bool ok1 = slab_entry(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, L.o, L.inv_dir, L.nx, L.ny, L.nz, t1);
    const bool near_ok = neg ? ok1 : ok0, far_ok = neg ? ok0 : ok1;
    float near_t = neg ? t1 : t0; const float far_t = neg ? t0 : t1;
    if (far_ok && (ANY ? far_t < L.t_max : far_t <= L.t_max * 1.0009765625f)) { S.put<ANY>(L.sp, far_ref, far_t); L.sp++; }   // see lane_step_interior
    uint32_t nxt = near_ref;
    if (!(near_ok && near_t < L.t_max)) nxt = lane_pop_t<ANY>(L, S, near_t);
    lane_settle<ANY>(L, S, nxt, near_t, pend, pend_t);
}
// Tests the parked leaf's primitives (aggregate.rs:99-110).  Returns true when an any-hit ray is done.
template <bool ANY>
SGD bool lane_test_pending(const TraceScene& ts, Lane& L, uint32_t pend, float pend_t) {
    if (!ANY && !(pend_t < L.t_max)) return false;                      // the reference would have culled its box by now
    uint32_t pi = pend & ~kLeafBit;
    for (;;) {
        const float4* tv = ts.tri_verts + 3 * (size_t)pi;
        const float4 v0 = __ldg(tv), v1 = __ldg(tv + 1), v2 = __ldg(tv + 2);
        float b0, b1, b2, t;
        if (intersect_triangle(L.o, L.rp, L.t_max, f3(v0.x, v0.y, v0.z), f3(v1.x, v1.y, v1.z), f3(v2.x, v2.y, v2.z), b0, b1, b2, t)) {
            L.hit.prim = (int)pi; L.hit.t = t; L.hit.b0 = b0; L.hit.b1 = b1; L.hit.b2 = b2;
            if (ANY) return true;
            L.t_max = t;
        }
        if (__float_as_uint(v2.w) & kLastInLeaf) break;
        ++pi;
    }
    return false;
}

}  // namespace sg
