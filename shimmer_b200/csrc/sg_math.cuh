// Device arithmetic layer of the B200 wavefront path tracer.
//
// Arithmetic contract (DESIGN.md "Arithmetic contract"): this translation unit is compiled
// with -fmad=false, IEEE division and square root (nvcc defaults -prec-div=true
// -prec-sqrt=true), no fast-math.  A fused multiply-add appears ONLY where the reference
// calls f32::mul_add: difference_of_products (math.rs:173-178) and dot3
// (vecmath/tuple_fns.rs:68-79).  Everything else is separately rounded mul/add in the
// reference's evaluation order, so that first-hit indices, t and barycentrics are
// bit-identical to the CPU path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define SGD __device__ __forceinline__

namespace sg {

static constexpr float kPi = 3.14159265358979323846f;
static constexpr float kInvPi = 0.31830988618379067154f;
static constexpr float kPiOver2 = 1.57079632679489661923f;
static constexpr float kPiOver4 = 0.78539816339744830961f;
static constexpr float kMachineEps = 1.1920929e-07f * 0.5f;     // float.rs:16
// gamma(n) = n*eps/2 / (1 - n*eps/2), float.rs:88-90 -- evaluated in f32 exactly like the reference
SGD float gamma_n(int n) { return ((float)n * kMachineEps) / (1.0f - (float)n * kMachineEps); }

SGD float next_up(float v) {                                    // float.rs:53-68
    if (isinf(v) && v > 0.0f) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t u = __float_as_uint(v);
    u = (v >= 0.0f) ? u + 1u : u - 1u;
    return __uint_as_float(u);
}
SGD float next_down(float v) {                                  // float.rs:72-86
    if (isinf(v) && v < 0.0f) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t u = __float_as_uint(v);
    u = (v > 0.0f) ? u - 1u : u + 1u;
    return __uint_as_float(u);
}
// NaN-preserving variants for the interval arithmetic of the sphere (sg_sphere.cuh): the reference's bit increment keeps
// x86's / ARM's default NaN (0xFFC00000 / 0x7FC00000) a NaN, whereas CUDA's canonical NaN 0x7FFFFFFF + 1 would wrap to
// -0.0; Sphere::basic_intersect relies on f32::min/max ignoring NaN interval bounds for axis-aligned rays.  The hot
// shading paths (offset_ray_origin, Point3fi construction) keep the two-instruction-cheaper versions above: they only see
// NaN when the path is already degenerate.
SGD float next_up_n(float v) { return isnan(v) ? v : next_up(v); }
SGD float next_down_n(float v) { return isnan(v) ? v : next_down(v); }
SGD float sqr(float x) { return x * x; }
SGD float clampf(float x, float lo, float hi) { float r = x; if (r < lo) r = lo; if (r > hi) r = hi; return r; }   // f32::clamp
SGD float dop(float a, float b, float c, float d) {             // math.rs:173-178
    float cd = c * d;
    float diff = fmaf(a, b, -cd);
    float err = fmaf(-c, d, cd);
    return diff + err;
}
SGD double dop_d(double a, double b, double c, double d) {      // math.rs:190-195
    double cd = c * d;
    double diff = fma(a, b, -cd);
    double err = fma(-c, d, cd);
    return diff + err;
}
SGD float sop(float a, float b, float c, float d) { return dop(a, b, -c, d); }     // math.rs:182-184
SGD float lerpf(float t, float a, float b) { return a * (1.0f - t) + b * t; }      // math.rs:246-252
SGD float safe_asin(float x) { return asinf(clampf(x, -1.0f, 1.0f)); }             // math.rs:266-268
SGD float safe_sqrt(float x) { return sqrtf(fmaxf(0.0f, x)); }                     // math.rs:278-281

// ---- float3 helpers (Point3f / Vector3f / Normal3f all map to float3) ----
SGD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SGD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SGD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SGD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SGD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SGD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
SGD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }   // three true divisions, vector.rs:1218
SGD float3 abs3(float3 a) { return f3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
SGD float dot3(float3 v, float3 w) { return fmaf(v.x, w.x, sop(v.y, w.y, v.z, w.z)); }   // tuple_fns.rs:68-79
SGD float absdot3(float3 v, float3 w) { return fabsf(dot3(v, w)); }
SGD float3 cross3(float3 a, float3 b) {                         // tuple_fns.rs:40-52
    return f3(dop(a.y, b.z, a.z, b.y), dop(a.z, b.x, a.x, b.z), dop(a.x, b.y, a.y, b.x));
}
SGD float len2(float3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }              // length_fns.rs:6-13 (unfused)
SGD float len3(float3 v) { return sqrtf(len2(v)); }
SGD float3 normalize3(float3 v) { float l = len3(v); return v / l; }               // normalize.rs:9-13
SGD float dist2(float3 a, float3 b) { return len2(a - b); }
SGD float maxcomp(float3 v) { return fmaxf(v.x, fmaxf(v.y, v.z)); }                 // tuple.rs:180-182
SGD int maxcomp_index(float3 v) {                                                   // tuple.rs:184-198
    if (v.x > v.y) return v.x > v.z ? 0 : 2;
    return v.y > v.z ? 1 : 2;
}
SGD float comp3(float3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }
SGD float3 faceforward3(float3 a, float3 b) { return dot3(a, b) < 0.0f ? -a : a; }  // tuple_fns.rs:208-220
SGD float3 gram_schmidt3(float3 v, float3 w) { return v - dot3(v, w) * w; }         // vector.rs:517-519
SGD void coord_system(float3 v, float3& a2, float3& a3) {                           // vector.rs:1034-1042
    float sign = copysignf(1.0f, v.z);
    float a = -1.0f / (sign + v.z);
    float b = v.x * v.y * a;
    a2 = f3(1.0f + sign * sqr(v.x) * a, sign * b, -sign * v.x);
    a3 = f3(b, sign + sqr(v.y) * a, -v.y);
}
SGD float angle_between3(float3 v1, float3 v2) {                                    // tuple_fns.rs:162-183
    if (dot3(v1, v2) < 0.0f) return kPi - 2.0f * safe_asin(len3(v1 + v2) / 2.0f);
    return 2.0f * safe_asin(len3(v2 - v1) / 2.0f);
}
SGD float spherical_tri_area(float3 a, float3 b, float3 c) {                        // spherical.rs:5-7
    return fabsf(2.0f * atan2f(dot3(a, cross3(b, c)), 1.0f + dot3(a, b) + dot3(a, c) + dot3(b, c)));
}

// ---- Point3fi as (lo, hi) per axis: interval.rs:48-58,68-84, point.rs:911-919,1000-1026 ----
struct P3fi { float3 lo, hi; };
SGD void ival(float v, float e, float& lo, float& hi) {
    if (e == 0.0f) { lo = v; hi = v; } else { lo = next_down(v - e); hi = next_up(v + e); }
}
SGD P3fi p3fi_make(float3 p, float3 e) {
    P3fi r;
    ival(p.x, e.x, r.lo.x, r.hi.x); ival(p.y, e.y, r.lo.y, r.hi.y); ival(p.z, e.z, r.lo.z, r.hi.z);
    return r;
}
SGD P3fi p3fi_exact(float3 p) { P3fi r; r.lo = p; r.hi = p; return r; }
SGD float3 p3fi_mid(const P3fi& p) { return f3((p.lo.x + p.hi.x) / 2.0f, (p.lo.y + p.hi.y) / 2.0f, (p.lo.z + p.hi.z) / 2.0f); }
SGD float3 p3fi_err(const P3fi& p) { return f3((p.hi.x - p.lo.x) / 2.0f, (p.hi.y - p.lo.y) / 2.0f, (p.hi.z - p.lo.z) / 2.0f); }

SGD float3 offset_ray_origin(const P3fi& pi, float3 n, float3 w) {                  // ray.rs:53-72
    float d = dot3(abs3(n), p3fi_err(pi));
    float3 off = d * n;
    if (dot3(w, n) < 0.0f) off = -off;
    float3 po = p3fi_mid(pi) + off;
    if (off.x > 0.0f) po.x = next_up(po.x); else if (off.x < 0.0f) po.x = next_down(po.x);
    if (off.y > 0.0f) po.y = next_up(po.y); else if (off.y < 0.0f) po.y = next_down(po.y);
    if (off.z > 0.0f) po.z = next_up(po.z); else if (off.z < 0.0f) po.z = next_down(po.z);
    return po;
}

// ---- xoshiro256++ (rand 0.8.5 SmallRng on 64-bit targets; sampler.rs:103-132) ----
struct Rng {
    uint64_t s0, s1, s2, s3;
    SGD static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    SGD static uint64_t splitmix_step(uint64_t& state) {
        state += 0x9e3779b97f4a7c15ULL;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    SGD void seed_from_u64(uint64_t state) {        // SeedableRng::seed_from_u64 (SplitMix64 fill)
        s0 = splitmix_step(state); s1 = splitmix_step(state); s2 = splitmix_step(state); s3 = splitmix_step(state);
    }
    SGD uint64_t next_u64() {
        uint64_t r = rotl(s0 + s3, 23) + s0;
        uint64_t t = s1 << 17;
        s2 ^= s0; s3 ^= s1; s1 ^= s2; s0 ^= s3; s2 ^= t; s3 = rotl(s3, 45);
        return r;
    }
    // Standard f32: 24 high bits of next_u32() (= next_u64() >> 32) scaled by 2^-24
    SGD float get_1d() { return (float)((uint32_t)(next_u64() >> 32) >> 8) * (1.0f / 16777216.0f); }
};
SGD uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// (pixel, sample) -> stream key; DESIGN.md "RNG streams".  The reference has no canonical
// map (sampler.rs:117-121 is a no-op), so the map is ours and is shared with the CPU oracle.
SGD uint64_t stream_key(uint64_t seed, uint32_t pixel_index, uint32_t sample_index) {
    return mix64(mix64(seed + 0x9e3779b97f4a7c15ULL) ^ (((uint64_t)pixel_index << 32) | (uint64_t)sample_index));
}

// Rust `f as i32`: saturating, NaN -> 0
SGD int f2i_sat(float f) { return __float2int_rz(f); }   // cvt.rzi.s32.f32 saturates and maps NaN to 0

}  // namespace sg
