// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this code.  The product (libshimmer_gpu.so) never links or calls it.
//
// CPU restatement of jalberse/shimmer's arithmetic layer (L1 in SURVEY.md section 1).
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src).  Compile with -ffp-contract=off: the reference is Rust, which
// never contracts a*b+c; it fuses only where it calls f32::mul_add explicitly.
//
// PARITY STATUS: the reference cannot be built here (no cargo/rustc, nightly-only
// crate, un-vendored crates.io deps, missing rgb2spec blobs -- SURVEY.md 8c), so
// this restatement is pinned against the reference's own unit-test vectors
// (tests/test_oracle_kat.py) and, for third-party crates (rand SmallRng,
// num-complex), against their published algorithms: for those "parity unpinned".
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <algorithm>

namespace orc {

typedef float Float;                                   // float.rs:1-4 (use_f64 never set)
static const Float PI_F = 3.14159265358979323846f;     // float.rs:14
static const Float INV_PI = 0.31830988618379067154f;   // math.rs consts
static const Float INV_4PI = 0.07957747154594766788f;
static const Float INV_2PI = 0.15915494309189533577f;
static const Float PI_OVER_2 = 1.57079632679489661923f;
static const Float PI_OVER_4 = 0.78539816339744830961f;
static const Float F_INF = std::numeric_limits<Float>::infinity();
static const Float MACHINE_EPSILON = 1.1920929e-07f * 0.5f;  // float.rs:16 (f32::EPSILON * 0.5)

inline uint32_t float_to_bits(Float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }  // float.rs:24-34
inline Float bits_to_float(uint32_t u) { Float f; std::memcpy(&f, &u, 4); return f; }      // float.rs:42-49

// float.rs:53-68
inline Float next_float_up(Float v) {
    if (std::isinf(v) && v > 0.0f) return v;
    if (v == -0.0f) v = 0.0f;
    uint32_t ui = float_to_bits(v);
    if (v >= 0.0f) ui += 1; else ui -= 1;
    return bits_to_float(ui);
}
// float.rs:72-86
inline Float next_float_down(Float v) {
    if (std::isinf(v) && v < 0.0f) return v;
    if (v == 0.0f) v = -0.0f;
    uint32_t ui = float_to_bits(v);
    if (v > 0.0f) ui -= 1; else ui += 1;
    return bits_to_float(ui);
}
// float.rs:88-90
inline Float gamma_n(int n) { return ((Float)n * MACHINE_EPSILON) / (1.0f - (Float)n * MACHINE_EPSILON); }

inline Float sqr(Float x) { return x * x; }
// Rust f32::max / f32::min ignore a NaN operand == fmaxf/fminf.
inline Float fmax_(Float a, Float b) { return std::fmax(a, b); }
inline Float fmin_(Float a, Float b) { return std::fmin(a, b); }
// Rust f32::clamp: NaN propagates; otherwise max-then-min.
inline Float clampf(Float x, Float lo, Float hi) { Float r = x; if (r < lo) r = lo; if (r > hi) r = hi; return r; }

// math.rs:173-178
inline Float difference_of_products(Float a, Float b, Float c, Float d) {
    Float cd = c * d;
    Float difference = std::fma(a, b, -cd);
    Float error = std::fma(-c, d, cd);
    return difference + error;
}
inline double difference_of_products_d(double a, double b, double c, double d) {  // math.rs:190-195
    double cd = c * d;
    double difference = std::fma(a, b, -cd);
    double error = std::fma(-c, d, cd);
    return difference + error;
}
// math.rs:182-184
inline Float sum_of_products(Float a, Float b, Float c, Float d) { return difference_of_products(a, b, -c, d); }
// math.rs:246-252
inline Float lerp(Float t, Float a, Float b) { return a * (1.0f - t) + b * t; }
// math.rs:266-268 ; :272-274 (safe_acos calls asin in the reference -- quirk, off the path)
inline Float safe_asin(Float x) { return std::asin(clampf(x, -1.0f, 1.0f)); }
// math.rs:278-281
inline Float safe_sqrt(Float x) { return std::sqrt(fmax_(0.0f, x)); }

// math.rs:322-333
template <class Pred> inline int find_interval(int size, Pred pred) {
    int first = 1, last = size - 2;
    while (last > 0) {
        int half = last >> 1, middle = first + half;
        bool r = pred(middle);
        first = r ? middle + 1 : first;
        last = r ? last - (half + 1) : half;
    }
    int v = first - 1;
    if (v < 0) v = 0;
    if (v > size - 2) v = size - 2;
    return v;
}

// ---- vecmath ---------------------------------------------------------------
struct V3 { Float x, y, z; };
struct V2 { Float x, y; };
inline V3 v3(Float x, Float y, Float z) { V3 r = {x, y, z}; return r; }
inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, Float s) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(Float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, Float s) { return v3(a.x / s, a.y / s, a.z / s); }   // vector.rs:1218 true divisions
inline V3 vabs(V3 a) { return v3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline Float comp(V3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
// tuple_fns.rs:68-79
inline Float dot(V3 v, V3 w) { return std::fma(v.x, w.x, sum_of_products(v.y, w.y, v.z, w.z)); }
inline Float abs_dot(V3 v, V3 w) { return std::fabs(dot(v, w)); }
// tuple_fns.rs:40-52
inline V3 cross(V3 a, V3 b) {
    return v3(difference_of_products(a.y, b.z, a.z, b.y),
              difference_of_products(a.z, b.x, a.x, b.z),
              difference_of_products(a.x, b.y, a.y, b.x));
}
// length_fns.rs:6-21 (unfused)
inline Float length_squared(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline Float length(V3 v) { return std::sqrt(length_squared(v)); }
// normalize.rs:9-13
inline V3 normalize(V3 v) { Float len = length(v); return v / len; }
inline Float distance_squared(V3 a, V3 b) { return length_squared(a - b); }
// tuple.rs:184-198 ; :159-161
inline int max_component_index(V3 v) {
    if (v.x > v.y) { return v.x > v.z ? 0 : 2; } else { return v.y > v.z ? 1 : 2; }
}
inline Float max_component_value(V3 v) { return fmax_(v.x, fmax_(v.y, v.z)); }
inline V3 permute(V3 v, int a, int b, int c) { return v3(comp(v, a), comp(v, b), comp(v, c)); }
// tuple_fns.rs:208-220
inline V3 face_forward(V3 a, V3 b) { return dot(a, b) < 0.0f ? -a : a; }
// vector.rs:517-519
inline V3 gram_schmidt(V3 v, V3 w) { return v - dot(v, w) * w; }
// vector.rs:1034-1042
inline void coordinate_system(V3 v, V3* v2, V3* v3o) {
    Float sign = std::copysign(1.0f, v.z);
    Float a = -1.0f / (sign + v.z);
    Float b = v.x * v.y * a;
    *v2 = v3(1.0f + sign * sqr(v.x) * a, sign * b, -sign * v.x);
    *v3o = v3(b, sign + sqr(v.y) * a, -v.y);
}
// tuple_fns.rs:162-183
inline Float angle_between(V3 v1, V3 v2) {
    if (dot(v1, v2) < 0.0f) return PI_F - 2.0f * safe_asin(length(v1 + v2) / 2.0f);
    return 2.0f * safe_asin(length(v2 - v1) / 2.0f);
}
// spherical.rs:5-7
inline Float spherical_triangle_area(V3 a, V3 b, V3 c) {
    return std::fabs(2.0f * std::atan2(dot(a, cross(b, c)), 1.0f + dot(a, b) + dot(a, c) + dot(b, c)));
}

// ---- Interval / Point3fi (interval.rs:48-58,68-84; point.rs:911-919,1000-1026) ----
struct P3fi { V3 lo, hi; };
inline void interval_from_value_and_error(Float v, Float err, Float* lo, Float* hi) {
    if (err == 0.0f) { *lo = v; *hi = v; }
    else { *lo = next_float_down(v - err); *hi = next_float_up(v + err); }
}
inline P3fi p3fi_from_value_and_error(V3 p, V3 e) {
    P3fi r;
    interval_from_value_and_error(p.x, e.x, &r.lo.x, &r.hi.x);
    interval_from_value_and_error(p.y, e.y, &r.lo.y, &r.hi.y);
    interval_from_value_and_error(p.z, e.z, &r.lo.z, &r.hi.z);
    return r;
}
inline P3fi p3fi_exact(V3 p) { P3fi r = {p, p}; return r; }
inline V3 p3fi_mid(const P3fi& p) { return v3((p.lo.x + p.hi.x) / 2.0f, (p.lo.y + p.hi.y) / 2.0f, (p.lo.z + p.hi.z) / 2.0f); }
inline V3 p3fi_error(const P3fi& p) { return v3((p.hi.x - p.lo.x) / 2.0f, (p.hi.y - p.lo.y) / 2.0f, (p.hi.z - p.lo.z) / 2.0f); }
inline bool p3fi_is_exact(const P3fi& p) { return p.hi.x - p.lo.x == 0.0f && p.hi.y - p.lo.y == 0.0f && p.hi.z - p.lo.z == 0.0f; }

// ray.rs:53-72
inline V3 offset_ray_origin(const P3fi& pi, V3 n, V3 w) {
    Float d = dot(vabs(n), p3fi_error(pi));
    V3 offset = d * n;
    if (dot(w, n) < 0.0f) offset = -offset;
    V3 po = p3fi_mid(pi) + offset;
    if (offset.x > 0.0f) po.x = next_float_up(po.x); else if (offset.x < 0.0f) po.x = next_float_down(po.x);
    if (offset.y > 0.0f) po.y = next_float_up(po.y); else if (offset.y < 0.0f) po.y = next_float_down(po.y);
    if (offset.z > 0.0f) po.z = next_float_up(po.z); else if (offset.z < 0.0f) po.z = next_float_down(po.z);
    return po;
}

// ---- rand 0.8.5 SmallRng on 64-bit = xoshiro256++ (third-party, Cargo.lock; parity unpinned) ----
struct Rng {
    uint64_t s[4];
    static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    // SeedableRng::seed_from_u64 for Xoshiro256PlusPlus: SplitMix64 fill
    void seed_from_u64(uint64_t state) {
        for (int i = 0; i < 4; ++i) {
            state += 0x9e3779b97f4a7c15ULL;
            uint64_t z = state;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    uint64_t next_u64() {
        uint64_t result = rotl(s[0] + s[3], 23) + s[0];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
        s[2] ^= t; s[3] = rotl(s[3], 45);
        return result;
    }
    uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
    // rand::distributions::Standard for f32: 24 high bits * 2^-24; sampler.rs:123-125
    Float get_1d() { return (Float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};
inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// Deterministic (pixel, sample) -> stream map.  The reference has none
// (`start_pixel_sample` is a no-op, sampler.rs:117-121, and every rayon thread
// clones the same generator, integrator.rs:252-253), so this is OUR choice,
// shared verbatim by the CUDA path: DESIGN.md "RNG streams".
inline uint64_t stream_key(uint64_t seed, uint32_t pixel_index, uint32_t sample_index) {
    return mix64(mix64(seed + 0x9e3779b97f4a7c15ULL) ^ (((uint64_t)pixel_index << 32) | (uint64_t)sample_index));
}

}  // namespace orc
