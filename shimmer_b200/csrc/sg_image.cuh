// MIP pyramid construction on the device: Image::generate_pyramid (image.rs:699-787) with Image::float_resize_up (:1007-1111) /
// resample_weights (:1113-1141) for images whose resolution is not a power of two.  HBM-bound streaming kernels: one thread per
// output texel channel, coalesced along x; every level is read once and written once.
// As written in the reference: resample_weights evaluates the windowed sinc at `first_pixel + 0.5` for all four taps (pbrt:
// first_pixel + j + 0.5), so after normalisation each tap weighs ~0.25.
#pragma once
#include "sg_math.cuh"

namespace sg {

struct ResampleWeight { int first_pixel; float w[4]; };

SGD float sin_over_x(float x) { if (1.0f - x * x == 1.0f) return 1.0f; return sinf(x) / x; }                 // math.rs:413-420
SGD float sincf_(float x) { return sin_over_x(kPi * x); }
SGD float windowed_sinc(float x, float radius, float tau) { if (fabsf(x) > radius) return 0.0f; return sincf_(x) * sincf_(x / tau); }

static __global__ void k_resample_weights(int old_res, int new_res, ResampleWeight* wt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= new_res) return;
    const float filter_radius = 2.0f, tau = 2.0f;
    const float center = ((float)i + 0.5f) * (float)old_res / (float)new_res;
    ResampleWeight r;
    r.first_pixel = max(0, f2i_sat(floorf(center - filter_radius + 0.5f)));
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float pos = (float)r.first_pixel + 0.5f; r.w[j] = windowed_sinc(pos - center, filter_radius, tau); }   // sic: no `+ j`
    const float inv = 1.0f / (r.w[0] + r.w[1] + r.w[2] + r.w[3]);
#pragma unroll
    for (int j = 0; j < 4; ++j) r.w[j] *= inv;
    wt[i] = r;
}
SGD int remap_coord(int p, int res, int wrap) {                                                           // remap_pixel_coords image.rs:134-177
    if (p >= 0 && p < res) return p;
    if (wrap == SG_WRAP_CLAMP) return p < 0 ? 0 : res - 1;
    const int r = p - (p / res) * res;                                                                    // repeat
    return r < 0 ? r + res : r;
}
static __global__ void k_resize_up(const float* __restrict__ in, int rx, int ry, int nc, int wrap, const ResampleWeight* __restrict__ xw,
                                   const ResampleWeight* __restrict__ yw, int nx, int ny, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nc) return;
    const int c = (int)(i % nc); const long long px = i / nc;
    const int x = (int)(px % nx), y = (int)(px / nx);
    const ResampleWeight wx = xw[x], wy = yw[y];
    int xs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xs[k] = remap_coord(wx.first_pixel + k, rx, wrap);
    float col[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float* row = in + (size_t)remap_coord(wy.first_pixel + j, ry, wrap) * rx * nc + c;
        col[j] = wx.w[0] * __ldg(row + (size_t)xs[0] * nc) + wx.w[1] * __ldg(row + (size_t)xs[1] * nc) + wx.w[2] * __ldg(row + (size_t)xs[2] * nc) +
                 wx.w[3] * __ldg(row + (size_t)xs[3] * nc);
    }
    out[i] = fmaxf(0.0f, wy.w[0] * col[0] + wy.w[1] * col[1] + wy.w[2] * col[2] + wy.w[3] * col[3]);
}
// one 2x2 box-filter step, image.rs:733-768
static __global__ void k_downsample(const float* __restrict__ in, int rx, int ry, int nc, int nx, int ny, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)nx * ny * nc) return;
    const int c = (int)(i % nc); const long long px = i / nc;
    const int x = (int)(px % nx), y = (int)(px / nx);
    const size_t d1 = rx == 1 ? 0 : (size_t)nc, d2 = ry == 1 ? 0 : (size_t)nc * rx;
    const size_t src = ((size_t)(2 * y) * rx + 2 * x) * nc + c;
    out[i] = 0.25f * (__ldg(in + src) + __ldg(in + src + d1) + __ldg(in + src + d2) + __ldg(in + src + d1 + d2));
}

}  // namespace sg
